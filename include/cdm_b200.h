/* cdm_b200 — C ABI of the B200-native ContextUnet / DDPM hot path.
 *
 * The reference (Tengis0618/CAMELS-Diffusion-Model) has no FFI of its own: its
 * only seam is the Python nn.Module / function API (ContextUnet.py:5-60,
 * code/diffusion_utilities.py:13-145, code/train_diffusion_paper.py:320,548,556).
 * Each entry point below replaces the torch (ATen/cuDNN) call the reference
 * makes at the cited line; the Python mirror in camels-diffusion-model_b200/
 * binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the name ends in _host; `stream` is a cudaStream_t passed as void*.
 *   - activations are NHWC bf16; statistics, schedule scalars, x_t, eps, noise
 *     are fp32.  The library allocates nothing: the caller owns all buffers.
 *   - return 0 on success, negative cdm_status on failure; cdm_last_error()
 *     gives a thread-local message.  No CPU fallback exists: on a device that
 *     is not sm_100 every compute entry point returns CDM_ERR_ARCH.
 *   - all launches are asynchronous on `stream` and CUDA-graph capturable.
 */
#ifndef CDM_B200_H
#define CDM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CDM_OK = 0,
  CDM_ERR_ARG = -1,    /* bad shape / null pointer / unsupported size */
  CDM_ERR_ARCH = -2,   /* device is not sm_100 */
  CDM_ERR_CUDA = -3,   /* CUDA runtime / driver error (see cdm_last_error) */
} cdm_status;

int cdm_version(void);
const char* cdm_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x). */
int cdm_device_ok(void);

/* ---- epilogue flags for the implicit-GEMM kernels ------------------------ */
#define CDM_EPI_RELU 1      /* max(y,0) after scale/shift                      */
#define CDM_EPI_SHORTCUT 2  /* y += w_c[co]*x[n,px] + b_c[co]   (fresh 1x1 shortcut) */
#define CDM_EPI_POOL 4      /* 2x2 max-pool, output [n][H/2][W/2][cout]        */
#define CDM_EPI_FILM 8      /* y = film_scale[n][co]*y + film_shift[..][co]    */
#define CDM_EPI_GNSTATS 16  /* emit per-(n,slot,group) sum / sum-of-squares    */

/* A-operand feeding strategy of cdm_conv3x3. */
#define CDM_CONV_MODE_COPIES 0  /* three kw-shifted TMA copies, aligned views  */
#define CDM_CONV_MODE_SHIFT24 1 /* one halo tile, pitch 24, row-shifted views  */
#define CDM_CONV_MODE_SHIFT18 2 /* one halo tile, pitch 18, row-shifted views  */

/* 3x3, stride 1, pad 1 convolution as tcgen05 implicit GEMM.
 * Replaces nn.Conv2d(+BatchNorm2d eval +ReLU) of ResidualConvBlock
 * (code/diffusion_utilities.py:26-37,42-47), the MaxPool2d of UnetDown (:109),
 * the FiLM of ContextUnet.forward (ContextUnet.py:57-58), torch.cat + out.0
 * (ContextUnet.py:36,59) and the fresh 1x1 shortcut (diffusion_utilities.py:54).
 *   y[n,h,w,co] = scale[co] * sum_{kh,kw,ci} in[n,h+kh-1,w+kw-1,ci] * weight[co,kh,kw,ci] + shift[co]
 * `in` is the channel concatenation of src0 (c0 channels) and src1 (c1). */
typedef struct {
  const void* src0;   /* bf16 [n_img][H][W][c0] */
  const void* src1;   /* bf16 [n_img][H][W][c1] or NULL */
  int c0, c1;         /* multiples of 64 */
  int n_img, H, W;    /* H, W multiples of 16 */
  const void* weight; /* bf16 [cout][3][3][c0+c1] */
  int cout;           /* multiple of 128 */
  const float* scale; /* fp32 [cout] */
  const float* shift; /* fp32 [cout] */
  int flags;          /* CDM_EPI_* */
  void* out;          /* bf16 [n_img][H'][W'][cout]  ([sc_reps*n_img].. with CDM_EPI_SHORTCUT) */
  /* CDM_EPI_SHORTCUT: image n fans out to sc_reps outputs, out[r*n_img + n] uses shortcut row r
   * (the two classifier-free-guidance passes share x and init_conv; only the shortcut differs) */
  const float* sc_x;   /* fp32 [n_img][H][W] */
  int sc_reps;
  const float* sc_tab; /* fp32 [steps][sc_reps][2][cout]: w_c then b_c */
  /* CDM_EPI_FILM */
  const float* film_scale; /* fp32 [n_img][cout] */
  const float* film_shift; /* fp32 [steps][film_shift_rows][cout] */
  int film_shift_rows;     /* 1 (t shared by the batch) or n_img */
  /* device int32 step index used to pick the sc_tab / film_shift row; NULL = 0 */
  const int* step_ptr;
  /* CDM_EPI_GNSTATS: fp32 [n_img][(H/16)*(W/16)*8][8][2] */
  float* gn_partial;
  int mode; /* CDM_CONV_MODE_* */
} cdm_conv3x3_args;
int cdm_conv3x3(const cdm_conv3x3_args* a, void* stream);

/* Dense GEMM  C[m, n] = sum_k A[m,k] * Bw[n,k] + shift[n % shift_mod], bf16 in,
 * fp32 accumulate, bf16 out.  A is the row concatenation along K of a0 (k0
 * columns) and a1 (k1).  Replaces nn.ConvTranspose2d of UnetUp
 * (code/diffusion_utilities.py:86, with torch.cat :96) and of up0
 * (ContextUnet.py:27).
 *   out_mode 0: C row-major [M][N]                               (up0)
 *   out_mode 1: 2x2 stride-2 pixel shuffle: row m = (img,h,w) of an
 *               [n_img][H][W] grid, n = (kh*2+kw)*128 + co  ->
 *               out[img][2h+kh][2w+kw][co], N must be 512        (UnetUp) */
typedef struct {
  const void* a0; /* bf16 [M][k0] */
  const void* a1; /* bf16 [M][k1] or NULL */
  int k0, k1;     /* multiples of 64 */
  int M, N;       /* N multiple of 128 */
  const void* bw; /* bf16 [N][k0+k1] */
  const float* shift;
  int shift_mod;
  int out_mode;
  int H, W; /* out_mode 1 */
  void* out;
} cdm_gemm_args;
int cdm_gemm(const cdm_gemm_args* a, void* stream);

/* init_conv.conv1: nn.Conv2d(1, cout, 3, 1, 1) + eval BatchNorm2d + ReLU
 * (code/diffusion_utilities.py:26-30 with in_channels=1, ContextUnet.py:14). fp32 math. */
typedef struct {
  const float* x; /* fp32 [n_img][H][W] */
  int n_img, H, W;
  const float* weight; /* fp32 [9][cout], tap-major */
  int cout;
  const float* scale; /* fp32 [cout] */
  const float* shift; /* fp32 [cout] */
  int relu;
  void* out; /* bf16 [n_img][H][W][cout] */
} cdm_conv_in_args;
int cdm_conv_in(const cdm_conv_in_args* a, void* stream);

/* out.1-out.3: GroupNorm(8, C) + ReLU applied on load, then nn.Conv2d(C, 1, 3, 1, 1)
 * (ContextUnet.py:37-39). src is the raw out.0 output; eps-prediction comes out in fp32. */
typedef struct {
  const void* src; /* bf16 [n_img][H][W][C] */
  int n_img, H, W, C; /* C == 128 */
  const float* mean_rstd; /* fp32 [n_img][8][2] from cdm_gn_finalize */
  const float* gamma;     /* fp32 [C] */
  const float* beta;      /* fp32 [C] */
  const float* weight;    /* fp32 [9][C], tap-major */
  const float* bias;      /* fp32 [1] */
  float* out;             /* fp32 [n_img][H][W] */
} cdm_conv_out_args;
int cdm_conv_out(const cdm_conv_out_args* a, void* stream);

/* EmbedFC.forward (code/diffusion_utilities.py:137-145): out = W2 gelu(W1 v + b1) + b2, fp32. */
int cdm_embed_fc(const float* in, int rows, int din, const float* w1, const float* b1, const float* w2,
                 const float* b2, int emb, float* out, void* stream);

/* to_vec = AvgPool2d(h/4) + GELU (ContextUnet.py:17): src bf16 [n][P][C] -> out bf16 [n][C]. */
int cdm_avgpool_gelu(const void* src, int n_img, int P, int C, void* out, void* stream);

/* up0.1-up0.2 + FiLM: GroupNorm(groups, C) + ReLU, then film_scale*y + film_shift
 * (ContextUnet.py:28-29,57).  film_* may both be NULL. */
typedef struct {
  const void* src; /* bf16 [n_img][P][C] */
  int n_img, P, C, groups;
  const float* gamma;
  const float* beta;
  float eps;
  const float* film_scale; /* fp32 [n_img][C] */
  const float* film_shift; /* fp32 [steps][film_rows][C] */
  int film_rows;           /* 1 or n_img */
  const int* step_ptr;     /* device int32 row selector, NULL = 0 */
  void* out;               /* bf16 [n_img][P][C] */
} cdm_gn_relu_film_args;
int cdm_gn_relu_film(const cdm_gn_relu_film_args* a, void* stream);

/* Deterministic reduction of the CDM_EPI_GNSTATS partials into mean / rstd per (image, group). */
int cdm_gn_finalize(const float* partial, int n_img, int slots, float count, float eps, float* mean_rstd,
                    void* stream);

/* One reverse-diffusion update: classifier-free-guidance mix + denoise_add_noise
 * (code/train_diffusion_paper.py:548-553,600-611):
 *   eps = eps_u + w (eps_c - eps_u)   if reps == 2 and guide_w > 0, else eps_c
 *   x  <- (x - eps * k2[i]) / sa[i] + sb[i] * z,     z = 0 at i == 1
 * coef[i] = {(1-a_i)/sqrt(1-ab_i), sqrt(a_i), sqrt(b_i), 0}.  Each fp32 op is rounded
 * separately so the trajectory is bit-identical to the reference's elementwise chain. */
typedef struct {
  float* x;         /* fp32 [n][hw], updated in place */
  const float* eps; /* fp32 [reps*n][hw]: conditional block first, unconditional second */
  int n, hw, reps;
  float guide_w;
  const float* coef;   /* fp32 [timesteps+1][4] */
  const int* step_ptr; /* device int32 holding i (CUDA-graph replay), or NULL -> `step` */
  int step, timesteps;
  const float* z;          /* host-fed noise: z + (timesteps - i) * z_iter_stride, or NULL -> in-kernel Philox */
  long long z_iter_stride; /* elements */
  unsigned long long seed;
  float* snap;          /* optional snapshot ring [n_snap][n][hw] ... */
  const int* snap_slot; /* ... slot per step i ([timesteps+1], -1 = none)  (paper.py:617-618) */
} cdm_ddpm_step_args;
int cdm_ddpm_step(const cdm_ddpm_step_args* a, void* stream);
int cdm_step_advance(int* step_ptr, int delta, void* stream);

/* perturb_input (code/train_diffusion_paper.py:320-321; :112 for the sqrt form):
 *   out = ca[t] * x + cb[t] * noise.  noise == NULL -> drawn in-kernel (Philox) and written to noise_out. */
typedef struct {
  const float* x;
  const float* noise;
  float* out;
  int n, hw;
  const float* ca; /* fp32 [T+1] */
  const float* cb; /* fp32 [T+1] */
  const long long* t_idx; /* int64 [n] per-sample timestep, or NULL -> t_shared */
  int t_shared;
  const int* step_ptr; /* device int32 overriding t_shared (CUDA-graph replay), or NULL */
  unsigned long long seed;
  unsigned int stream_id;
  float* noise_out;
} cdm_perturb_args;
int cdm_perturb(const cdm_perturb_args* a, void* stream);

/* F.mse_loss(reduction='none').mean([1,2,3]) (+ weighted accumulation for NLL / ELBO,
 * code/train_diffusion_paper.py:119-127,173-178): mse_out[s] = mse; acc[s] += weight_tab[t]*mse. */
typedef struct {
  const float* pred;
  const float* target;
  int n, hw;
  const float* weight_tab; /* fp32 [T+1] or NULL (weight 1) */
  const long long* t_idx;  /* int64 [n] or NULL -> t_shared */
  int t_shared;
  const int* step_ptr; /* device int32 overriding t_shared, or NULL */
  float* mse_out; /* fp32 [n] or NULL */
  float* acc;     /* fp32 [n] or NULL */
} cdm_mse_accum_args;
int cdm_mse_accum(const cdm_mse_accum_args* a, void* stream);

/* Measurement probe: every CTA streams `tile_bytes` TMA tiles from an
 * L2-resident buffer into a shared-memory ring; returns nothing, caller times it. */
int cdm_probe_tma_l2(const void* buf, int n_rows, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDM_B200_H */
