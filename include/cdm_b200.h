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
/* Multiprocessor count of the current device (148 on B200): per-CTA partial-sum workspaces are sized from it. */
int cdm_num_sms(void);

/* ---- epilogue flags for the implicit-GEMM kernels ------------------------ */
#define CDM_EPI_RELU 1      /* max(y,0) after scale/shift                      */
#define CDM_EPI_SHORTCUT 2  /* y += w_c[co]*x[n,px] + b_c[co]   (fresh 1x1 shortcut) */
#define CDM_EPI_POOL 4      /* 2x2 max-pool, output [n][H/2][W/2][cout]        */
#define CDM_EPI_FILM 8      /* y = film_scale[n][co]*y + film_shift[..][co]    */
#define CDM_EPI_GNSTATS 16  /* emit per-(n,slot,group) sum / sum-of-squares    */
#define CDM_EPI_BNSTATS 32  /* per-channel sum / sum-of-squares of the (bf16-rounded) output over the whole
                             * launch -> bn_sums, all ranks of `xr` included: the batch statistics of a
                             * train-mode nn.BatchNorm2d come out of the convolution itself (MODE 3 / 4 only) */
#define CDM_EPI_GELU 64     /* exact-erf GELU after scale/shift instead of ReLU (the activation the reference's
                             * comment names, code/diffusion_utilities.py:29,36; the code itself runs nn.ReLU) */
#define CDM_EPI_RESSCALE 128 /* y *= res_scale after the shortcut add (the reference's disabled `/ 1.414`,
                              * code/diffusion_utilities.py:59: res_scale = 1 / 1.414) */
#define CDM_EPI_BNBWD 256   /* the launch computes dy of a Conv-BatchNorm-ReLU layer (a data-gradient convolution):
                              * its epilogue also accumulates sum g and sum g*xhat per channel (g = dy under the ReLU
                              * mask of the layer's forward z) -> bn_sums, all ranks of `xr` included: what
                              * cdm_chan_reduce mode 1 computes with one more pass over dy and z (MODE 3 / 4 only) */
#define CDM_EPI_ALL 511     /* every defined bit; any other bit in `flags` is rejected with CDM_ERR_ARG */

/* A-operand feeding strategy of cdm_conv3x3. */
#define CDM_CONV_MODE_COPIES 0  /* three kw-shifted TMA copies, aligned views  */
#define CDM_CONV_MODE_SHIFT24 1 /* one halo tile, pitch 24, row-shifted views  */
#define CDM_CONV_MODE_SHIFT18 2 /* one halo tile, pitch 18, row-shifted views  */
#define CDM_CONV_MODE_SWAPPED 3 /* weights = M operand, 256 pixels = N operand (M128 N256 K16); H % 32 == 0 */
#define CDM_CONV_MODE_SWAPPED_TMA 4 /* as 3, output through shared-memory staging + TMA stores (bit-identical) */

/* 3x3, stride 1, pad 1 convolution as tcgen05 implicit GEMM.
 * Replaces nn.Conv2d(+BatchNorm2d eval +ReLU) of ResidualConvBlock
 * (code/diffusion_utilities.py:26-37,42-47), the MaxPool2d of UnetDown (:109),
 * the FiLM of ContextUnet.forward (ContextUnet.py:57-58), torch.cat + out.0
 * (ContextUnet.py:36,59) and the fresh 1x1 shortcut (diffusion_utilities.py:54).
 *   y[n,h,w,co] = scale[co] * sum_{kh,kw,ci} in[n,h+kh-1,w+kw-1,ci] * weight[co,kh,kw,ci] + shift[co]
 * `in` is the channel concatenation of src0 (c0 channels) and src1 (c1). */
struct cdm_xrank_s; /* peer group of the cross-rank exchange, defined with cdm_chan_reduce below */
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
  /* CDM_EPI_BNSTATS: workspace fp32 [cdm_num_sms()][2][cout]; bn_sums fp32 [2][cout] = (sum, sum of squares) per channel
   * over every rank of `xr` (NULL: this rank only) */
  float* bn_partial;
  float* bn_sums;
  const struct cdm_xrank_s* xr;
  float res_scale; /* CDM_EPI_RESSCALE */
  /* CDM_EPI_BNBWD: forward z of the layer whose dy this launch produces (bf16 NHWC, the shape of `out`), its
   * scale / shift (mask: z*scale+shift > 0), batch mean and rstd, fp32 [cout]; sums land in bn_partial / bn_sums */
  const void* bwd_z;
  const float* bwd_scale; const float* bwd_shift; const float* bwd_mean; const float* bwd_rstd;
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
 *               out[img][2h+kh][2w+kw][co], N must be 512        (UnetUp)
 * Kernel choice is internal and does not change results (same K order, bit-identical rows): weights resident in
 * shared memory for N = 512 with M >= 8192 and for N >= 256 * (SM count + 1) with K <= 256, M >= 1024 and
 * 256 % shift_mod == 0; a streaming kernel otherwise. */
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
  /* optional fp32 scratch: with few output tiles and a long K (M*N <= 37 tiles, K >= 4096) the K range is split over
   * the idle SMs, each slice stores its tile here and a second kernel adds the slices in a fixed order */
  float* workspace;
  long long workspace_floats;
} cdm_gemm_args;
int cdm_gemm(const cdm_gemm_args* a, void* stream);

/* init_conv.conv1: nn.Conv2d(1, cout, 3, 1, 1) + eval BatchNorm2d + ReLU
 * (code/diffusion_utilities.py:26-30 with in_channels=1, ContextUnet.py:14). fp32 math.
 * W is 16, 32 or 64, H a multiple of 4, cout a multiple of 128 (<= 512); x is 16-byte aligned (its rows are copied
 * in 16-byte chunks); anything else returns CDM_ERR_ARG. */
typedef struct {
  const float* x; /* fp32 [n_img][H][W], 16-byte aligned */
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
  float* mean_rstd_out;    /* optional fp32 [n_img][groups][2] (kept for the backward pass) */
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
  /* in-kernel noise: global index of x[0] in the whole (all-rank) sample batch.  The Philox counter is the GLOBAL
   * element index, so the draw of a sample is the same whichever rank holds it and two ranks never share noise */
  long long sample_offset;
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
  long long sample_offset; /* in-kernel noise: global index of x[0] (see cdm_ddpm_step_args) */
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
  /* optional second weighted accumulator over the same MSE: BASELINE config 5 sums mse/(2 b_t) (NLL,
   * code/train_diffusion_elbo.py:108-149) and 0.5 (1/(1-ab_t) - 1) mse (ELBO, :91-103) in one sweep */
  const float* weight_tab2; /* fp32 [T+1] or NULL (weight 1) */
  float* acc2;              /* fp32 [n] or NULL */
} cdm_mse_accum_args;
int cdm_mse_accum(const cdm_mse_accum_args* a, void* stream);

/* ======================= training path (code/train_diffusion_paper.py:349-366) =======================
 * Activations / activation gradients: NHWC bf16 with a pixel stride `ld*` (elements) so channel slices of
 * wider tensors are usable in place.  Statistics and parameter gradients: fp32.  All reductions are
 * two-stage with a fixed order (deterministic). */

/* Per-channel reductions over P rows -> out[2][C]:
 *   mode 0: sum z, sum z^2                       (train-mode nn.BatchNorm2d batch statistics)
 *   mode 1: sum g, sum g*xhat, g = dy*[relu mask] (BatchNorm2d + ReLU backward)
 *   mode 2: sum a, 0                             (bias gradients)
 * workspace: fp32 [workspace_blocks][2][C]. */
/* Peer group of the fused reduce + cross-rank exchange (data-parallel BatchNorm statistics; what the reference
 * would get from nn.SyncBatchNorm + NCCL).  Every rank owns one symmetric buffer: fp32 slots [2][world][512]
 * followed by int32 flags [world], zero-filled once; peer_slots / peer_flags are DEVICE arrays of `world` device
 * pointers to those regions on every rank (NVLink peer mappings; rank's own entry included).  seq / ticket: device
 * int32 scalars owned by the caller, zero-initialised.  NULL or world == 1: rank-local reduction. */
typedef struct cdm_xrank_s {
  int rank, world;
  const unsigned long long* peer_slots;
  const unsigned long long* peer_flags;
  int* seq;
  unsigned int* ticket;
} cdm_xrank;

typedef struct {
  const void* a; int lda;
  const void* z; int ldz;
  const float* scale; const float* shift; const float* mean; const float* rstd;
  int relu, mode;
  long long P; int C;
  float* workspace; int workspace_blocks;
  float* out;
  const cdm_xrank* xr; /* NULL: rank-local sums; else `out` = sum over all ranks, bit-identical on every rank */
} cdm_chan_reduce_args;
int cdm_chan_reduce(const cdm_chan_reduce_args* a, void* stream);
/* out[i] = sum over ranks of sum_b partial[b][i] (b in fixed order), i < n <= 512: the final pass of a two-stage
 * reduction fused with its exchange over peer memory (one kernel, graph-capturable, deterministic). */
int cdm_xrank_sum(const float* partial, int n_blocks, int n, float* out, const cdm_xrank* xr, void* stream);
/* How long a rank waits for a late peer before the exchange kernel gives up with a CUDA error (wall-clock seconds,
 * default 600 or the CDM_XRANK_TIMEOUT_S environment variable).  A late peer (checkpoint I/O, dataloader) is normal. */
int cdm_xrank_set_timeout(double seconds);

/* sums[2][C] (after the optional cross-rank all-reduce) -> scale = gamma*rstd, shift = beta - mean*scale,
 * mean, rstd; running_mean/var updated with `momentum` and the UNBIASED variance (torch semantics). */
int cdm_bn_finalize(const float* sums, int C, float count, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                    float* mean, float* rstd, void* stream);

/* y = act(z*scale+shift) [+ w_c*x + b_c (init_conv shortcut)]; optional yf = film_scale[n]*y + film_shift[n|0].
 * sums != NULL: the launch also does cdm_bn_finalize's work — scale / shift are derived from sums[2][C] (same
 * arithmetic, bit-identical), written with mean / rstd to the *_out vectors for the backward pass, and the running
 * statistics are updated (one launch per BatchNorm instead of two; `scale` / `shift` are ignored). */
typedef struct {
  const void* z; long long P; int C, relu;
  const float* scale; const float* shift;
  void* y;
  const float* sc_x; const float* sc_w; const float* sc_b;
  const float* film_scale; const float* film_shift; int film_rows, px_per_img;
  void* yf;
  const float* sums; const float* gamma; const float* beta;
  float count, eps, momentum;
  float* running_mean; float* running_var; /* both or neither */
  float* scale_out; float* shift_out; float* mean_out; float* rstd_out;
} cdm_bn_apply_args;
int cdm_bn_apply(const cdm_bn_apply_args* a, void* stream);

/* dz = scale*(g - S0/N - xhat*S1/N) with sums = {S0[C], S1[C]} from cdm_chan_reduce mode 1. */
typedef struct {
  const void* dy; int lddy;
  const void* z; long long P; int C, relu;
  const float* scale; const float* shift; const float* mean; const float* rstd;
  const float* sums; float count;
  void* dz;
} cdm_bn_bwd_args;
int cdm_bn_bwd_apply(const cdm_bn_bwd_args* a, void* stream);

/* nn.MaxPool2d(2) forward / backward (first-maximum tie rule, as torch). */
int cdm_maxpool2_fwd(const void* y, int n_img, int H, int W, int C, void* out, void* stream);
int cdm_maxpool2_bwd(const void* dpool, int lddp, const void* y, int n_img, int H, int W, int C, void* dy,
                     void* stream);
/* a += b (gradient accumulation at skip connections). */
int cdm_add_bf16(void* a, int lda, const void* b, int ldb, long long P, int C, void* stream);
/* dv[n][2H][2W][C] -> [n][H][W][(kh,kw,c)]: A operand of the ConvTranspose2d(2,2) dgrad / wgrad GEMMs. */
int cdm_space_to_depth(const void* dv, int n_img, int H, int W, int C, void* out, void* stream);
/* FiLM backward: dy = fs*dyf; dfs[n][c] = sum_px dyf*y; dfb[n][c] = sum_px dyf. */
int cdm_film_bwd(const void* dyf, int lddyf, const void* y, int n_img, int px, int C, const float* fs, void* dy,
                 float* dfs, float* dfb, void* stream);

/* nn.GroupNorm + ReLU (+FiLM) backward; per-(image, channel) parameter-gradient partials. */
typedef struct {
  const void* x; const void* dyf; int lddyf;
  int n_img, P, C, groups;
  const float* mean_rstd; const float* gamma; const float* beta;
  const float* film_scale;
  void* dx;
  float* dgamma_nc; float* dbeta_nc; float* dfs; float* dfb;
} cdm_gn_bwd_args;
int cdm_gn_bwd(const cdm_gn_bwd_args* a, void* stream);
int cdm_rows_sum(const float* in, int rows, int C, float* out, void* stream);

/* to_vec forward that also keeps the pre-GELU mean, and its backward (accumulates into dx). */
int cdm_avgpool_gelu_train(const void* src, int n_img, int P, int C, float* pre, void* out, void* stream);
int cdm_avgpool_gelu_bwd(const float* pre, const float* dh, int n_img, int P, int C, void* dx, void* stream);

/* Weight gradient of the K=9 (1->C) and N=1 (C->1) convolutions: out[tap][c] = sum_px s[px+d(tap)]*v[px][c];
 * flip=1 negates d(tap); mean_rstd != NULL applies GroupNorm(8)+ReLU to v on load. workspace fp32 [blocks][9][C]. */
typedef struct {
  const float* s; const void* v;
  int n_img, H, W, C, flip;
  const float* mean_rstd; const float* gamma; const float* beta;
  float* workspace; int workspace_blocks;
  float* out;
} cdm_outer_wgrad_args;
int cdm_outer_wgrad(const cdm_outer_wgrad_args* a, void* stream);

/* EmbedFC backward (recomputes the hidden layer): scratch pre/h/dpre are fp32 [rows][emb]. */
typedef struct {
  const float* in; int rows, din, emb;
  const float* w1; const float* b1; const float* w2;
  const float* dout;
  float* pre; float* h; float* dpre;
  float* dw1; float* db1; float* dw2; float* db2;
} cdm_embed_bwd_args;
int cdm_embed_bwd(const cdm_embed_bwd_args* a, void* stream);

/* F.mse_loss(pred, target) summed (loss_sum[0] = sum of squares) and d pred = 2 (pred-target) * inv_count. */
int cdm_mse_grad(const float* pred, const float* target, long long n, float inv_count, float* dpred, float* partial,
                 int partial_blocks, float* loss_sum, void* stream);
/* torch.optim.Adam defaults over a device table of {p, g, m, v, n} (5 x 8 bytes per tensor).  lr_dev / step_dev
 * (device scalars, optional) override lr / step so that the launch can be replayed from a CUDA graph. */
int cdm_adam_step(const void* table, int n_tensors, long long max_numel, float lr, float beta1, float beta2, float eps,
                  int step, const float* lr_dev, const int* step_dev, void* stream);

/* Weight-gradient GEMM  C[m][tap*tap_stride + n] += sum_px A[px][m_off+m] * B[px + (kh-1,kw-1)][n_off+n]
 * with both operands read straight from NHWC bf16 activations (the reduction index is the pixel, the
 * channels are contiguous: tensor-core "MN-major" operands), fp32 accumulate, split-K with fp32 atomics:
 * the caller zero-fills C.  taps = 9 is the 3x3 convolution wgrad (what autograd / cuDNN computes for
 * nn.Conv2d in loss.backward(), code/train_diffusion_paper.py:363): A = dL/d(conv output), B = conv input,
 * C = dW as [cout][3][3][cin].  taps = 1 with H = n_img = 1, W = rows is a plain A^T B (transposed-conv and
 * up0 wgrad).  M, N multiples of 128; channel windows let A / B be slices of wider tensors. */
typedef struct {
  const void* a; /* bf16 [n_img][H][W][a_c] */
  int a_c;
  const void* b; /* bf16 [n_img][H][W][b_c] */
  int b_c;
  int n_img, H, W;
  int taps;
  int m_off, M;
  int n_off, N;
  float* c;
  int ldc, tap_stride;
  int k_split; /* 0 = choose so that ~2 CTAs per SM have work */
  float* workspace;            /* split-K partial tiles: taps == 9 needs >= 3 * (M/128) * (N/128) * k_split * 128 * 384
                                * floats, taps == 1 >= (M/128) * (N/128) * k_split * 128 * 128.  Given (and large enough):
                                * every K slice writes its tile with plain stores and a second kernel adds the slices in
                                * a fixed order (deterministic, and ~7 M fewer fp32 atomics per 3x3 layer).  NULL: fp32
                                * atomics straight into c (run-to-run differences in the last bits). */
  long long workspace_floats;
  float* probe; /* measurement only (NULL in production): fp32 [148][4] = per-CTA issuer cycles blocked on the TMA
                 * ring, blocked on the accumulator, total, epilogue cycles (taps == 9 path) */
} cdm_gemm_tn_args;
int cdm_gemm_tn(const cdm_gemm_tn_args* a, void* stream);

/* Radial power spectrum of n_maps square N x N fp32 maps (N a power of two <= 64), one CTA per map.
 * Replaces power_spectrum (code/diffusion_utilities.py:302-368), which compare_power_spectra (:370-431)
 * calls per image: np.fft.fftn(box, norm="ortho"), |F|^2, mean over the modes of each radial bin, * dl^ndims.
 * Bin membership is the reference's own rule int(round(|k|/dk)) evaluated once on the host and passed as a
 * CSR list: bin_start int32 [n_bins+1], bin_items int32 [N*N] (flat mode indices l*N+k grouped by bin, ascending
 * within a bin).  pk fp64 [n_maps][n_bins] = scale * mean (0 for an empty bin); scale = dl^2. */
int cdm_power_spectrum(const float* maps, int n_maps, int N, const int* bin_start, const int* bin_items, int n_bins,
                       double scale, double* pk, void* stream);

/* Per-map pixel histogram with explicit fp64 bin edges (ascending, n_bins+1 of them): np.histogram(map.ravel(),
 * edges) of compare_distributions (code/train_diffusion_paper.py:861-876) — half-open bins, last bin closed,
 * values outside [edges[0], edges[n_bins]] dropped.  counts int32 [n_maps][n_bins]; n_bins <= 12288. */
int cdm_pixel_histogram(const float* maps, int n_maps, int P, const double* edges, int n_bins, int* counts,
                        void* stream);

/* ---- data preparation (code/train_diffusion_paper.py:232-262) ------------------------------------------------ */
/* out[0] = min, out[1] = max of n fp32 values (x 16-byte aligned).  workspace: fp32 scratch the caller zero-fills
 * ONCE (>= 17 floats; 2 per block + a ticket counter in the last slot, re-armed by the kernel). */
int cdm_minmax(const float* x, long long n, float* workspace, int workspace_floats, float* out, void* stream);
/* The map pipeline of :254-261 fused with the resize: with (min, max) of the RAW maps in raw_minmax (device),
 *   v -> (min <= 0 ? v - min + 1e-8 : v) / max' -> log10 -> (. - lmin) / (lmax - lmin)
 * evaluated only at the taps of F.interpolate(size=(Ho,Wo), mode='bilinear') (align_corners=False), fp32,
 * every operation rounded separately in the reference's order.  in fp32 [n][Hi][Wi] -> out fp32 [n][Ho][Wo]. */
int cdm_preprocess_maps(const float* in, int n, int Hi, int Wi, const float* raw_minmax, int Ho, int Wo, float* out,
                        void* stream);
/* Parameter table of :232-252: col_min / col_max over the rows of x [rows][cols]; out [rows*repeat][out_cols] =
 * (x - min) / (max - min + 1e-8) with every row repeated `repeat` times (np.repeat(param_data, 15, axis=0)),
 * columns cut (cols > out_cols) or zero-padded (cols < out_cols) to the number of conditioning parameters. */
int cdm_normalize_params(const float* x, int rows, int cols, int repeat, int out_cols, float* out, float* col_min,
                         float* col_max, void* stream);

/* All bf16 operand layouts of a training step in one launch (the weights change every optimizer step): `table` is
 * a device array of n_rows rows of 12 x int64 {src fp32*, dst bf16*, d1, d2, d3, s0, s1, s2, s3, off, vec_start, 0}:
 *   dst[i0][i1][i2][i3] = (bf16) src[off + i0 s0 + i1 s1 + i2 s2 + i3 s3],  d3 % 8 == 0,
 * vec_start = running count of 8-element output vectors, total_vec their total.  Replaces the per-tensor
 * w.permute(..).contiguous().to(bf16) / w.flip(2,3).permute(..) chains of the torch path. */
int cdm_pack_bf16(const void* table, int n_rows, long long total_vec, void* stream);
/* dst[b][c][r] (bf16) = src[b][r][c] (fp32): batched transposing cast through shared memory, both innermost dimensions
 * contiguous (strides in elements; R, Cc multiples of 64).  The two tensor-core layouts of up0.0.weight
 * (ConvTranspose2d IOHW [ci][co][16*16]) are such transposes: [ci][khw][co] (b = ci, r = co, c = khw) and
 * [khw][co][ci] (b = co, r = ci, c = khw). */
int cdm_pack_transpose_bf16(const float* src, void* dst, int batches, int R, int Cc, long long src_batch_stride,
                            long long src_row_stride, long long dst_batch_stride, long long dst_col_stride, void* stream);

/* ======================= composite entry points: one eval forward / one sampling step per C call ===============
 * ContextUnet.forward in eval mode (ContextUnet.py:42-60) and one iteration of sample_ddpm's loop
 * (code/train_diffusion_paper.py:594-618).  A plan packs the fp32 PyTorch-layout parameters into the tensor-core
 * layouts once (bf16 K-major weights, eval BatchNorm folded to scale / shift), carves the activation workspace and
 * encodes every layer's tensor maps; cdm_forward_eval is then 26 kernel launches and nothing else (no descriptor
 * encoding, no allocation, no host sync; CUDA-graph capturable).  The library allocates no device memory: arena
 * and workspace are the caller's.  A plan is used from one host thread on one stream at a time. */
typedef struct cdm_plan cdm_plan;

/* Parameter / buffer tensors in ContextUnet.state_dict() order without the num_batches_tracked entries:
 * 102 parameters + 36 BatchNorm running statistics = 138 fp32 tensors, PyTorch shapes (Conv2d OIHW,
 * ConvTranspose2d IOHW, Linear [out][in]). */
int cdm_plan_n_tensors(void);
const char* cdm_plan_tensor_name(int i);              /* state_dict key, e.g. "down1.model.0.conv1.0.weight" */
long long cdm_plan_tensor_numel(int i, int n_cfeat);  /* element count for a context width of n_cfeat */

#define CDM_CONV_MODE_DEFAULT CDM_CONV_MODE_SWAPPED_TMA
typedef struct {
  int n_cfeat;      /* width of the context vector (1..6 in the reference's sweeps) */
  int batch, reps;  /* batch inputs x, each evaluated `reps` times (2 = the conditional + unconditional passes of
                     * classifier-free guidance, run as ONE 2*batch-image forward); images in flight = batch*reps */
  const float* const* tensors; /* HOST array of cdm_plan_n_tensors() DEVICE pointers; the tensors must outlive the
                                * plan (small fp32 vectors are used in place); after updating them in place call
                                * cdm_plan_refresh */
  void* arena;      /* packed weights, 256-byte aligned, >= cdm_plan_arena_bytes(n_cfeat) */
  long long arena_bytes;
  void* workspace;  /* activations, 256-byte aligned, >= cdm_plan_workspace_bytes(batch, reps) */
  long long workspace_bytes;
  int conv_mode;    /* CDM_CONV_MODE_*; 0 selects CDM_CONV_MODE_DEFAULT */
} cdm_plan_desc;
long long cdm_plan_arena_bytes(int n_cfeat);
long long cdm_plan_workspace_bytes(int batch, int reps);
int cdm_plan_create(const cdm_plan_desc* d, void* stream, cdm_plan** out); /* packing kernels run on `stream` */
int cdm_plan_refresh(cdm_plan* p, void* stream);                            /* re-pack after an in-place update */
void cdm_plan_destroy(cdm_plan* p);
/* Named view into the workspace (tests / taps): "x0", "d1", "d2", "hidden", "u0f", "u1f", "p64", "q64", "eps", ... */
int cdm_plan_buffer(const cdm_plan* p, const char* name, void** ptr, long long* bytes);
/* EmbedFC with the plan's weights: which = 0 contextembed1, 1 timeembed1, 2 contextembed2, 3 timeembed2
 * (ContextUnet.py:51-54).  in fp32 [rows][n_cfeat or 1] -> out fp32 [rows][256 or 128]. */
int cdm_plan_embed(const cdm_plan* p, int which, const float* in, int rows, float* out, void* stream);

typedef struct {
  const float* x;      /* fp32 [batch][64][64] */
  const float* sc_tab; /* fp32 [steps][reps][2][128]: the fresh 1x1 shortcut draws (w_c, b_c) per pass */
  const float* cemb1;  /* fp32 [reps*batch][256]  contextembed1(c) */
  const float* temb1;  /* fp32 [steps][temb_rows][256] timeembed1(t) */
  const float* cemb2;  /* fp32 [reps*batch][128] */
  const float* temb2;  /* fp32 [steps][temb_rows][128] */
  int temb_rows;       /* 1 (t shared by the batch) or reps*batch */
  const int* step_ptr; /* device int32 selecting the sc_tab / temb row, NULL = row 0 */
  float* eps;          /* fp32 [reps*batch][64][64]; NULL = the plan's "eps" buffer */
} cdm_forward_args;
int cdm_forward_eval(cdm_plan* p, const cdm_forward_args* f, void* stream);
/* Measurement: the forward's launches by name, and one forward with a CUDA event after every launch on `stream`
 * (synchronises; ms_host[cdm_plan_n_launches()] = per-launch durations).  bench.py's roofline figures. */
int cdm_plan_n_launches(void);
const char* cdm_plan_launch_name(int i);
int cdm_plan_profile(cdm_plan* p, const cdm_forward_args* f, float* ms_host, void* stream);

typedef struct {
  cdm_forward_args fwd; /* x, step_ptr and eps are taken from the fields below */
  float* x;             /* fp32 [batch][64][64] = x_t, updated in place to x_{t-1} */
  int* step_ptr;        /* device int32 holding i; decremented by the call */
  float guide_w;
  const float* coef;    /* fp32 [timesteps+1][4] (cdm_ddpm_step_args) */
  int timesteps;
  const float* z;       /* host-fed noise or NULL -> in-kernel Philox */
  long long z_iter_stride;
  unsigned long long seed;
  long long sample_offset;
  float* snap;
  const int* snap_slot;
} cdm_sample_step_args;
int cdm_sample_step(cdm_plan* p, const cdm_sample_step_args* s, void* stream);

/* Measurement probe: every CTA streams `tile_bytes` TMA tiles from an
 * L2-resident buffer into a shared-memory ring; returns nothing, caller times it. */
int cdm_probe_tma_l2(const void* buf, int n_rows, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDM_B200_H */
