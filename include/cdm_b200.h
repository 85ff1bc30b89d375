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
#define CDM_EPI_SHORTCUT 2  /* y += w_c[co]*x[n,px] + b_c[co]   (G1 shortcut)  */
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
  void* out;          /* bf16 [n_img][H'][W'][cout] */
  /* CDM_EPI_SHORTCUT (cout must be 128) */
  const float* sc_x;   /* fp32 [sc_nx][H][W]; image n reads sc_x[n % sc_nx] */
  int sc_nx;
  const float* sc_tab; /* fp32 [steps][n_img/sc_nx][2][cout]: w_c then b_c */
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

/* Measurement probe: every CTA streams `tile_bytes` TMA tiles from an
 * L2-resident buffer into a shared-memory ring; returns nothing, caller times it. */
int cdm_probe_tma_l2(const void* buf, int n_rows, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDM_B200_H */
