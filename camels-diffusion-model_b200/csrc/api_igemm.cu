// C-ABI entry points for the tcgen05 kernels (+ library-wide error plumbing).
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include <utility>

#include "common.h"
#include "igemm.cuh"
#include "launch.h"

namespace cdm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s (no CUDA device? this library has no CPU path)", cudaGetErrorString(e));
    return CDM_ERR_CUDA;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    return CDM_ERR_CUDA;
  }
  if (major != 10) {
    set_error("device compute capability %d.x is not sm_100: kernels are built for sm_100a only", major);
    return CDM_ERR_ARCH;
  }
  return CDM_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is reached through the runtime so that the .so loads on a box without a driver.
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CDM_ERR_CUDA;
  cuuint64_t gd[5];
  cuuint64_t gs[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gs[i] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return CDM_ERR_CUDA;
  }
  return CDM_OK;
}

int num_sms() {
  static int cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!cache[dev]) cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev);
  return cache[dev];
}

}  // namespace cdm

using namespace cdm;

extern "C" int cdm_version(void) { return 200; }
extern "C" const char* cdm_last_error(void) { return g_err; }
extern "C" int cdm_device_ok(void) { return check_device(); }
extern "C" int cdm_num_sms(void) { return check_device() == CDM_OK ? num_sms() : 0; }

// ---------------------------------------------------------------------------------------------------------------
// cdm_conv3x3 / cdm_gemm are split into a PREPARE step (argument checks, tensor maps, kernel choice: everything that
// depends only on shapes and pointers) and a LAUNCH step, so that a cdm_plan (plan.cu) encodes each layer's tensor
// maps once and a forward is nothing but kernel launches.
// ---------------------------------------------------------------------------------------------------------------
namespace cdm {

static int g_attr_done[64][16];  // [device][kernel variant]: cudaFuncSetAttribute is per device, not per process

template <typename K>
static int set_smem_attr(K kernel, int variant, int bytes) {
  int dev = 0;
  CDM_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 63;
  if (dev == 63 || !g_attr_done[dev][variant]) {
    CDM_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    g_attr_done[dev][variant] = 1;
  }
  return CDM_OK;
}

// Launch with programmatic stream serialization (PDL): the kernel may become resident while its stream predecessor
// is still running; conv3x3_sw_kernel orders itself with griddepcontrol.wait.  CDM_PDL=0 disables it (A/B runs).
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDM_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)block, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static int act_tmap(CUtensorMap* m, const void* base, int c, int W, int H, int n, uint32_t bw, uint32_t bh) {
  uint64_t dims[4] = {(uint64_t)c, (uint64_t)W, (uint64_t)H, (uint64_t)n};
  uint64_t str[3] = {(uint64_t)c * 2, (uint64_t)W * c * 2, (uint64_t)H * W * c * 2};
  uint32_t box[4] = {64, bw, bh, 1};
  return make_tmap_bf16(m, base, 4, dims, str, box);
}

int conv_prepare(const cdm_conv3x3_args* a, ConvLaunch* L) {
  CDM_CHECK_ARG(a != nullptr);
  CDM_CHECK_ARG(a->src0 && a->weight && a->scale && a->shift && a->out);
  CDM_CHECK_ARG(a->c0 > 0 && a->c0 % 64 == 0 && a->c1 >= 0 && a->c1 % 64 == 0);
  CDM_CHECK_ARG((a->c1 == 0) == (a->src1 == nullptr));
  CDM_CHECK_ARG(a->n_img > 0 && a->H >= 16 && a->W >= 16 && a->H % 16 == 0 && a->W % 16 == 0);
  CDM_CHECK_ARG(a->cout > 0 && a->cout % 128 == 0 && a->cout <= 256);
  CDM_CHECK_ARG(a->mode >= 0 && a->mode <= 4);
#ifdef CDM_PROBES
  CDM_CHECK_ARG((a->flags & ~(CDM_EPI_ALL | (127 << 24))) == 0);
#else
  CDM_CHECK_ARG((a->flags & ~CDM_EPI_ALL) == 0);  // unknown bits are an error, never silently forwarded to the kernel
#endif
  if (a->flags & CDM_EPI_SHORTCUT)
    CDM_CHECK_ARG(a->sc_x && a->sc_tab && a->sc_reps >= 1);
  if (a->flags & CDM_EPI_FILM) CDM_CHECK_ARG(a->film_scale && a->film_shift && a->film_shift_rows >= 1);
  if (a->flags & CDM_EPI_GNSTATS) CDM_CHECK_ARG(a->gn_partial && a->cout == 128 && !(a->flags & CDM_EPI_POOL));
  if (a->flags & CDM_EPI_BNSTATS)
    CDM_CHECK_ARG(a->bn_partial && a->bn_sums && a->mode >= 3 && a->H % 32 == 0 && a->W % 8 == 0 &&
                  !(a->flags & (CDM_EPI_POOL | CDM_EPI_SHORTCUT)));
  if (a->flags & CDM_EPI_BNBWD)
    CDM_CHECK_ARG(a->bn_partial && a->bn_sums && a->mode >= 3 && a->H % 32 == 0 && a->W % 8 == 0 && a->bwd_z &&
                  a->bwd_scale && a->bwd_shift && a->bwd_mean && a->bwd_rstd &&
                  !(a->flags & (CDM_EPI_POOL | CDM_EPI_SHORTCUT | CDM_EPI_BNSTATS | CDM_EPI_GNSTATS)));
  int rc = check_device();
  if (rc) return rc;
  memset(L, 0, sizeof(*L));

  // MODE 3 / 4 (swapped operands, 8 px x 32 row patches) need H % 32 == 0; otherwise fall back to SHIFT18
  const bool swapped = a->mode >= 3 && a->H % 32 == 0 && a->W % 8 == 0;
  const int mode = swapped ? a->mode : (a->mode >= 3 ? 2 : a->mode);
  const int cin = a->c0 + a->c1;
  {
    uint64_t dims[2] = {(uint64_t)9 * cin, (uint64_t)a->cout};
    uint64_t str[1] = {(uint64_t)9 * cin * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&L->b, a->weight, 2, dims, str, box);
    if (rc) return rc;
  }
  ConvKParams& p = L->p;
  p.H = a->H;
  p.W = a->W;
  p.n_img = a->n_img;
  p.chunks0 = a->c0 / 64;
  p.chunks = cin / 64;
  p.n_tiles = a->cout / 128;
  p.cout = a->cout;
  p.strips_x = a->W / 16;
  p.strips_y = a->H / 16;
  p.n_units = a->n_img * p.strips_x * p.strips_y * p.n_tiles;
  p.flags = a->flags;
  p.scale = a->scale;
  p.shift = a->shift;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.sc_x = a->sc_x;
  p.sc_tab = a->sc_tab;
  p.sc_reps = a->sc_reps > 0 ? a->sc_reps : 1;
  p.film_scale = a->film_scale;
  p.film_shift = a->film_shift;
  p.film_shift_rows = a->film_shift_rows;
  p.step_ptr = a->step_ptr;
  p.gn_partial = a->gn_partial;
  p.bn_partial = a->bn_partial;
  p.res_scale = (a->flags & CDM_EPI_RESSCALE) ? a->res_scale : 1.f;
  p.bwd_z = reinterpret_cast<const bf16*>(a->bwd_z);
  p.bwd_scale = a->bwd_scale, p.bwd_shift = a->bwd_shift, p.bwd_mean = a->bwd_mean, p.bwd_rstd = a->bwd_rstd;
  const int sms = num_sms();
  uint32_t pitch = 18, box_rows = 18;
  if (swapped) {
    const int units32 = a->n_img * (a->W / 8) * (a->H / 32) * p.n_tiles;
    // small launches (batch-1 sampling): when 8 x 32 patches would leave three quarters of the SMs idle, use 8 x 8
    // patches — four times the units (still one wave), a quarter of the MMA chain each.  Not with the statistics
    // epilogues, whose partial-sum layout (and summation order) is tied to the 32-row patch.
    const bool small = units32 * 4 <= sms && !(a->flags & (CDM_EPI_GNSTATS | CDM_EPI_BNSTATS | CDM_EPI_BNBWD));
    const bool tma = mode == 4;
    L->variant = (small ? kConvSw8 : kConvSw32) + (tma ? 2 : 0);
    p.n_units = small ? units32 * 4 : units32;
    pitch = 10;
    box_rows = small ? 10 : 34;
    if (a->flags & (CDM_EPI_BNSTATS | CDM_EPI_BNBWD)) {
      L->bn_fold = 1;
      L->bn_sums = a->bn_sums;
      L->xr = a->xr;
    }
    // output map of the TMA-store epilogue: box {64 ch, 8 px, 4 rows}; all images the epilogue may fan out to
    const int reps = (a->flags & CDM_EPI_SHORTCUT) ? p.sc_reps : 1;
    const int sh = (a->flags & CDM_EPI_POOL) ? 1 : 0;
    rc = act_tmap(&L->out, a->out, a->cout, a->W >> sh, a->H >> sh, a->n_img * reps, 8, 4);
    if (rc) return rc;
  } else {
    static const int pitch_of_mode[3] = {16, 24, 18};
    pitch = pitch_of_mode[mode];
    L->variant = mode;
  }
  rc = act_tmap(&L->a0, a->src0, a->c0, a->W, a->H, a->n_img, pitch, box_rows);
  if (rc) return rc;
  if (a->src1) {
    rc = act_tmap(&L->a1, a->src1, a->c1, a->W, a->H, a->n_img, pitch, box_rows);
    if (rc) return rc;
  } else {
    L->a1 = L->a0;
  }
  if (!swapped) L->out = L->a0;
  L->grid = p.n_units < sms ? p.n_units : sms;
  return CDM_OK;
}

int conv_launch(const ConvLaunch& L, cudaStream_t st) {
  int rc = CDM_OK;
  switch (L.variant) {
    case 0: {
      constexpr int smem = conv_smem_bytes<0>();
      if ((rc = set_smem_attr(conv3x3_kernel<0>, 0, smem))) return rc;
      conv3x3_kernel<0><<<L.grid, kConvThreads, smem, st>>>(L.a0, L.a1, L.b, L.p);
      break;
    }
    case 1: {
      constexpr int smem = conv_smem_bytes<1>();
      if ((rc = set_smem_attr(conv3x3_kernel<1>, 1, smem))) return rc;
      conv3x3_kernel<1><<<L.grid, kConvThreads, smem, st>>>(L.a0, L.a1, L.b, L.p);
      break;
    }
    case 2: {
      constexpr int smem = conv_smem_bytes<2>();
      if ((rc = set_smem_attr(conv3x3_kernel<2>, 2, smem))) return rc;
      conv3x3_kernel<2><<<L.grid, kConvThreads, smem, st>>>(L.a0, L.a1, L.b, L.p);
      break;
    }
    case kConvSw32: {
      constexpr int smem = conv_sw_smem_bytes<32, false>();
      if ((rc = set_smem_attr(conv3x3_sw_kernel<32, false>, kConvSw32, smem))) return rc;
      CDM_CHECK_CUDA(launch_pdl(conv3x3_sw_kernel<32, false>, L.grid, kSwThreads, smem, st, L.a0, L.a1, L.b, L.out, L.p));
      break;
    }
    case kConvSw8: {
      constexpr int smem = conv_sw_smem_bytes<8, false>();
      if ((rc = set_smem_attr(conv3x3_sw_kernel<8, false>, kConvSw8, smem))) return rc;
      CDM_CHECK_CUDA(launch_pdl(conv3x3_sw_kernel<8, false>, L.grid, kSwThreads, smem, st, L.a0, L.a1, L.b, L.out, L.p));
      break;
    }
    case kConvSw32 + 2: {
      constexpr int smem = conv_sw_smem_bytes<32, true>();
      if ((rc = set_smem_attr(conv3x3_sw_kernel<32, true>, kConvSw32 + 2, smem))) return rc;
      CDM_CHECK_CUDA(launch_pdl(conv3x3_sw_kernel<32, true>, L.grid, kSwThreads, smem, st, L.a0, L.a1, L.b, L.out, L.p));
      break;
    }
    case kConvSw8 + 2: {
      constexpr int smem = conv_sw_smem_bytes<8, true>();
      if ((rc = set_smem_attr(conv3x3_sw_kernel<8, true>, kConvSw8 + 2, smem))) return rc;
      CDM_CHECK_CUDA(launch_pdl(conv3x3_sw_kernel<8, true>, L.grid, kSwThreads, smem, st, L.a0, L.a1, L.b, L.out, L.p));
      break;
    }
    default:
      set_error("conv_launch: bad variant %d", L.variant);
      return CDM_ERR_ARG;
  }
  CDM_CHECK_LAUNCH();
  if (L.bn_fold) {  // fold the per-CTA rows (and the other ranks' sums) into bn_sums
    launch_xrank_sum(L.p.bn_partial, L.grid, 2 * L.p.cout, L.bn_sums, L.xr, st);
    CDM_CHECK_LAUNCH();
  }
  return CDM_OK;
}

int gemm_prepare(const cdm_gemm_args* a, GemmLaunch* G) {
  CDM_CHECK_ARG(a != nullptr);
  CDM_CHECK_ARG(a->a0 && a->bw && a->shift && a->out);
  CDM_CHECK_ARG(a->k0 > 0 && a->k0 % 64 == 0 && a->k1 >= 0 && a->k1 % 64 == 0);
  CDM_CHECK_ARG((a->k1 == 0) == (a->a1 == nullptr));
  CDM_CHECK_ARG(a->M > 0 && a->N > 0 && a->N % 128 == 0 && a->shift_mod > 0 && a->shift_mod % 128 == 0);
  if (a->out_mode == 1) CDM_CHECK_ARG((a->H & (a->H - 1)) == 0 && (a->W & (a->W - 1)) == 0);
  CDM_CHECK_ARG(a->out_mode == 0 || (a->out_mode == 1 && a->N == 512 && a->H > 0 && a->W > 0 &&
                                     a->M % (a->H * a->W) == 0));
  int rc = check_device();
  if (rc) return rc;
  memset(G, 0, sizeof(*G));
  {
    uint64_t dims[2] = {(uint64_t)a->k0, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->k0 * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&G->a0, a->a0, 2, dims, str, box);
    if (rc) return rc;
  }
  if (a->a1) {
    uint64_t dims[2] = {(uint64_t)a->k1, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->k1 * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&G->a1, a->a1, 2, dims, str, box);
    if (rc) return rc;
  } else {
    G->a1 = G->a0;
  }
  const int K = a->k0 + a->k1;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&G->b, a->bw, 2, dims, str, box);
    if (rc) return rc;
  }
  const int sms = num_sms();
  // small N (transposed convolutions): keep the weight tiles resident, stream A once per weight group
  const int n_tiles_total = a->N / 128;
  const int n_res = K <= 256 ? 2 : (K <= 512 ? 1 : 0);
  // few_n: the 2x2 transposed convolutions (N = 512): one weight group per CTA.  many_n: up0 at sampling batch
  // sizes (N = 65536): several groups per CTA, swapped in turn; needs the same shift rows for every group.
  const bool few_n = n_res > 0 && n_tiles_total <= 4 && n_tiles_total % n_res == 0 && a->M >= 128 * 64;
  const bool many_n = n_res == 2 && n_tiles_total % 2 == 0 && n_tiles_total / 2 > sms && a->out_mode == 0 &&
                      256 % a->shift_mod == 0 && a->M >= 1024;
  int h_shift = 0, w_shift = 0;
  for (int v = a->H; v > 1; v >>= 1) ++h_shift;
  for (int v = a->W; v > 1; v >>= 1) ++w_shift;
  if (few_n || many_n) {
    GemmBresKParams& q = G->q;
    q.M = a->M;
    q.N = a->N;
    q.chunks0 = a->k0 / 64;
    q.chunks = K / 64;
    q.m_tiles = (a->M + 127) / 128;
    q.n_res = n_res;
    q.n_groups = n_tiles_total / n_res;
    q.shift = a->shift;
    q.shift_mod = a->shift_mod;
    q.out_mode = a->out_mode;
    q.H = a->H;
    q.W = a->W;
    q.h_shift = h_shift;
    q.w_shift = w_shift;
    q.out = reinterpret_cast<bf16*>(a->out);
    const int groups_per_cta = (q.n_groups + sms - 1) / sms;
    G->grid = many_n ? (q.n_groups + groups_per_cta - 1) / groups_per_cta : (sms / q.n_groups) * q.n_groups;
    G->variant = 1;
    G->out = G->a0;
    if (a->out_mode == 1) {
      // out[img][2h + kh][2w + kw][128] = five dimensions {c, kw, w, kh, img*H + h}; an epilogue warp stores its 32
      // consecutive rows m (pixels) x 64 channels as one box
      const int W = a->W, bw = W >= 32 ? 32 : W;
      CDM_CHECK_ARG(32 % bw == 0 && a->M % bw == 0);
      uint64_t dims[5] = {128, 2, (uint64_t)W, 2, (uint64_t)(a->M / W)};
      uint64_t str[4] = {256, 512, (uint64_t)512 * W, (uint64_t)1024 * W};
      uint32_t box[5] = {64, 1, (uint32_t)bw, 1, (uint32_t)(32 / bw)};
      rc = make_tmap_bf16(&G->out, a->out, 5, dims, str, box);
      if (rc) return rc;
    }
    return CDM_OK;
  }
  GemmKParams& p = G->p;
  p.M = a->M;
  p.N = a->N;
  p.chunks0 = a->k0 / 64;
  p.chunks = K / 64;
  p.m_tiles = (a->M + 127) / 128;
  p.n_tiles = a->N / 128;
  p.n_units = p.m_tiles * p.n_tiles;
  p.shift = a->shift;
  p.shift_mod = a->shift_mod;
  p.out_mode = a->out_mode;
  p.H = a->H;
  p.W = a->W;
  p.h_shift = h_shift;
  p.w_shift = w_shift;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.k_split = 1;
  // few output tiles but a long K (the up0 data gradient: 2 tiles, K = 65536): split K over the idle SMs
  if (a->workspace && a->out_mode == 0 && p.n_units * 4 <= sms && p.chunks >= 64) {
    int ks = sms / p.n_units;
    if (ks > p.chunks / 8) ks = p.chunks / 8;
    const long long need = (long long)ks * p.m_tiles * 128 * a->N;
    if (ks > 1 && a->workspace_floats >= need) {
      p.k_split = ks;
      p.partial = a->workspace;
      p.n_units *= ks;
    }
  }
  G->grid = p.n_units < sms ? p.n_units : sms;
  G->variant = 0;
  G->M = a->M;
  return CDM_OK;
}

int gemm_launch(const GemmLaunch& G, cudaStream_t st) {
  int rc = CDM_OK;
  if (G.variant == 1) {
    constexpr int smem_b = gemm_bres_smem_bytes();
    if ((rc = set_smem_attr(gemm_bres_kernel, 8, smem_b))) return rc;
    gemm_bres_kernel<<<G.grid, kBresThreads, smem_b, st>>>(G.a0, G.a1, G.b, G.out, G.q);
    CDM_CHECK_LAUNCH();
    return CDM_OK;
  }
  constexpr int smem = gemm_smem_bytes();
  if ((rc = set_smem_attr(gemm_kernel, 9, smem))) return rc;
  gemm_kernel<<<G.grid, kConvThreads, smem, st>>>(G.a0, G.a1, G.b, G.p);
  CDM_CHECK_LAUNCH();
  if (G.p.partial) {
    const long long quads = ((long long)G.M * G.p.N + 3) / 4;
    gemm_splitk_reduce_kernel<<<(int)((quads + 255) / 256), 256, 0, st>>>(G.p.partial, G.p.k_split, G.M,
                                                                          G.p.m_tiles * 128, G.p.N, G.p.shift,
                                                                          G.p.shift_mod, G.p.out);
    CDM_CHECK_LAUNCH();
  }
  return CDM_OK;
}

}  // namespace cdm

extern "C" int cdm_conv3x3(const cdm_conv3x3_args* a, void* stream) {
  ConvLaunch L;
  int rc = conv_prepare(a, &L);
  if (rc) return rc;
  return conv_launch(L, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int cdm_gemm(const cdm_gemm_args* a, void* stream) {
  GemmLaunch G;
  int rc = gemm_prepare(a, &G);
  if (rc) return rc;
  return gemm_launch(G, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int cdm_gemm_tn(const cdm_gemm_tn_args* a, void* stream) {
  CDM_CHECK_ARG(a != nullptr && a->a && a->b && a->c);
  CDM_CHECK_ARG(a->n_img > 0 && a->H > 0 && a->W > 0 && (a->taps == 1 || a->taps == 9));
  CDM_CHECK_ARG(a->a_c % 64 == 0 && a->b_c % 64 == 0 && a->M > 0 && a->N > 0 && a->M % 128 == 0 && a->N % 128 == 0);
  CDM_CHECK_ARG(a->m_off >= 0 && a->m_off % 64 == 0 && a->m_off + a->M <= a->a_c);
  CDM_CHECK_ARG(a->n_off >= 0 && a->n_off % 64 == 0 && a->n_off + a->N <= a->b_c);
  CDM_CHECK_ARG(a->ldc > 0);
  int rc = check_device();
  if (rc) return rc;
  if (a->taps == 9 && a->W % 8 == 0 && a->H % 16 == 0) {
    // 3x3 weight gradient: 8 px x 16 row patches, one kernel row (three taps) per instruction (gemm_tn9_kernel)
    CUtensorMap mA, mB;
    {
      uint64_t dims[4] = {(uint64_t)a->a_c, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
      uint64_t str[3] = {(uint64_t)a->a_c * 2, (uint64_t)a->W * a->a_c * 2, (uint64_t)a->H * a->W * a->a_c * 2};
      uint32_t box[4] = {64, 8, 16, 1};
      rc = make_tmap_bf16(&mA, a->a, 4, dims, str, box);
      if (rc) return rc;
    }
    {
      uint64_t dims[4] = {(uint64_t)a->b_c, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
      uint64_t str[3] = {(uint64_t)a->b_c * 2, (uint64_t)a->W * a->b_c * 2, (uint64_t)a->H * a->W * a->b_c * 2};
      uint32_t box[4] = {64, 10, 16, 1};
      rc = make_tmap_bf16(&mB, a->b, 4, dims, str, box);
      if (rc) return rc;
    }
    GemmTnKParams p;
    memset(&p, 0, sizeof(p));
    p.m_tiles = a->M / 128;
    p.n_tiles = a->N / 128;
    p.taps = 9;
    p.bw = 8;
    p.bh = 16;
    p.tiles_x = a->W / 8;
    p.tiles_y = a->H / 16;
    p.k_blocks = p.tiles_x * p.tiles_y * a->n_img;
    const int base_units = 3 * p.m_tiles * p.n_tiles;
    // one unit per CTA and no second wave: 150 units on 148 SMs would double the makespan
    int ks = a->k_split > 0 ? a->k_split : num_sms() / base_units;
    if (ks > p.k_blocks) ks = p.k_blocks;
    if (ks < 1) ks = 1;
    p.k_split = ks;
    p.n_units = base_units * ks;
    p.a_off = a->m_off;
    p.b_off = a->n_off;
    p.C = a->c;
    p.ldc = a->ldc;
    p.tap_stride = a->tap_stride;
    p.probe = a->probe;
    if (a->workspace) {
      CDM_CHECK_ARG(a->workspace_floats >= (long long)p.n_units * 128 * 384);
      p.partial = a->workspace;
    }
    constexpr int smem9 = gemm_tn9_smem_bytes();
    if ((rc = set_smem_attr(gemm_tn9_kernel, 10, smem9))) return rc;
    const int grid9 = p.n_units < num_sms() ? p.n_units : num_sms();
    gemm_tn9_kernel<<<grid9, kConvThreads, smem9, reinterpret_cast<cudaStream_t>(stream)>>>(mA, mB, p);
    CDM_CHECK_LAUNCH();
    if (p.partial) {
      const long long per_slice = (long long)base_units * 128 * 384;
      gemm_tn9_reduce_kernel<<<(int)((per_slice + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
          p.partial, p.k_split, p.m_tiles, p.n_tiles, p.C, p.ldc, p.tap_stride);
      CDM_CHECK_LAUNCH();
    }
    return CDM_OK;
  }
  // a K block is 128 rows: [rows] mode (H == 1, n_img == 1) takes 128 consecutive rows (ragged tail zero-filled
  // by TMA); image mode takes a bw x bh pixel rectangle so that a tap shift is a coordinate offset
  int bw = 128, bh = 1;
  if (a->H > 1 || a->n_img > 1) {
    bw = a->W;
    CDM_CHECK_ARG(bw <= 128 && 128 % bw == 0);
    bh = 128 / bw;
    CDM_CHECK_ARG(a->H % bh == 0);
  } else {
    CDM_CHECK_ARG(a->taps == 1);
  }
  CUtensorMap mA, mB;
  {
    uint64_t dims[4] = {(uint64_t)a->a_c, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->a_c * 2, (uint64_t)a->W * a->a_c * 2, (uint64_t)a->H * a->W * a->a_c * 2};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
    rc = make_tmap_bf16(&mA, a->a, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)a->b_c, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->b_c * 2, (uint64_t)a->W * a->b_c * 2, (uint64_t)a->H * a->W * a->b_c * 2};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
    rc = make_tmap_bf16(&mB, a->b, 4, dims, str, box);
    if (rc) return rc;
  }
  GemmTnKParams p;
  memset(&p, 0, sizeof(p));
  p.m_tiles = a->M / 128;
  p.n_tiles = a->N / 128;
  p.taps = a->taps;
  p.bw = bw;
  p.bh = bh;
  p.tiles_x = (a->W + bw - 1) / bw;
  p.tiles_y = (a->H + bh - 1) / bh;
  p.k_blocks = p.tiles_x * p.tiles_y * a->n_img;
  const int base_units = p.taps * p.m_tiles * p.n_tiles;
  int ks = a->k_split > 0 ? a->k_split : (2 * num_sms()) / base_units;  // at most two full waves of units
  if (ks > p.k_blocks) ks = p.k_blocks;
  if (ks < 1) ks = 1;
  p.k_split = ks;
  p.n_units = base_units * ks;
  p.a_off = a->m_off;
  p.b_off = a->n_off;
  p.C = a->c;
  p.ldc = a->ldc;
  p.tap_stride = a->tap_stride;
  if (a->workspace && ks > 1 && a->workspace_floats >= (long long)p.n_units * 128 * 128) p.partial = a->workspace;
  constexpr int smem = gemm_tn_smem_bytes();
  if ((rc = set_smem_attr(gemm_tn_kernel, 11, smem))) return rc;
  const int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  gemm_tn_kernel<<<grid, kConvThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(mA, mB, p);
  CDM_CHECK_LAUNCH();
  if (p.partial) {
    const long long per_slice = (long long)base_units * 128 * 128;
    gemm_tn_reduce_kernel<<<(int)((per_slice + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p.partial, p.k_split, p.taps, p.m_tiles, p.n_tiles, p.C, p.ldc, p.tap_stride);
    CDM_CHECK_LAUNCH();
  }
  return CDM_OK;
}

extern "C" int cdm_probe_tma_l2(const void* buf, int n_rows, int iters, void* stream) {
  CDM_CHECK_ARG(buf && n_rows >= 128 && n_rows % 128 == 0 && iters > 0);
  int rc = check_device();
  if (rc) return rc;
  CUtensorMap m;
  uint64_t dims[2] = {64, (uint64_t)n_rows};
  uint64_t str[1] = {128};
  uint32_t box[2] = {64, 128};
  rc = make_tmap_bf16(&m, buf, 2, dims, str, box);
  if (rc) return rc;
  const int smem = 8 * 16384 + 256 + 1024;
  if ((rc = set_smem_attr(probe_tma_l2_kernel, 12, smem))) return rc;
  probe_tma_l2_kernel<<<num_sms(), 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(m, n_rows, iters);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
