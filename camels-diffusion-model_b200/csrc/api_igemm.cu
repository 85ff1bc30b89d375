// C-ABI entry points for the tcgen05 kernels (+ library-wide error plumbing).
#include <stdarg.h>
#include <string.h>

#include "common.h"
#include "igemm.cuh"

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

static int num_sms() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

template <int MODE>
static int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvKParams& p,
                       cudaStream_t st) {
  constexpr int smem = conv_smem_bytes<MODE>();
  static bool attr_set = false;
  if (!attr_set) {
    CDM_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  conv3x3_kernel<MODE><<<grid, kConvThreads, smem, st>>>(a0, a1, b, p);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

}  // namespace cdm

using namespace cdm;

extern "C" int cdm_version(void) { return 100; }
extern "C" const char* cdm_last_error(void) { return g_err; }
extern "C" int cdm_device_ok(void) { return check_device(); }

extern "C" int cdm_conv3x3(const cdm_conv3x3_args* a, void* stream) {
  CDM_CHECK_ARG(a != nullptr);
  CDM_CHECK_ARG(a->src0 && a->weight && a->scale && a->shift && a->out);
  CDM_CHECK_ARG(a->c0 > 0 && a->c0 % 64 == 0 && a->c1 >= 0 && a->c1 % 64 == 0);
  CDM_CHECK_ARG((a->c1 == 0) == (a->src1 == nullptr));
  CDM_CHECK_ARG(a->n_img > 0 && a->H >= 16 && a->W >= 16 && a->H % 16 == 0 && a->W % 16 == 0);
  CDM_CHECK_ARG(a->cout > 0 && a->cout % 128 == 0 && a->cout <= 256);
  CDM_CHECK_ARG(a->mode >= 0 && a->mode <= 3);
  if (a->flags & CDM_EPI_SHORTCUT)
    CDM_CHECK_ARG(a->sc_x && a->sc_tab && a->sc_reps >= 1);
  if (a->flags & CDM_EPI_FILM) CDM_CHECK_ARG(a->film_scale && a->film_shift && a->film_shift_rows >= 1);
  if (a->flags & CDM_EPI_GNSTATS) CDM_CHECK_ARG(a->gn_partial && a->cout == 128 && !(a->flags & CDM_EPI_POOL));
  if (a->flags & CDM_EPI_BNSTATS)
    CDM_CHECK_ARG(a->bn_partial && a->bn_sums && a->mode == 3 && a->H % 32 == 0 && a->W % 8 == 0 &&
                  !(a->flags & (CDM_EPI_POOL | CDM_EPI_SHORTCUT)));
  int rc = check_device();
  if (rc) return rc;

  // MODE 3 (swapped operands, 8 px x 32 row patches) needs H % 32 == 0; otherwise fall back to SHIFT18
  const int mode = (a->mode == 3 && (a->H % 32 != 0 || a->W % 8 != 0)) ? 2 : a->mode;
  static const int pitch_of_mode[4] = {16, 24, 18, 10};
  const uint32_t pitch = pitch_of_mode[mode];
  const uint32_t box_rows = mode == 3 ? 34 : 18;
  CUtensorMap mA0, mA1, mB;
  {
    uint64_t dims[4] = {(uint64_t)a->c0, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->c0 * 2, (uint64_t)a->W * a->c0 * 2, (uint64_t)a->H * a->W * a->c0 * 2};
    uint32_t box[4] = {64, pitch, box_rows, 1};
    rc = make_tmap_bf16(&mA0, a->src0, 4, dims, str, box);
    if (rc) return rc;
  }
  if (a->src1) {
    uint64_t dims[4] = {(uint64_t)a->c1, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->c1 * 2, (uint64_t)a->W * a->c1 * 2, (uint64_t)a->H * a->W * a->c1 * 2};
    uint32_t box[4] = {64, pitch, box_rows, 1};
    rc = make_tmap_bf16(&mA1, a->src1, 4, dims, str, box);
    if (rc) return rc;
  } else {
    mA1 = mA0;
  }
  const int cin = a->c0 + a->c1;
  {
    uint64_t dims[2] = {(uint64_t)9 * cin, (uint64_t)a->cout};
    uint64_t str[1] = {(uint64_t)9 * cin * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&mB, a->weight, 2, dims, str, box);
    if (rc) return rc;
  }
  ConvKParams p;
  memset(&p, 0, sizeof(p));
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
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mode == 3) {
    const int units32 = a->n_img * (a->W / 8) * (a->H / 32) * p.n_tiles;
    const int grid_cap = num_sms();
    // small launches (batch-1 sampling): when 8 x 32 patches would leave three quarters of the SMs idle, use 8 x 8
    // patches — four times the units (still one wave), a quarter of the MMA chain each.  Not with the statistics epilogues, whose
    // partial-sum layout (and summation order) is tied to the 32-row patch.
    const bool small = units32 * 4 <= grid_cap && !(a->flags & (CDM_EPI_GNSTATS | CDM_EPI_BNSTATS));  // one wave
    if (small) {
      p.n_units = units32 * 4;
      constexpr int smem = conv_sw_smem_bytes<8>();
      static bool attr_set8 = false;
      if (!attr_set8) {
        CDM_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_sw_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set8 = true;
      }
      // the halo box of the 8-row patch is {64, 10, 10, 1}: its own tensor maps
      CUtensorMap sA0, sA1;
      {
        uint64_t dims[4] = {(uint64_t)a->c0, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
        uint64_t str[3] = {(uint64_t)a->c0 * 2, (uint64_t)a->W * a->c0 * 2, (uint64_t)a->H * a->W * a->c0 * 2};
        uint32_t box[4] = {64, 10, 10, 1};
        rc = make_tmap_bf16(&sA0, a->src0, 4, dims, str, box);
        if (rc) return rc;
      }
      if (a->src1) {
        uint64_t dims[4] = {(uint64_t)a->c1, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
        uint64_t str[3] = {(uint64_t)a->c1 * 2, (uint64_t)a->W * a->c1 * 2, (uint64_t)a->H * a->W * a->c1 * 2};
        uint32_t box[4] = {64, 10, 10, 1};
        rc = make_tmap_bf16(&sA1, a->src1, 4, dims, str, box);
        if (rc) return rc;
      } else {
        sA1 = sA0;
      }
      const int grid = p.n_units < grid_cap ? p.n_units : grid_cap;
      conv3x3_sw_kernel<8><<<grid, kSwThreads, smem, st>>>(sA0, sA1, mB, p);
      CDM_CHECK_LAUNCH();
      return CDM_OK;
    }
    p.n_units = units32;
    constexpr int smem = conv_sw_smem_bytes<32>();
    static bool attr_set = false;
    if (!attr_set) {
      CDM_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_sw_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    const int grid = p.n_units < num_sms() ? p.n_units : num_sms();
    conv3x3_sw_kernel<32><<<grid, kSwThreads, smem, st>>>(mA0, mA1, mB, p);
    CDM_CHECK_LAUNCH();
    if (a->flags & CDM_EPI_BNSTATS) {  // fold the per-CTA rows (and the other ranks' sums) into bn_sums
      launch_xrank_sum(a->bn_partial, grid, 2 * a->cout, a->bn_sums, a->xr, st);
      CDM_CHECK_LAUNCH();
    }
    return CDM_OK;
  }
  switch (mode) {
    case 0: return launch_conv<0>(mA0, mA1, mB, p, st);
    case 1: return launch_conv<1>(mA0, mA1, mB, p, st);
    default: return launch_conv<2>(mA0, mA1, mB, p, st);
  }
}

extern "C" int cdm_gemm(const cdm_gemm_args* a, void* stream) {
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
  CUtensorMap mA0, mA1, mB;
  {
    uint64_t dims[2] = {(uint64_t)a->k0, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->k0 * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&mA0, a->a0, 2, dims, str, box);
    if (rc) return rc;
  }
  if (a->a1) {
    uint64_t dims[2] = {(uint64_t)a->k1, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->k1 * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&mA1, a->a1, 2, dims, str, box);
    if (rc) return rc;
  } else {
    mA1 = mA0;
  }
  const int K = a->k0 + a->k1;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap_bf16(&mB, a->bw, 2, dims, str, box);
    if (rc) return rc;
  }
  // small N (transposed convolutions): keep the weight tiles resident, stream A once per weight group
  const int n_tiles_total = a->N / 128;
  const int n_res = K <= 256 ? 2 : (K <= 512 ? 1 : 0);
  // few_n: the 2x2 transposed convolutions (N = 512): one weight group per CTA.  many_n: up0 at sampling batch
  // sizes (N = 65536): several groups per CTA, swapped in turn; needs the same shift rows for every group.
  const bool few_n = n_res > 0 && n_tiles_total <= 4 && n_tiles_total % n_res == 0 && a->M >= 128 * 64;
  const bool many_n = n_res == 2 && n_tiles_total % 2 == 0 && n_tiles_total / 2 > num_sms() && a->out_mode == 0 &&
                      256 % a->shift_mod == 0 && a->M >= 1024;
  if (few_n || many_n) {
    GemmBresKParams q;
    memset(&q, 0, sizeof(q));
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
    for (int v = a->H; v > 1; v >>= 1) ++q.h_shift;
    for (int v = a->W; v > 1; v >>= 1) ++q.w_shift;
    q.out = reinterpret_cast<bf16*>(a->out);
    constexpr int smem_b = gemm_bres_smem_bytes();
    static bool attr_b = false;
    if (!attr_b) {
      CDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_bres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
      attr_b = true;
    }
    const int groups_per_cta = (q.n_groups + num_sms() - 1) / num_sms();
    const int grid_b = many_n ? (q.n_groups + groups_per_cta - 1) / groups_per_cta : (num_sms() / q.n_groups) * q.n_groups;
    gemm_bres_kernel<<<grid_b, kBresThreads, smem_b, reinterpret_cast<cudaStream_t>(stream)>>>(mA0, mA1, mB, q);
    CDM_CHECK_LAUNCH();
    return CDM_OK;
  }
  GemmKParams p;
  memset(&p, 0, sizeof(p));
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
  for (int v = a->H; v > 1; v >>= 1) ++p.h_shift;
  for (int v = a->W; v > 1; v >>= 1) ++p.w_shift;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.k_split = 1;
  // few output tiles but a long K (the up0 data gradient: 2 tiles, K = 65536): split K over the idle SMs
  if (a->workspace && a->out_mode == 0 && p.n_units * 4 <= num_sms() && p.chunks >= 64) {
    int ks = num_sms() / p.n_units;
    if (ks > p.chunks / 8) ks = p.chunks / 8;
    const long long need = (long long)ks * p.m_tiles * 128 * a->N;
    if (ks > 1 && a->workspace_floats >= need) {
      p.k_split = ks;
      p.partial = a->workspace;
      p.n_units *= ks;
    }
  }
  constexpr int smem = gemm_smem_bytes();
  static bool attr_set = false;
  if (!attr_set) {
    CDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  gemm_kernel<<<grid, kConvThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(mA0, mA1, mB, p);
  CDM_CHECK_LAUNCH();
  if (p.partial) {
    const long long quads = ((long long)a->M * a->N + 3) / 4;
    gemm_splitk_reduce_kernel<<<(int)((quads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p.partial, p.k_split, a->M, p.m_tiles * 128, a->N, a->shift, a->shift_mod, p.out);
    CDM_CHECK_LAUNCH();
  }
  return CDM_OK;
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
    static bool attr9_set = false;
    if (!attr9_set) {
      CDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem9));
      attr9_set = true;
    }
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
  static bool attr_set = false;
  if (!attr_set) {
    CDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
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
  static bool attr_set = false;
  if (!attr_set) {
    CDM_CHECK_CUDA(cudaFuncSetAttribute(probe_tma_l2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  probe_tma_l2_kernel<<<num_sms(), 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(m, n_rows, iters);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
