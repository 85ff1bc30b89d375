// tcgen05 implicit-GEMM kernels of the ContextUnet hot path (sm_100a only).
//
//   conv3x3_kernel : 3x3/s1/p1 convolution, NHWC bf16, K-split over two sources
//                    (replaces nn.Conv2d + eval BatchNorm2d + ReLU [+ MaxPool2d,
//                    FiLM, shortcut, GroupNorm statistics] — see cdm_b200.h)
//   gemm_kernel    : plain K-major GEMM (ConvTranspose2d 2x2/s2 and the 16x16/s16
//                    up0 transposed conv are GEMMs + a scatter epilogue)
//
// Shared structure (one CTA per SM, persistent over work units, 8 warps):
//   warp 0  lane 0 : TMA producer of the activation (A) operand
//   warp 3  lane 0 : TMA producer of the weight (B) operand
//   warp 1  lane 0 : tcgen05.mma issuer (accumulators in TMEM, double buffered)
//   warp 2         : TMEM allocator
//   warps 4..7     : epilogue (tcgen05.ld -> fp32 math -> bf16 -> smem transpose -> coalesced stores)
#pragma once
#include "../../include/cdm_b200.h"
#include "kparams.h"
#include "ptx.cuh"

namespace cdm {


// Measurement probes (tools/gpu_probe.py) exist only in the -DCDM_PROBES build (libcdm_b200_probes.so); in the
// production library the predicates are compile-time false and the public ABI rejects the probe flag bits.
#ifdef CDM_PROBES
#define CDM_PROBE_BIT(flags, bit) (((flags) >> (bit)) & 1)
#else
#define CDM_PROBE_BIT(flags, bit) false
#endif

// --------------------------------------------------------------------------
// conv3x3: geometry
//   M-tile (one 128-row UMMA) = a patch of 8 pixels (along W) x 16 rows; TMEM
//   lane m = row*8 + px.  The A descriptor walks the 16 rows with its stride
//   byte offset (SBO = smem pitch * 128 B), so a patch is a strided window of
//   one resident halo tile and every tap (kh,kw) is just a different start
//   address: + kh*pitch rows, + kw rows.
//   CTA work unit = strip of 16x16 pixels = 2 patches (2 accumulators), one
//   128-wide slice of Cout.
// --------------------------------------------------------------------------

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

template <int MODE>
struct ConvCfg;
template <>
struct ConvCfg<0> {  // three kw-shifted copies; descriptor starts stay 1024B-atom aligned
  static constexpr int PITCH = 16, A_PER_CHUNK = 3, TAPS_PER_A = 3, NA = 3;
  static constexpr int A_BYTES = 18 * PITCH * 128, A_STRIDE = 36864;
};
template <>
struct ConvCfg<1> {  // one halo tile, pitch 24 (kh shift stays atom aligned, kw shift is +128 B)
  static constexpr int PITCH = 24, A_PER_CHUNK = 1, TAPS_PER_A = 9, NA = 2;
  static constexpr int A_BYTES = 18 * PITCH * 128, A_STRIDE = 55296;
};
template <>
struct ConvCfg<2> {  // one halo tile, pitch 18 (no padding columns; SBO = 2304 B)
  static constexpr int PITCH = 18, A_PER_CHUNK = 1, TAPS_PER_A = 9, NA = 2;
  static constexpr int A_BYTES = 18 * PITCH * 128, A_STRIDE = 41984;
};

constexpr int kNB = 4;               // weight stages
constexpr int kBBytes = 128 * 128;   // 128 Cout rows x 64 ch x 2 B
constexpr int kStageBytes = 4 * 8192;  // epilogue transpose buffers, one per epilogue warp
constexpr int kConvThreads = 256;

template <int MODE>
constexpr int conv_smem_bytes() {
  return ConvCfg<MODE>::NA * ConvCfg<MODE>::A_STRIDE + kNB * kBBytes + kStageBytes + 2 * 256 * 4 + 256 + 1024;
}

// Epilogue of one 128x128 accumulator tile held by this warp's 32 TMEM lanes.
// Converts to bf16 into the warp-private staging buffer (XOR-swizzled 16 B units,
// conflict free for both the row-per-thread writes and the row-contiguous reads).
struct EpiCtx {
  const float* s_scale;  // smem, this n-tile's 128 entries
  const float* s_shift;
  int flags;
  // shortcut
  float xpix;
  const float* sc_w;
  const float* sc_b;
  // film
  const float* film_scale;
  const float* film_shift;
  float res_scale;
};

__device__ __forceinline__ void epi_tile_to_staging(uint32_t taddr, uint32_t stg, int lane, const EpiCtx& e,
                                                    float* gn_sum, float* gn_sq) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t v[32];
    tmem_ld_x32(taddr + cc * 32, v);
    tmem_wait_ld();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = cc * 32 + i;
      float y = fmaf(__uint_as_float(v[i]), e.s_scale[c], e.s_shift[c]);
      if (e.flags & CDM_EPI_RELU) y = fmaxf(y, 0.f);
      if (e.flags & CDM_EPI_GELU) y = gelu_erf(y);
      f[i] = y;
    }
    if (e.flags & CDM_EPI_SHORTCUT) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] += fmaf(__ldg(e.sc_w + cc * 32 + i), e.xpix, __ldg(e.sc_b + cc * 32 + i));
    }
    if (e.flags & CDM_EPI_RESSCALE) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] *= e.res_scale;
    }
    if (e.flags & CDM_EPI_FILM) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        f[i] = fmaf(__ldg(e.film_scale + cc * 32 + i), f[i], __ldg(e.film_shift + cc * 32 + i));
    }
    if (e.flags & CDM_EPI_GNSTATS) {
      float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        s0 += f[i];
        q0 = fmaf(f[i], f[i], q0);
        s1 += f[16 + i];
        q1 = fmaf(f[16 + i], f[16 + i], q1);
      }
      gn_sum[cc * 2] = s0;
      gn_sq[cc * 2] = q0;
      gn_sum[cc * 2 + 1] = s1;
      gn_sq[cc * 2 + 1] = q1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int u = cc * 4 + j;  // 16-byte unit inside this pixel's 256-byte row
      const uint32_t addr = stg + lane * 256 + ((u ^ (lane & 7)) << 4);
      st_shared_v4(addr, pack_bf16x2(f[j * 8 + 0], f[j * 8 + 1]), pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3]),
                   pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5]), pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7]));
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB, const ConvKParams p) {
  using C = ConvCfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smemA = smem;
  uint8_t* smemB = smemA + C::NA * C::A_STRIDE;
  uint8_t* smemStage = smemB + kNB * kBBytes;
  float* s_scale = reinterpret_cast<float*>(smemStage + kStageBytes);
  float* s_shift = s_scale + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::NA;
  uint64_t* b_full = a_empty + C::NA;
  uint64_t* b_empty = b_full + kNB;
  uint64_t* t_full = b_empty + kNB;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.cout; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kNB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int cin = p.chunks * 64;
  const int units_per_img = p.strips_x * p.strips_y * p.n_tiles;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------ A producer (activations)
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    int sa = 0;
    uint32_t pa = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int img = u / units_per_img;
      const int s = (u % units_per_img) / p.n_tiles;
      const int sx = s % p.strips_x, sy = s / p.strips_x;
      for (int ch = 0; ch < p.chunks; ++ch) {
        const CUtensorMap* m = ch < p.chunks0 ? &mapA0 : &mapA1;
        const int c_off = (ch < p.chunks0 ? ch : ch - p.chunks0) * 64;
        for (int j = 0; j < C::A_PER_CHUNK; ++j) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_arrive_expect_tx(&a_full[sa], C::A_BYTES);
          const int w0 = sx * 16 - 1 + (MODE == 0 ? j : 0);
          tma_load_4d(smemA + sa * C::A_STRIDE, m, &a_full[sa], c_off, w0, sy * 16 - 1, img);
          if (++sa == C::NA) {
            sa = 0;
            pa ^= 1;
          }
        }
      }
    }
  } else if (warp == 3 && lane == 0) {
    // ------------------------------------------------ B producer (weights)
    tma_prefetch_desc(&mapB);
    int sb = 0;
    uint32_t pb = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int n_tile = u % p.n_tiles;
      for (int ch = 0; ch < p.chunks; ++ch)
        for (int j = 0; j < C::A_PER_CHUNK; ++j)
          for (int t = 0; t < C::TAPS_PER_A; ++t) {
            const int tap = MODE == 0 ? t * 3 + j : t;
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_arrive_expect_tx(&b_full[sb], kBBytes);
            tma_load_2d(smemB + sb * kBBytes, &mapB, &b_full[sb], tap * cin + ch * 64, n_tile * 128);
            if (++sb == kNB) {
              sb = 0;
              pb ^= 1;
            }
          }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      bool first = true;
      for (int ch = 0; ch < p.chunks; ++ch)
        for (int j = 0; j < C::A_PER_CHUNK; ++j) {
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smemA + sa * C::A_STRIDE);
          for (int t = 0; t < C::TAPS_PER_A; ++t) {
            const int kh = MODE == 0 ? t : t / 3;
            const int kw = MODE == 0 ? 0 : t % 3;  // MODE 0: the copy itself is kw-shifted
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t b_base = smem_u32(smemB + sb * kBBytes);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint32_t a_addr = a_base + (uint32_t)(kh * C::PITCH + mt * 8 + kw) * 128u;
              const uint32_t d_tmem = tmem_base + (uint32_t)((buf * 2 + mt) * 128);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32, C::PITCH * 128),
                          umma_desc_sw128(b_base + k * 32, 1024), idesc, (first && k == 0) ? 0u : 1u);
            }
            first = false;
            umma_commit(&b_empty[sb]);
            if (++sb == kNB) {
              sb = 0;
              pb ^= 1;
            }
          }
          umma_commit(&a_empty[sa]);
          if (++sa == C::NA) {
            sa = 0;
            pa ^= 1;
          }
        }
      umma_commit(&t_full[buf]);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue
    const int q = warp - 4;  // TMEM lane quadrant == warp % 4
    const uint32_t stg = smem_u32(smemStage + q * 8192);
    const int step = p.step_ptr ? *p.step_ptr : 0;
    int it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      const int img = u / units_per_img;
      const int n_tile = (u % units_per_img) % p.n_tiles;
      const int s = (u % units_per_img) / p.n_tiles;
      const int sx = s % p.strips_x, sy = s / p.strips_x;
      mbar_wait(&t_full[buf], (it >> 1) & 1);
      tc_fence_after();

      EpiCtx e;
      e.s_scale = s_scale + n_tile * 128;
      e.s_shift = s_shift + n_tile * 128;
      e.flags = p.flags;
      e.xpix = 0.f;
      e.res_scale = p.res_scale;
      e.sc_w = e.sc_b = e.film_scale = e.film_shift = nullptr;
      if (p.flags & CDM_EPI_FILM) {
        e.film_scale = p.film_scale + (size_t)img * p.cout + n_tile * 128;
        const int r = p.film_shift_rows == 1 ? 0 : img;
        e.film_shift = p.film_shift + ((size_t)step * p.film_shift_rows + r) * p.cout + n_tile * 128;
      }
      // The G1 shortcut can fan one input image out to `sc_reps` outputs (the CFG cond / uncond
      // passes share x and init_conv, and differ only in the fresh 1x1 shortcut).
      const int reps = (p.flags & CDM_EPI_SHORTCUT) ? p.sc_reps : 1;

#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int oh0 = sy * 16, ow0 = sx * 16 + mt * 8;
        const int my_r = 4 * q + (lane >> 3), my_j = lane & 7;  // this thread's pixel inside the patch
        if (p.flags & CDM_EPI_SHORTCUT)
          e.xpix = __ldg(p.sc_x + ((size_t)img * p.H + oh0 + my_r) * p.W + ow0 + my_j);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + mt) * 128);
#pragma unroll 1
        for (int rep = 0; rep < reps; ++rep) {
          if (p.flags & CDM_EPI_SHORTCUT) {
            const float* row = p.sc_tab + ((size_t)(step * p.sc_reps + rep) * 2) * p.cout + n_tile * 128;
            e.sc_w = row;
            e.sc_b = row + p.cout;
          }
          const int oimg = rep * p.n_img + img;
          float gsum[8], gsq[8];
          epi_tile_to_staging(taddr, stg, lane, e, gsum, gsq);
          if (p.flags & CDM_EPI_GNSTATS) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                gsum[g] += __shfl_xor_sync(0xffffffffu, gsum[g], o);
                gsq[g] += __shfl_xor_sync(0xffffffffu, gsq[g], o);
              }
            }
            if (lane == 0) {
              const int slots = p.strips_x * p.strips_y * 8;
              const int slot = (s * 2 + mt) * 4 + q;
              float* dst = p.gn_partial + ((size_t)oimg * slots + slot) * 16;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                dst[g * 2] = gsum[g];
                dst[g * 2 + 1] = gsq[g];
              }
            }
          }
          __syncwarp();
          if (!(p.flags & CDM_EPI_POOL)) {
            // 32 pixels x 256 B; one instruction stores two pixels = 512 contiguous bytes
#pragma unroll 4
            for (int i2 = 0; i2 < 16; ++i2) {
              const int px = 2 * i2 + (lane >> 4), un = lane & 15;
              const uint4 val = ld_shared_v4(stg + px * 256 + ((un ^ (px & 7)) << 4));
              const int r = 4 * q + (px >> 3), jx = px & 7;
              bf16* g = p.out + (((size_t)oimg * p.H + oh0 + r) * p.W + ow0 + jx) * p.cout + n_tile * 128;
              reinterpret_cast<uint4*>(g)[un] = val;
            }
          } else {
            // 2x2 max-pool inside the warp's 4 rows x 8 px -> 2 rows x 4 px
            const int Ho = p.H >> 1, Wo = p.W >> 1;
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
              const int item = i2 * 32 + lane;
              const int pp = item >> 4, un = item & 15;
              const int pr = pp >> 2, pj = pp & 3;
              const int p00 = (2 * pr) * 8 + 2 * pj;
              const uint4 a = ld_shared_v4(stg + p00 * 256 + ((un ^ (p00 & 7)) << 4));
              const uint4 b = ld_shared_v4(stg + (p00 + 1) * 256 + ((un ^ ((p00 + 1) & 7)) << 4));
              const uint4 c = ld_shared_v4(stg + (p00 + 8) * 256 + ((un ^ ((p00 + 8) & 7)) << 4));
              const uint4 d = ld_shared_v4(stg + (p00 + 9) * 256 + ((un ^ ((p00 + 9) & 7)) << 4));
              uint4 m;
              m.x = bf16x2_max(bf16x2_max(a.x, b.x), bf16x2_max(c.x, d.x));
              m.y = bf16x2_max(bf16x2_max(a.y, b.y), bf16x2_max(c.y, d.y));
              m.z = bf16x2_max(bf16x2_max(a.z, b.z), bf16x2_max(c.z, d.z));
              m.w = bf16x2_max(bf16x2_max(a.w, b.w), bf16x2_max(c.w, d.w));
              const int orow = (oh0 >> 1) + 2 * q + pr, ocol = (ow0 >> 1) + pj;
              bf16* g = p.out + (((size_t)oimg * Ho + orow) * Wo + ocol) * p.cout + n_tile * 128;
              reinterpret_cast<uint4*>(g)[un] = m;
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// --------------------------------------------------------------------------
// conv3x3_sw_kernel ("swapped", MODE 3): the weight tile is the M operand (128 output channels)
// and a patch of 8 px x 32 rows = 256 pixels is the N operand of ONE M128 N256 K16 instruction.
// Per FLOP the tensor core then reads 25 % less shared memory than the M128 N128 shape (12 KB per
// 128 cycles instead of 8 KB per 64), which is what bounded MODE 0-2.  TMEM lane = channel, column =
// pixel, so scale/shift/FiLM/shortcut coefficients are per-thread scalars, 2x2 max-pooling and the
// GroupNorm sums are register-local, and only the NHWC store needs a transpose through shared memory.
// Halo tile: TMA box {64 ch, 10, 34, 1}, pitch 10 (SBO 1280 B), tap (kh,kw) = start row kh*10 + kw.
// --------------------------------------------------------------------------
// ROWS = patch height: 32 (N = 256 pixels per instruction, the throughput shape) or 8 (N = 64: four times as many
// units, each with a quarter of the MMA chain — the latency shape for launches that cannot fill the machine).
constexpr int kSwPitch = 10, kSwNA = 3;
template <int ROWS>
struct SwCfg {
  static constexpr int kHalo = ROWS + 2, kNPix = 8 * ROWS, kABytes = kHalo * kSwPitch * 128;
  static constexpr int kAStride = (kABytes + 1023) / 1024 * 1024, kChunksPerHalf = ROWS / 8;
  // Weight ring depth.  ROWS = 32: 4 stages (see below).  ROWS = 8 (the latency shape: one unit per CTA, N = 64): a
  // 16 KB weight tile is consumed in 4 short MMAs (~70 ns) but takes an L2 round trip (~1 us) to arrive, so with 4
  // stages the 18 (K = 1152) or 36 (K = 2304) tiles of a unit arrive in 4.5 / 9 latency-bound turns of the ring;
  // the small halo tiles leave room for 9 stages = 144 KB in flight, which turns the stream bandwidth-bound (and,
  // under PDL, lets half of a layer's weights arrive while the previous kernel drains).
  static constexpr int kNB = ROWS == 8 ? 9 : 4;
};
constexpr int kSwEpiWarps = 8, kSwThreads = 128 + 32 * kSwEpiWarps;
// Weight ring (ROWS = 32): 4 stages.  The issuer is blocked on it ~30 % of its cycles (tools/gpu_probe.py waits), but that is
// back-pressure from the tensor pipe, not TMA latency: 6 stages changed neither the wait share nor the run time.
// The shape itself is at the shared-memory limit: operand reads 12 KB / 128 clk = 96 B/clk plus TMA fills
// (16 KB weights / 512 clk + 43.5 KB halo / 4608 clk = 41 B/clk) against 128 B/clk per SM.
// TMAST (TMA-store epilogue): the NHWC output leaves through shared memory + cp.async.bulk.tensor stores.  A store box
// is {64 ch, 8 px, 4 rows} = 32 rows of 128 B (one 32-pixel accumulator chunk of one channel half), written by the
// two epilogue warps that own those 64 channels; 4 warp pairs x 2 buffers x 4 KB of staging replace the scale/shift
// staging area (the per-thread coefficients come straight from global memory instead).
constexpr int kSwStoreBox = 32 * 128, kSwStageBytes = 4 * 2 * kSwStoreBox;
template <int ROWS, bool TMAST = false>
constexpr int conv_sw_smem_bytes() {
  return kSwNA * SwCfg<ROWS>::kAStride + SwCfg<ROWS>::kNB * kBBytes + (TMAST ? kSwStageBytes : 2 * 256 * 4) + 256 + 1024;
}

template <int ROWS, bool TMAST = false>
__global__ void __launch_bounds__(kSwThreads, 1)
conv3x3_sw_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
                  const ConvKParams p) {
  constexpr int kSwABytes = SwCfg<ROWS>::kABytes, kSwAStride = SwCfg<ROWS>::kAStride, kNPix = SwCfg<ROWS>::kNPix;
  constexpr int kSwNB = SwCfg<ROWS>::kNB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smemA = smem;  // pixel halo tiles
  uint8_t* smemB = smemA + kSwNA * kSwAStride;  // weight tiles
  uint8_t* smemStage = smemB + kSwNB * kBBytes;  // TMAST: store staging (1024-byte aligned); else scale / shift
  float* s_scale = reinterpret_cast<float*>(smemStage);
  float* s_shift = s_scale + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemStage + (TMAST ? kSwStageBytes : 2 * 256 * 4));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kSwNA;
  uint64_t* b_full = a_empty + kSwNA;
  uint64_t* b_empty = b_full + kSwNB;
  uint64_t* t_full = b_empty + kSwNB;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (!TMAST) {
    for (int i = threadIdx.x; i < p.cout; i += blockDim.x) {
      s_scale[i] = p.scale[i];
      s_shift[i] = p.shift[i];
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSwNA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kSwNB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], kSwEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int cin = p.chunks * 64;
  const int px_tiles = p.W >> 3, py_tiles = p.H / ROWS;  // patches of 8 px x ROWS rows
  const int units_per_img = px_tiles * py_tiles * p.n_tiles;
  // PDL: barriers, TMEM and tensor-map prefetch above (and the weight stream below: weights are never written by a
  // stream predecessor that triggers early) may overlap the tail of the previous kernel; every role that touches
  // activations — the halo producer, and the epilogue with its table reads and its stores — waits for it first.
  // Small launches (batch-1 sampling: 64 CTAs on 148 SMs) hide their set-up and the launch latency this way.
  if (threadIdx.x == 0) griddep_launch();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    griddep_wait();
    int sa = 0;
    uint32_t pa = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int img = u / units_per_img;
      const int s = (u % units_per_img) / p.n_tiles;
      const int sx = s % px_tiles, sy = s / px_tiles;
      for (int ch = 0; ch < p.chunks; ++ch) {
        const CUtensorMap* m = ch < p.chunks0 ? &mapA0 : &mapA1;
        const int c_off = (ch < p.chunks0 ? ch : ch - p.chunks0) * 64;
        mbar_wait(&a_empty[sa], pa ^ 1);
        mbar_arrive_expect_tx(&a_full[sa], kSwABytes);
        tma_load_4d(smemA + sa * kSwAStride, m, &a_full[sa], c_off, sx * 8 - 1, sy * ROWS - 1, img);
        if (++sa == kSwNA) {
          sa = 0;
          pa ^= 1;
        }
      }
    }
  } else if (warp == 3 && lane == 0) {
    tma_prefetch_desc(&mapB);
    int sb = 0;
    uint32_t pb = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int n_tile = u % p.n_tiles;
      for (int ch = 0; ch < p.chunks; ++ch)
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          if (CDM_PROBE_BIT(p.flags, 27) && (tap & 1)) {  // probe: half the weight traffic (stale tiles, wrong numbers)
            mbar_arrive(&b_full[sb]);
          } else {
            mbar_arrive_expect_tx(&b_full[sb], kBBytes);
            tma_load_2d(smemB + sb * kBBytes, &mapB, &b_full[sb], tap * cin + ch * 64, n_tile * 128);
          }
          if (++sb == kSwNB) {
            sb = 0;
            pb ^= 1;
          }
        }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, kNPix);
    int sa = 0, sb = 0, it = 0;
    uint32_t pa = 0, pb = 0;
    // probe (flag bit 28): cycles the issuer spends blocked on each barrier class -> gn_partial[cta][0..3]
    const bool prof = CDM_PROBE_BIT(p.flags, 28);
    long long w_t = 0, w_a = 0, w_b = 0, c0 = 0;
    const long long c_start = prof ? clock64() : 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      if (prof) c0 = clock64();
      mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
      if (prof) w_t += clock64() - c0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kNPix);
      bool first = true;
      for (int ch = 0; ch < p.chunks; ++ch) {
        if (prof) c0 = clock64();
        mbar_wait(&a_full[sa], pa);
        if (prof) w_a += clock64() - c0;
        tc_fence_after();
        const uint32_t px_base = smem_u32(smemA + sa * kSwAStride);
        for (int tap = 0; tap < 9; ++tap) {
          if (prof) c0 = clock64();
          mbar_wait(&b_full[sb], pb);
          if (prof) w_b += clock64() - c0;
          tc_fence_after();
          const uint32_t w_base = smem_u32(smemB + sb * kBBytes);
          const uint32_t px_addr = px_base + (uint32_t)((tap / 3) * kSwPitch + tap % 3) * 128u;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, umma_desc_sw128(w_base + k * 32, 1024), umma_desc_sw128(px_addr + k * 32, kSwPitch * 128),
                      idesc, (first && k == 0) ? 0u : 1u);
          first = false;
          umma_commit(&b_empty[sb]);
          if (++sb == kSwNB) {
            sb = 0;
            pb ^= 1;
          }
        }
        umma_commit(&a_empty[sa]);
        if (++sa == kSwNA) {
          sa = 0;
          pa ^= 1;
        }
      }
      umma_commit(&t_full[buf]);
    }
    if (prof) {
      float* d = p.gn_partial + (size_t)blockIdx.x * 8;
      d[0] = (float)w_t;
      d[1] = (float)w_a;
      d[2] = (float)w_b;
      d[3] = (float)(clock64() - c_start);
    }
  } else if (warp >= 4) {
    // 8 epilogue warps: TMEM lane quadrant q = warp % 4 (hardware rule), column half = (warp - 4) / 4
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;  // 0..255 among the epilogue threads
    const int co_l = q * 32 + lane;    // this thread's channel inside the 128-wide tile
    griddep_wait();
    const int step = p.step_ptr ? *p.step_ptr : 0;
    const bool pool = p.flags & CDM_EPI_POOL;
    float bn_s0 = 0.f, bn_q0 = 0.f, bn_s1 = 0.f, bn_q1 = 0.f;  // CDM_EPI_BNSTATS, per n_tile (cout <= 256)
    // TMAST: this warp pair (64 channels of one column half) owns two 4 KB staging boxes; lane 0 of its first warp
    // issues the stores and tracks their bulk groups
    const int ph = q >> 1, pair = half * 2 + ph;
    const bool st_elect = TMAST && (q & 1) == 0 && lane == 0;
    const uint32_t stg_pair = smem_u32(smemStage) + (uint32_t)(pair * 2 * kSwStoreBox);
    int sidx = 0;  // stores issued so far by this pair (buffer = sidx & 1)
    // One 32-pixel chunk (f[i]: pixel row i >> 3, column i & 7 of this thread's channel) -> staging box -> TMA store.
    // Lanes 2k / 2k+1 own channels c, c+1: one shuffle hands the even lane both channels of pixel ia and the odd lane
    // both channels of pixel ia + 4; the 128-byte-swizzle chunk index of a row is XORed with (row & 7), so rows that
    // differ in bit 2 land in different bank halves and every st.shared is one conflict-free 128-byte wavefront.
    auto stage_store = [&](const float (&f)[32], int c_crd, int w_crd, int h_crd, int n_crd) {
      const int odd = lane & 1;
      const uint32_t base = stg_pair + (uint32_t)((sidx & 1) * kSwStoreBox) + (uint32_t)(odd * 4 * 128 + (lane & 6) * 2);
      const uint32_t jsw = (uint32_t)((q & 1) * 4 + (lane >> 3)) ^ (uint32_t)(odd * 4);
      if (CDM_PROBE_BIT(p.flags, 25)) {  // probe: no register -> shared-memory staging (stale bytes are stored)
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += f[i];
        if (acc == 123.456f) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base), "r"(__float_as_uint(acc)) : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int ia = (k >> 2) * 8 + (k & 3), ib = ia + 4;
          const float recv = __shfl_xor_sync(0xffffffffu, odd ? f[ia] : f[ib], 1);
          const uint32_t w = odd ? pack_bf16x2(recv, f[ib]) : pack_bf16x2(f[ia], recv);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (uint32_t)(ia * 128) + ((jsw ^ (uint32_t)(k & 3)) << 4)),
                       "r"(w)
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      if (st_elect) tma_store_wait_read<0>();  // the OTHER buffer's store (one chunk ago) has left shared memory
      named_bar_sync(2 + pair, 64);
      if (st_elect && !CDM_PROBE_BIT(p.flags, 24)) {  // probe bit 24: staging only, no TMA store
        tma_store_4d(&mapOut, stg_pair + (uint32_t)((sidx & 1) * kSwStoreBox), c_crd, w_crd, h_crd, n_crd);
        tma_store_commit();
      }
      ++sidx;
    };
    int it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      const int img = u / units_per_img;
      const int n_tile = (u % units_per_img) % p.n_tiles;
      const int s = (u % units_per_img) / p.n_tiles;
      const int sx = s % px_tiles, sy = s / px_tiles;
      const int oh0 = sy * ROWS, ow0 = sx * 8;
      const int co = n_tile * 128 + co_l;
      const bool prof = CDM_PROBE_BIT(p.flags, 28);
      const long long e0 = prof ? clock64() : 0;
      mbar_wait(&t_full[buf], (it >> 1) & 1);
      const long long e1 = prof ? clock64() : 0;
      tc_fence_after();
      const float sc = TMAST ? __ldg(p.scale + co) : s_scale[co], sh = TMAST ? __ldg(p.shift + co) : s_shift[co];
      float fsv = 1.f, fbv = 0.f;
      if (p.flags & CDM_EPI_FILM) {
        fsv = __ldg(p.film_scale + (size_t)img * p.cout + co);
        fbv = __ldg(p.film_shift + ((size_t)step * p.film_shift_rows + (p.film_shift_rows == 1 ? 0 : img)) * p.cout + co);
      }
      float bw_sc = 0.f, bw_sh = 0.f;
      if (p.flags & CDM_EPI_BNBWD) {
        bw_sc = __ldg(p.bwd_scale + co);
        bw_sh = __ldg(p.bwd_shift + co);
      }
      const int reps = (p.flags & CDM_EPI_SHORTCUT) ? p.sc_reps : 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kNPix);
      // Classifier-free guidance fan-out (init_conv.conv2: one image -> the conditional and the unconditional copy,
      // which differ only in the shortcut row): read the accumulator and the shortcut input ONCE and emit both
      // outputs; as two passes of the generic loop below this epilogue was the bottleneck of its layer.
      const bool fan2 = (p.flags & ~(CDM_EPI_RELU | (1 << 28))) == CDM_EPI_SHORTCUT && reps == 2;
      if (fan2) {
        const float* row0 = p.sc_tab + ((size_t)(step * 2) * 2) * p.cout;
        const float* row1 = row0 + 2 * p.cout;
        const float w0 = __ldg(row0 + co), b0 = __ldg(row0 + p.cout + co);
        const float w1 = __ldg(row1 + co), b1 = __ldg(row1 + p.cout + co);
        const int odd = lane & 1;
#pragma unroll 1
        for (int c8 = half * SwCfg<ROWS>::kChunksPerHalf; c8 < (half + 1) * SwCfg<ROWS>::kChunksPerHalf; ++c8) {
          uint32_t v[32];
          tmem_ld_x32(taddr + c8 * 32, v);
          const float* xr = p.sc_x + ((size_t)img * p.H + oh0 + 4 * c8) * p.W + ow0;
          float xs[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) xs[i] = __ldg(xr + (i >> 3) * p.W + (i & 7));
          tmem_wait_ld();
          float base[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float y = fmaf(__uint_as_float(v[i]), sc, sh);
            base[i] = (p.flags & CDM_EPI_RELU) ? fmaxf(y, 0.f) : y;
          }
#pragma unroll
          for (int rep = 0; rep < 2; ++rep) {
            const float wc = rep ? w1 : w0, bc = rep ? b1 : b0;
            if constexpr (TMAST) {
              float f2[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) f2[i] = base[i] + fmaf(wc, xs[i], bc);
              stage_store(f2, n_tile * 128 + ph * 64, ow0, oh0 + 4 * c8, rep * p.n_img + img);
              continue;
            }
            bf16* gbase = p.out + (((size_t)(rep * p.n_img + img) * p.H + oh0 + 4 * c8) * p.W + ow0) * p.cout +
                          n_tile * 128 + (co_l & ~1);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float f0 = base[i] + fmaf(wc, xs[i], bc), f1 = base[i + 1] + fmaf(wc, xs[i + 1], bc);
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? f0 : f1, 1);
              const uint32_t w = odd ? pack_bf16x2(recv, f1) : pack_bf16x2(f0, recv);
              const int px = i + odd;
              *reinterpret_cast<uint32_t*>(gbase + ((size_t)(px >> 3) * p.W + (px & 7)) * p.cout) = w;
            }
          }
        }
      }
#pragma unroll 1
      for (int rep = 0; rep < (fan2 ? 0 : reps); ++rep) {
        float wcv = 0.f, bcv = 0.f;
        if (p.flags & CDM_EPI_SHORTCUT) {
          const float* row = p.sc_tab + ((size_t)(step * p.sc_reps + rep) * 2) * p.cout;
          wcv = __ldg(row + co);
          bcv = __ldg(row + p.cout + co);
        }
        const int oimg = rep * p.n_img + img;
        float gs = 0.f, gq = 0.f;
#pragma unroll 1
        for (int c8 = half * SwCfg<ROWS>::kChunksPerHalf; c8 < (half + 1) * SwCfg<ROWS>::kChunksPerHalf; ++c8) {  // 32 pixels: patch rows 4*c8 .. 4*c8+3, 8 px each
          if (CDM_PROBE_BIT(p.flags, 29)) continue;  // probe: no epilogue work at all
          uint32_t v[32];
          tmem_ld_x32(taddr + c8 * 32, v);
          // CDM_EPI_BNBWD: this thread's channel of the layer's forward z at the chunk's 32 pixels (a warp-level load
          // covers 32 consecutive channels = 64 bytes of one pixel); issued before the accumulator wait
          unsigned short zr[32];
          if (p.flags & CDM_EPI_BNBWD) {
            const unsigned short* zp = reinterpret_cast<const unsigned short*>(p.bwd_z) +
                                       (((size_t)img * p.H + oh0 + 4 * c8) * p.W + ow0) * p.cout + co;
#pragma unroll
            for (int i = 0; i < 32; ++i) zr[i] = __ldg(zp + ((size_t)(i >> 3) * p.W + (i & 7)) * p.cout);
          }
          tmem_wait_ld();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float y = fmaf(__uint_as_float(v[i]), sc, sh);
            if (p.flags & CDM_EPI_RELU) y = fmaxf(y, 0.f);
            f[i] = y;
          }
          if (p.flags & CDM_EPI_BNBWD) {  // g = stored (bf16) dy under the ReLU mask of the forward; sum g, sum g*z
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float zf = __uint_as_float((uint32_t)zr[i] << 16);
              const float r = __bfloat162float(__float2bfloat16_rn(f[i]));
              const float g = fmaf(zf, bw_sc, bw_sh) > 0.f ? r : 0.f;
              ts += g;
              tq = fmaf(g, zf, tq);
            }
            if (n_tile == 0) {
              bn_s0 += ts;
              bn_q0 += tq;
            } else {
              bn_s1 += ts;
              bn_q1 += tq;
            }
          }
          if (p.flags & CDM_EPI_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
          }
          if (p.flags & CDM_EPI_SHORTCUT) {
            const float* xr = p.sc_x + ((size_t)img * p.H + oh0 + 4 * c8) * p.W + ow0;
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += fmaf(wcv, __ldg(xr + (i >> 3) * p.W + (i & 7)), bcv);
          }
          if (p.flags & CDM_EPI_RESSCALE) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] *= p.res_scale;
          }
          if (p.flags & CDM_EPI_FILM) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaf(fsv, f[i], fbv);
          }
          if (p.flags & CDM_EPI_GNSTATS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              gs += f[i];
              gq = fmaf(f[i], f[i], gq);
            }
          }
          if (p.flags & CDM_EPI_BNSTATS) {  // statistics of what is stored (bf16), so BatchNorm normalises z exactly
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float r = __bfloat162float(__float2bfloat16_rn(f[i]));
              ts += r;
              tq = fmaf(r, r, tq);
            }
            if (n_tile == 0) {
              bn_s0 += ts;
              bn_q0 += tq;
            } else {
              bn_s1 += ts;
              bn_q1 += tq;
            }
          }
          // NHWC store without shared memory: lanes 2k / 2k+1 own channels co, co+1.  One shuffle per pixel
          // pair gives the even lane both channels of pixel i and the odd lane both channels of pixel i+1, so a
          // warp instruction writes 2 x 64 contiguous bytes (full sectors) as 32-bit bf16x2 words.
          const int odd = lane & 1;
          if (CDM_PROBE_BIT(p.flags, 30)) {  // probe: everything but the global stores
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += f[i];
            if (acc == 123.456f) p.out[0] = __float2bfloat16(acc);
            continue;
          }
          if (TMAST && !pool) {
            stage_store(f, n_tile * 128 + ph * 64, ow0, oh0 + 4 * c8, oimg);
          } else if (!pool) {
            bf16* gbase = p.out + (((size_t)oimg * p.H + oh0 + 4 * c8) * p.W + ow0) * p.cout + n_tile * 128 +
                          (co_l & ~1);
            int wrow = p.W;
            if (CDM_PROBE_BIT(p.flags, 26)) {  // probe: same stores, but into a per-CTA 64 KB window that stays in L2
              gbase = p.out + (size_t)blockIdx.x * 32768 + (size_t)(4 * c8) * 8 * p.cout + (co_l & ~1);
              wrow = 8;
            }
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? f[i] : f[i + 1], 1);
              const uint32_t w = odd ? pack_bf16x2(recv, f[i + 1]) : pack_bf16x2(f[i], recv);
              const int px = i + odd;  // rows 4*c8 + (px >> 3), column px & 7
              *reinterpret_cast<uint32_t*>(gbase + ((size_t)(px >> 3) * wrow + (px & 7)) * p.cout) = w;
            }
          } else {
            float m[8];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr)
#pragma unroll
              for (int pj = 0; pj < 4; ++pj) {
                const int i00 = (2 * pr) * 8 + 2 * pj;
                m[pr * 4 + pj] = fmaxf(fmaxf(f[i00], f[i00 + 1]), fmaxf(f[i00 + 8], f[i00 + 9]));
              }
            const int Ho = p.H >> 1, Wo = p.W >> 1;
            bf16* gbase = p.out + (((size_t)oimg * Ho + (oh0 >> 1) + 2 * c8) * Wo + (ow0 >> 1)) * p.cout +
                          n_tile * 128 + (co_l & ~1);
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? m[i] : m[i + 1], 1);
              const uint32_t w = odd ? pack_bf16x2(recv, m[i + 1]) : pack_bf16x2(m[i], recv);
              const int px = i + odd;  // pooled row 2*c8 + (px >> 2), pooled column px & 3
              *reinterpret_cast<uint32_t*>(gbase + ((size_t)(px >> 2) * Wo + (px & 3)) * p.cout) = w;
            }
          }
        }
        if (p.flags & CDM_EPI_GNSTATS) {
          // 16 channels per group = 16 consecutive lanes
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            gs += __shfl_xor_sync(0xffffffffu, gs, o);
            gq += __shfl_xor_sync(0xffffffffu, gq, o);
          }
          const int slots = (p.H >> 4) * (p.W >> 4) * 8;  // API layout; this mode owns 8 slots per patch:
          float* dst = p.gn_partial + ((size_t)oimg * slots + (size_t)s * 8) * 16;  // slot `half` + six zero slots
          if ((lane & 15) == 0) {
            const int g = q * 2 + (lane >> 4);
            dst[half * 16 + g * 2] = gs;
            dst[half * 16 + g * 2 + 1] = gq;
          }
          for (int z = et; z < 6 * 16; z += 256) dst[32 + z] = 0.f;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
      if (prof && et == 0) {  // warp 4: cycles idle (waiting for the accumulator) / busy (epilogue proper)
        float* d = p.gn_partial + (size_t)blockIdx.x * 8;
        d[4] = (it == 0 ? 0.f : d[4]) + (float)(e1 - e0);
        d[5] = (it == 0 ? 0.f : d[5]) + (float)(clock64() - e1);
      }
    }
    if (st_elect) tma_store_wait_read<0>();  // the staging boxes are read before the CTA (and its shared memory) goes
    if (p.flags & (CDM_EPI_BNSTATS | CDM_EPI_BNBWD)) {
      // the two column halves of a channel live in warps q and q+4: fold them through shared memory (the
      // scale/shift staging area is free once the unit loop is over), then one partial row per CTA
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float* fold = s_scale;  // [2 n_tiles][128][2] floats = 2 KB
      if (half == 1) {
        fold[(0 * 128 + co_l) * 2] = bn_s0;
        fold[(0 * 128 + co_l) * 2 + 1] = bn_q0;
        fold[(1 * 128 + co_l) * 2] = bn_s1;
        fold[(1 * 128 + co_l) * 2 + 1] = bn_q1;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0) {
        float* dst = p.bn_partial + (size_t)blockIdx.x * 2 * p.cout;
        float s0 = bn_s0 + fold[(0 * 128 + co_l) * 2], q0 = bn_q0 + fold[(0 * 128 + co_l) * 2 + 1];
        float s1 = bn_s1 + fold[(1 * 128 + co_l) * 2], q1 = bn_q1 + fold[(1 * 128 + co_l) * 2 + 1];
        if (p.flags & CDM_EPI_BNBWD) {  // sum g*xhat = rstd * (sum g*z - mean * sum g): linear, so the CTA rows add up
          q0 = (q0 - __ldg(p.bwd_mean + co_l) * s0) * __ldg(p.bwd_rstd + co_l);
          if (p.n_tiles == 2) q1 = (q1 - __ldg(p.bwd_mean + 128 + co_l) * s1) * __ldg(p.bwd_rstd + 128 + co_l);
        }
        dst[co_l] = s0;
        dst[p.cout + co_l] = q0;
        if (p.n_tiles == 2) {
          dst[128 + co_l] = s1;
          dst[p.cout + 128 + co_l] = q1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// --------------------------------------------------------------------------
// gemm: C[M][N] = A[M][K] * Bw[N][K]^T, work unit = one 128x128 output tile.
// --------------------------------------------------------------------------
constexpr int kGemmStages = 5;
constexpr int kGemmStageBytes = 2 * 16384;
constexpr int gemm_smem_bytes() { return kGemmStages * kGemmStageBytes + kStageBytes + 256 + 1024; }

__global__ void __launch_bounds__(kConvThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
            const __grid_constant__ CUtensorMap mapB, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smemAB = smem;
  uint8_t* smemStage = smemAB + kGemmStages * kGemmStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemStage + kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kGemmStages;
  uint64_t* t_full = empty + kGemmStages;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGemmStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int cps = (p.chunks + p.k_split - 1) / p.k_split;  // K chunks per slice

  // unit -> (m_tile, n_tile): n fastest so concurrently running CTAs share the A rows in L2
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
    int st = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int mn = u % (p.m_tiles * p.n_tiles), slice = u / (p.m_tiles * p.n_tiles);
      const int m_tile = mn / p.n_tiles, n_tile = mn % p.n_tiles;
      const int ch0 = slice * cps, ch1 = min(p.chunks, ch0 + cps);
      for (int ch = ch0; ch < ch1; ++ch) {
        const CUtensorMap* m = ch < p.chunks0 ? &mapA0 : &mapA1;
        const int c_off = (ch < p.chunks0 ? ch : ch - p.chunks0) * 64;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&full[st], kGemmStageBytes);
        tma_load_2d(smemAB + st * kGemmStageBytes, m, &full[st], c_off, m_tile * 128);
        tma_load_2d(smemAB + st * kGemmStageBytes + 16384, &mapB, &full[st], ch * 64, n_tile * 128);
        if (++st == kGemmStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    int st = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const int slice = u / (p.m_tiles * p.n_tiles);
      const int ch0 = slice * cps, ch1 = min(p.chunks, ch0 + cps);
      for (int ch = ch0; ch < ch1; ++ch) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smemAB + st * kGemmStageBytes);
        const uint32_t b_base = a_base + 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + buf * 128, umma_desc_sw128(a_base + k * 32, 1024),
                    umma_desc_sw128(b_base + k * 32, 1024), idesc, (ch == ch0 && k == 0) ? 0u : 1u);
        umma_commit(&empty[st]);
        if (++st == kGemmStages) {
          st = 0;
          ph ^= 1;
        }
      }
      umma_commit(&t_full[buf]);
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const uint32_t stg = smem_u32(smemStage + q * 8192);
    int it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      const int mn = u % (p.m_tiles * p.n_tiles), slice = u / (p.m_tiles * p.n_tiles);
      const int m_tile = mn / p.n_tiles, n_tile = mn % p.n_tiles;
      mbar_wait(&t_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128);
      if (p.partial) {  // split-K: this slice's fp32 tile row, no shift, no rounding (an empty slice stores zeros)
        float4* prow = reinterpret_cast<float4*>(
            p.partial + ((size_t)slice * p.m_tiles * 128 + m_tile * 128 + q * 32 + lane) * p.N + n_tile * 128);
        const bool live = slice * cps < p.chunks;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          if (live) {
            tmem_ld_x32(taddr + cc * 32, v);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            prow[cc * 8 + i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                           __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[buf]);
        continue;
      }
      const float* shift = p.shift + (n_tile * 128) % p.shift_mod;  // shift_mod is a multiple of 128
      // this thread's output row: 128 contiguous bf16 (256 B) written straight from registers
      const int m = m_tile * 128 + q * 32 + lane;
      bf16* g = nullptr;
      if (m < p.M) {
        if (p.out_mode == 0) {
          g = p.out + (size_t)m * p.N + n_tile * 128;
        } else {
          // H, W are powers of two (checked on the host): shifts instead of integer division
          const int w = m & (p.W - 1), h = (m >> p.w_shift) & (p.H - 1), img = m >> (p.w_shift + p.h_shift);
          const int kh = n_tile >> 1, kw = n_tile & 1;
          g = p.out + (((size_t)img * 2 * p.H + 2 * h + kh) * (2 * p.W) + 2 * w + kw) * 128;
        }
      }
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        tmem_ld_x32(taddr + cc * 32, v);
        tmem_wait_ld();
        if (g) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(shift + cc * 32 + j * 8));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(shift + cc * 32 + j * 8 + 4));
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]) + s0.x, __uint_as_float(v[j * 8 + 1]) + s0.y);
            o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]) + s0.z, __uint_as_float(v[j * 8 + 3]) + s0.w);
            o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]) + s1.x, __uint_as_float(v[j * 8 + 5]) + s1.y);
            o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]) + s1.z, __uint_as_float(v[j * 8 + 7]) + s1.w);
            reinterpret_cast<uint4*>(g)[cc * 4 + j] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// --------------------------------------------------------------------------

// out[m][n] = bf16( sum_slice partial[slice][m][n] + shift[n % shift_mod] ), slices added in order (deterministic).
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const float* __restrict__ partial, int k_split, int M,
                                                                 int M_pad, int N, const float* __restrict__ shift,
                                                                 int shift_mod, bf16* __restrict__ out) {
  const size_t idx = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (idx >= (size_t)M * N) return;
  const int n = (int)(idx % N);
  const size_t m = idx / N;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int sl = 0; sl < k_split; ++sl) {
    const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)sl * M_pad + m) * N + n);
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  const float* sh = shift + n % shift_mod;
  uint2 o;
  o.x = pack_bf16x2(acc.x + sh[0], acc.y + sh[1]);
  o.y = pack_bf16x2(acc.z + sh[2], acc.w + sh[3]);
  *reinterpret_cast<uint2*>(out + idx) = o;
}

// --------------------------------------------------------------------------
// gemm_bres: same contraction as gemm_kernel for SMALL N (the 2x2 transposed convolutions: N = 512,
// K = 256 / 512).  The weight rows a CTA needs (n_res 128-row tiles, all of K: 128 KB) are loaded once and
// stay resident in shared memory; the CTA then streams A row tiles through a 4-stage ring and issues
// n_res accumulations per stage, so L2->smem traffic per output tile drops from 2*K*256 B to K*256/n_res B
// and the kernel sits at the DRAM roofline of its output instead of the L2 refill rate.
// CTA b serves weight group b % n_groups and the row tiles b / n_groups, + gridDim.x / n_groups, ...
// With MORE groups than CTAs (up0: N = 65536 = 256 groups of two n-tiles, 16 row tiles) CTA b serves the groups
// b, b + gridDim.x, ... with all row tiles each: the resident weights are swapped between groups once the last
// MMA that reads them has completed (w_free), the A rows (1 MB, L2-resident) are streamed again per group, and
// L2->smem traffic per 128x256 output tile falls from 192 KB (gemm_kernel: A and B per tile) to 64 KB + 8 KB.
// --------------------------------------------------------------------------
constexpr int kBresStages = 4;
constexpr int kBresStageTile = 32 * 128;  // per epilogue warp: 32 rows x 128 B of bf16 output, 16-byte chunks XOR-swizzled
constexpr int gemm_bres_smem_bytes() {
  return 131072 + kBresStages * 16384 + 8 * kBresStageTile + 256 + 1024 + 1024;
}

constexpr int kBresThreads = 128 + 8 * 32;  // 4 role warps + 8 epilogue warps (one set of 4 per resident n-tile)
__global__ void __launch_bounds__(kBresThreads, 1)
gemm_bres_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
                 const GemmBresKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smemW = smem;            // [n_res][chunks][128 rows][64 ch]
  uint8_t* smemA = smem + 131072;   // ring of [128 rows][64 ch]
  uint8_t* smemStage = smemA + kBresStages * 16384;  // 8 x kBresStageTile, 1024-byte aligned (TMA-store source boxes)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemStage + 8 * kBresStageTile);
  uint64_t* full = bars;
  uint64_t* empty = full + kBresStages;
  uint64_t* t_full = empty + kBresStages;
  uint64_t* t_empty = t_full + 2;
  uint64_t* w_ready = t_empty + 2;
  uint64_t* w_free = w_ready + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_free + 1);
  float* s_shift = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);  // [n_res][128], 16-byte aligned
  const uint32_t stage_base = smem_u32(smemStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.n_res * 128; i += blockDim.x)
    s_shift[i] = p.shift[(((blockIdx.x % p.n_groups) * p.n_res) * 128 + i) % p.shift_mod];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBresStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 8);
    }
    mbar_init(w_ready, 1);
    mbar_init(w_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // groups <= CTAs: one group per CTA, its row tiles shared by gridDim.x / n_groups CTAs; else: several groups per
  // CTA with all row tiles each (the host guarantees the shift rows are the same for every group in that case)
  const bool many = p.n_groups > (int)gridDim.x;
  const int g_first = many ? (int)blockIdx.x : (int)blockIdx.x % p.n_groups, g_stride = many ? (int)gridDim.x : p.n_groups;
  const int first = many ? 0 : (int)blockIdx.x / p.n_groups, stride = many ? 1 : (int)gridDim.x / p.n_groups;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
    int st = 0, gi = 0;
    uint32_t ph = 0;
    for (int group = g_first; group < p.n_groups; group += g_stride, ++gi) {
      const int n_tile0 = group * p.n_res;
      if (gi > 0) mbar_wait(w_free, (uint32_t)((gi - 1) & 1));  // every MMA of the previous group has read its weights
      mbar_arrive_expect_tx(w_ready, (uint32_t)(p.n_res * p.chunks * 16384));
      for (int nt = 0; nt < p.n_res; ++nt)  // [chunk][n-tile]: the n-tiles of one K chunk are one contiguous N operand
        for (int ch = 0; ch < p.chunks; ++ch)
          tma_load_2d(smemW + (ch * p.n_res + nt) * 16384, &mapB, w_ready, ch * 64, (n_tile0 + nt) * 128);
      for (int mt = first; mt < p.m_tiles; mt += stride)
        for (int ch = 0; ch < p.chunks; ++ch) {
          const CUtensorMap* m = ch < p.chunks0 ? &mapA0 : &mapA1;
          const int c_off = (ch < p.chunks0 ? ch : ch - p.chunks0) * 64;
          mbar_wait(&empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&full[st], 16384);
          tma_load_2d(smemA + st * 16384, m, &full[st], c_off, mt * 128);
          if (++st == kBresStages) {
            st = 0;
            ph ^= 1;
          }
        }
    }
  } else if (warp == 1 && lane == 0) {
    // two resident n-tiles = ONE M128 N256 instruction per K step (12 KB of operand reads per 128 clk instead of
    // 2 x 8 KB per 2 x 64: the N128 shape is shared-memory bound at half the tensor rate)
    const uint32_t idesc = p.n_res == 2 ? umma_idesc_bf16(128, 256) : umma_idesc_bf16(128, 128);
    const uint32_t w_base = smem_u32(smemW);
    int st = 0, it = 0, gi = 0;
    uint32_t ph = 0;
    for (int group = g_first; group < p.n_groups; group += g_stride, ++gi) {
      mbar_wait(w_ready, (uint32_t)(gi & 1));
      tc_fence_after();
      for (int mt = first; mt < p.m_tiles; mt += stride, ++it) {
        const int buf = it & 1;
        mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int ch = 0; ch < p.chunks; ++ch) {
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smemA + st * 16384);
          const uint32_t b_base = w_base + (uint32_t)(ch * p.n_res * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + (uint32_t)(buf * p.n_res * 128), umma_desc_sw128(a_base + k * 32, 1024),
                      umma_desc_sw128(b_base + k * 32, 1024), idesc, (ch == 0 && k == 0) ? 0u : 1u);
          umma_commit(&empty[st]);
          if (++st == kBresStages) {
            st = 0;
            ph ^= 1;
          }
        }
        umma_commit(&t_full[buf]);
      }
      umma_commit(w_free);  // arrives when every MMA issued so far (all that read this group's weights) is done
    }
  } else if (warp >= 4) {
    // TMEM lane quadrant q = warp % 4 (hardware rule); warps 4-7 drain n-tile 0, warps 8-11 n-tile 1 (with one
    // resident n-tile the second set only keeps the barrier protocol): the epilogue, not the MMA, paced this kernel
    const int q = warp & 3, set = (warp - 4) >> 2;
    int it = 0;
    for (int group = g_first; group < p.n_groups; group += g_stride)
    for (int mt = first; mt < p.m_tiles; mt += stride, ++it) {
      const int n_tile0 = group * p.n_res;
      const int buf = it & 1;
      mbar_wait(&t_full[buf], (it >> 1) & 1);
      tc_fence_after();
      // The accumulator gives a thread 32 columns of ONE row, so a register-direct store would touch 32 different
      // 128-byte lines per warp instruction (16 B each).  Rows are transposed through a warp-private staging tile
      // instead: lane = row writes its 64 columns (128 B) as eight 16-byte chunks, chunk c of row r at
      // r*128 + ((c ^ (r & 7)) << 4) (conflict-free both ways); reading back, 8 lanes hold one row's 128
      // contiguous bytes and every warp-level global store writes FOUR FULL LINES.
      const uint32_t stage = stage_base + (uint32_t)((warp - 4) * kBresStageTile);
      const int sub_row = lane >> 3, sub_chunk = lane & 7;
      for (int nt = set; nt < p.n_res; nt += 2) {
        const int n_tile = n_tile0 + nt;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.n_res + nt) * 128);
        const float* shift = s_shift + nt * 128;  // shared memory: broadcast reads, no global-load latency
        if (p.out_mode == 1) {
          // Pixel-shuffle output through the TMA engine (UTMASTG): the staging tile in the layout above IS a
          // 128-byte-swizzle box {64 ch, kw 1, 32 px (or W px x 32/W rows), kh 1} of the output seen as
          // [img*H + h][kh][w][kw][128 ch]: this warp's 32 consecutive rows m are 32 consecutive pixels.  One elected
          // lane stores 4 KB per half instead of eight 16-byte loads + stores per lane (the LSU's shared-memory
          // wavefronts, at 50 % of their peak in ncu, halve), and the store leaves while the next half is read.
          const int m0 = mt * 128 + q * 32;
          const int w0 = p.W >= 32 ? (m0 & (p.W - 1)) : 0, hh0 = m0 >> p.w_shift;
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            if (lane == 0) tma_store_wait_read<0>();  // the previous store has left the staging tile
            __syncwarp();
#pragma unroll 1
            for (int c2 = 0; c2 < 2; ++c2) {
              const int col0 = (half * 2 + c2) * 32;
              uint32_t v[32];
              tmem_ld_x32(taddr + col0, v);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 s0 = *reinterpret_cast<const float4*>(shift + col0 + j * 8);
                const float4 s1 = *reinterpret_cast<const float4*>(shift + col0 + j * 8 + 4);
                const int chunk = c2 * 4 + j;
                st_shared_v4(stage + (uint32_t)(lane * 128 + ((chunk ^ (lane & 7)) << 4)),
                             pack_bf16x2(__uint_as_float(v[j * 8 + 0]) + s0.x, __uint_as_float(v[j * 8 + 1]) + s0.y),
                             pack_bf16x2(__uint_as_float(v[j * 8 + 2]) + s0.z, __uint_as_float(v[j * 8 + 3]) + s0.w),
                             pack_bf16x2(__uint_as_float(v[j * 8 + 4]) + s1.x, __uint_as_float(v[j * 8 + 5]) + s1.y),
                             pack_bf16x2(__uint_as_float(v[j * 8 + 6]) + s1.z, __uint_as_float(v[j * 8 + 7]) + s1.w));
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && m0 < p.M) {
              tma_store_5d(&mapOut, stage, half * 64, n_tile & 1, w0, n_tile >> 1, hh0);
              tma_store_commit();
            }
          }
          continue;
        }
        size_t row_off[8];  // element offset of the eight rows this lane stores (row = sub_row + 4*it); ~0 = beyond M
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = mt * 128 + q * 32 + sub_row + 4 * it;
          row_off[it] = m >= p.M ? ~(size_t)0 : (size_t)m * p.N + n_tile * 128;  // out_mode 0: plain row-major
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
#pragma unroll 1
          for (int c2 = 0; c2 < 2; ++c2) {
            const int col0 = (half * 2 + c2) * 32;
            uint32_t v[32];
            tmem_ld_x32(taddr + col0, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 s0 = *reinterpret_cast<const float4*>(shift + col0 + j * 8);
              const float4 s1 = *reinterpret_cast<const float4*>(shift + col0 + j * 8 + 4);
              const int chunk = c2 * 4 + j;
              st_shared_v4(stage + (uint32_t)(lane * 128 + ((chunk ^ (lane & 7)) << 4)),
                           pack_bf16x2(__uint_as_float(v[j * 8 + 0]) + s0.x, __uint_as_float(v[j * 8 + 1]) + s0.y),
                           pack_bf16x2(__uint_as_float(v[j * 8 + 2]) + s0.z, __uint_as_float(v[j * 8 + 3]) + s0.w),
                           pack_bf16x2(__uint_as_float(v[j * 8 + 4]) + s1.x, __uint_as_float(v[j * 8 + 5]) + s1.y),
                           pack_bf16x2(__uint_as_float(v[j * 8 + 6]) + s1.z, __uint_as_float(v[j * 8 + 7]) + s1.w));
            }
          }
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = sub_row + 4 * it;
            const uint4 o = ld_shared_v4(stage + (uint32_t)(row * 128 + ((sub_chunk ^ (row & 7)) << 4)));
            if (row_off[it] != ~(size_t)0)
              *reinterpret_cast<uint4*>(p.out + row_off[it] + half * 64 + sub_chunk * 8) = o;
          }
          __syncwarp();  // the tile is rewritten by the next half
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
    }
    if (lane == 0) tma_store_wait_read<0>();  // the staging tiles are read before the CTA's shared memory goes
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// --------------------------------------------------------------------------
// gemm_tn: C[m][n] (+)= sum_k A[k][m] * B[k][n]  — both operands "MN-major": the reduction
// index k is the ROW (a pixel / a sample), the output indices are the contiguous channels.
// This is every weight gradient of the network (conv3x3 wgrad per tap with a pixel-shifted B,
// transposed-conv wgrad, up0 wgrad).  A K block is 128 rows = one TMA box {64 ch, bw, bh, 1}
// per 64-channel half (the same NHWC bytes the forward reads as a K-major A operand, here
// described to the tensor core as MN-major).  Split-K over CTAs, fp32 atomics into C.
// --------------------------------------------------------------------------
struct GemmTnKParams {
  int m_tiles, n_tiles, taps, k_split;
  int k_blocks;                 // total 128-row K blocks
  int bw, bh;                   // box extent in w / h (bw * bh == 128)
  int tiles_x, tiles_y;         // K block index -> (img, ty, tx)
  int n_units;
  int a_off, b_off;             // first channel of the A / B windows
  float* C;
  int ldc, tap_stride;
  float* probe;
  float* partial;  // gemm_tn9: [unit][128][384] split-K tiles (NULL: atomics into C)
};
constexpr int kTnStages = 3;
constexpr int kTnStageBytes = 4 * 16384;
constexpr int gemm_tn_smem_bytes() { return kTnStages * kTnStageBytes + 256 + 1024; }

__global__ void __launch_bounds__(kConvThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const GemmTnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTnStages * kTnStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kTnStages;
  uint64_t* t_full = empty + kTnStages;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTnStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int kb_per_slice = (p.k_blocks + p.k_split - 1) / p.k_split;
  // unit -> (slice, tap, m_tile, n_tile); slice slowest so that concurrently running CTAs share K blocks in L2
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    int st = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int n_tile = u % p.n_tiles, m_tile = (u / p.n_tiles) % p.m_tiles;
      const int tap = (u / (p.n_tiles * p.m_tiles)) % p.taps, slice = u / (p.n_tiles * p.m_tiles * p.taps);
      const int dh = p.taps == 9 ? tap / 3 - 1 : 0, dw = p.taps == 9 ? tap % 3 - 1 : 0;
      const int kb0 = slice * kb_per_slice, kb1 = min(p.k_blocks, kb0 + kb_per_slice);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int tx = kb % p.tiles_x, ty = (kb / p.tiles_x) % p.tiles_y, img = kb / (p.tiles_x * p.tiles_y);
        mbar_wait(&empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&full[st], kTnStageBytes);
        uint8_t* sp = smem + st * kTnStageBytes;
        const int ca = p.a_off + m_tile * 128, cb = p.b_off + n_tile * 128;
        tma_load_4d(sp, &mapA, &full[st], ca, tx * p.bw, ty * p.bh, img);
        tma_load_4d(sp + 16384, &mapA, &full[st], ca + 64, tx * p.bw, ty * p.bh, img);
        tma_load_4d(sp + 32768, &mapB, &full[st], cb, tx * p.bw + dw, ty * p.bh + dh, img);
        tma_load_4d(sp + 49152, &mapB, &full[st], cb + 64, tx * p.bw + dw, ty * p.bh + dh, img);
        if (++st == kTnStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16_mn(128, 128);
    int st = 0, it = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int slice = u / (p.n_tiles * p.m_tiles * p.taps);
      const int kb0 = slice * kb_per_slice, kb1 = min(p.k_blocks, kb0 + kb_per_slice);
      const int buf = it & 1;
      mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + st * kTnStageBytes), b_base = a_base + 32768;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // 16 K rows (two 8-row swizzle atoms) per instruction
          umma_bf16(tmem_base + buf * 128, umma_desc_sw128_mn(a_base + ks * 2048, 16384, 1024),
                    umma_desc_sw128_mn(b_base + ks * 2048, 16384, 1024), idesc, (kb == kb0 && ks == 0) ? 0u : 1u);
        umma_commit(&empty[st]);
        if (++st == kTnStages) {
          st = 0;
          ph ^= 1;
        }
      }
      umma_commit(&t_full[buf]);
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int n_tile = u % p.n_tiles, m_tile = (u / p.n_tiles) % p.m_tiles;
      const int tap = (u / (p.n_tiles * p.m_tiles)) % p.taps, slice = u / (p.n_tiles * p.m_tiles * p.taps);
      const int kb0 = slice * kb_per_slice, kb1 = min(p.k_blocks, kb0 + kb_per_slice);
      const int buf = it & 1;
      mbar_wait(&t_full[buf], (it >> 1) & 1);
      tc_fence_after();
      if (p.partial) {  // split-K tile of this slice -> workspace (zeros for an empty slice); gemm_tn_reduce adds them
        float4* prow = reinterpret_cast<float4*>(p.partial + ((size_t)u * 128 + q * 32 + lane) * 128);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          if (kb1 > kb0) {
            tmem_ld_x32(taddr + cc * 32, v);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            prow[cc * 8 + i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                           __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
      } else if (kb1 > kb0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128);
        float* crow = p.C + (size_t)(m_tile * 128 + q * 32 + lane) * p.ldc + tap * p.tap_stride + n_tile * 128;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          tmem_ld_x32(taddr + cc * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(crow + cc * 32 + i, __uint_as_float(v[i]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// Fixed-order sum of the split-K tiles of gemm_tn_kernel: C[...] += sum_slice partial[slice][tap][m_tile][n_tile][row][col].
__global__ void __launch_bounds__(256) gemm_tn_reduce_kernel(const float* __restrict__ partial, int k_split, int taps,
                                                             int m_tiles, int n_tiles, float* __restrict__ C, int ldc,
                                                             int tap_stride) {
  const size_t tile = (size_t)128 * 128;
  const size_t per_slice = (size_t)taps * m_tiles * n_tiles * tile;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_slice) return;
  const int col = (int)(idx % 128), row = (int)((idx / 128) % 128);
  const int un = (int)(idx / tile);  // (tap * m_tiles + m_tile) * n_tiles + n_tile
  const int n_tile = un % n_tiles, m_tile = (un / n_tiles) % m_tiles, tap = un / (n_tiles * m_tiles);
  float acc = 0.f;
#pragma unroll 8
  for (int sl = 0; sl < k_split; ++sl) acc += partial[(size_t)sl * per_slice + idx];  // loads in flight, fixed order
  C[(size_t)(m_tile * 128 + row) * ldc + tap * tap_stride + n_tile * 128 + col] += acc;
}

// --------------------------------------------------------------------------
// gemm_tn9: the 3x3-convolution weight gradient, three taps per instruction.
//   dW[co][kh][kw][ci] += sum_px dz[px][co] * x[px + (kh-1, kw-1)][ci]
// gemm_tn_kernel spends one M128 N128 K16 instruction stream per tap and re-loads both operands for every
// tap: 8 KB of operand reads per 64 tensor-core cycles plus 64 KB of TMA fill per 512 cycles = 256 B/clk against
// the 128 B/clk of shared memory, i.e. half rate (measured 805 TFLOP/s).  Here a K block is a patch of 8 px x 16
// rows; for one kernel row kh the x operand is loaded ONCE with its left/right halo (box {64 ch, 10, 16}) and the
// three kw taps are three 64-channel "MN blocks" of ONE descriptor whose leading-dimension byte offset is 128 B =
// one pixel: block j of the B operand is the same shared-memory tile shifted by j pixels.  One M128 N192 K16
// instruction per channel half therefore produces dW[:, kh, 0..2, half]: operand reads 10 KB / 96 clk, fill 72 KB
// per 1536 clk = 151 B/clk.  Units = (K slice, kh, m_tile, n_tile); accumulators 2 x 192 TMEM columns; split-K
// partial sums are added to C with fp32 atomics (the caller zero-fills C).
// --------------------------------------------------------------------------
constexpr int kTn9Stages = 3;
constexpr int kTn9ABytes = 2 * 16384;            // dz: two channel halves of {64 ch, 8, 16}
constexpr int kTn9BHalf = 10 * 16 * 128;         // x : {64 ch, 10, 16} = 20480 B per channel half
constexpr int kTn9StageBytes = kTn9ABytes + 2 * kTn9BHalf;
constexpr int gemm_tn9_smem_bytes() { return kTn9Stages * kTn9StageBytes + 256 + 1024; }

__global__ void __launch_bounds__(kConvThreads, 1)
gemm_tn9_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const GemmTnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTn9Stages * kTn9StageBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kTn9Stages;
  uint64_t* t_full = empty + kTn9Stages;
  uint64_t* t_empty = t_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTn9Stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(t_full, 1);
    mbar_init(t_empty, 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int kb_per_slice = (p.k_blocks + p.k_split - 1) / p.k_split;
  // unit -> (slice, kh, m_tile, n_tile), slice slowest: concurrently running CTAs share K blocks in L2
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    int st = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int n_tile = u % p.n_tiles, m_tile = (u / p.n_tiles) % p.m_tiles;
      const int kh = (u / (p.n_tiles * p.m_tiles)) % 3, slice = u / (p.n_tiles * p.m_tiles * 3);
      const int kb0 = slice * kb_per_slice, kb1 = min(p.k_blocks, kb0 + kb_per_slice);
      const int ca = p.a_off + m_tile * 128, cb = p.b_off + n_tile * 128;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int tx = kb % p.tiles_x, ty = (kb / p.tiles_x) % p.tiles_y, img = kb / (p.tiles_x * p.tiles_y);
        mbar_wait(&empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&full[st], kTn9StageBytes);
        uint8_t* sp = smem + st * kTn9StageBytes;
        tma_load_4d(sp, &mapA, &full[st], ca, tx * 8, ty * 16, img);
        tma_load_4d(sp + 16384, &mapA, &full[st], ca + 64, tx * 8, ty * 16, img);
        tma_load_4d(sp + kTn9ABytes, &mapB, &full[st], cb, tx * 8 - 1, ty * 16 + kh - 1, img);
        tma_load_4d(sp + kTn9ABytes + kTn9BHalf, &mapB, &full[st], cb + 64, tx * 8 - 1, ty * 16 + kh - 1, img);
        if (++st == kTn9Stages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16_mn(128, 192);
    int st = 0, it = 0;
    uint32_t ph = 0;
    long long w_full = 0, w_acc = 0, c0 = 0;
    const long long c_start = p.probe ? clock64() : 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int slice = u / (p.n_tiles * p.m_tiles * 3);
      const int kb0 = slice * kb_per_slice, kb1 = min(p.k_blocks, kb0 + kb_per_slice);
      if (p.probe) c0 = clock64();
      mbar_wait(t_empty, (it & 1) ^ 1);
      if (p.probe) w_acc += clock64() - c0;
      tc_fence_after();
      for (int kb = kb0; kb < kb1; ++kb) {
        if (p.probe) c0 = clock64();
        mbar_wait(&full[st], ph);
        if (p.probe) w_full += clock64() - c0;
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + st * kTn9StageBytes), b_base = a_base + kTn9ABytes;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // 16 K rows = two image rows of the patch per instruction
          const uint64_t adesc = umma_desc_sw128_mn(a_base + ks * 2048, 16384, 1024);
          const uint32_t acc = (kb == kb0 && ks == 0) ? 0u : 1u;
          // B: 8-pixel atoms 1280 B apart (halo pitch 10), the three kw taps 128 B apart
          umma_bf16(tmem_base, adesc, umma_desc_sw128_mn(b_base + ks * 2560, 128, 1280), idesc, acc);
          umma_bf16(tmem_base + 192, adesc, umma_desc_sw128_mn(b_base + kTn9BHalf + ks * 2560, 128, 1280), idesc, acc);
        }
        umma_commit(&empty[st]);
        if (++st == kTn9Stages) {
          st = 0;
          ph ^= 1;
        }
      }
      umma_commit(t_full);
    }
    if (p.probe) {
      p.probe[blockIdx.x * 4 + 0] = (float)w_full;
      p.probe[blockIdx.x * 4 + 1] = (float)w_acc;
      p.probe[blockIdx.x * 4 + 2] = (float)(clock64() - c_start);
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int n_tile = u % p.n_tiles, m_tile = (u / p.n_tiles) % p.m_tiles;
      const int kh = (u / (p.n_tiles * p.m_tiles)) % 3, slice = u / (p.n_tiles * p.m_tiles * 3);
      const int kb0 = slice * kb_per_slice, kb1 = min(p.k_blocks, kb0 + kb_per_slice);
      mbar_wait(t_full, it & 1);
      const long long e0 = p.probe ? clock64() : 0;
      tc_fence_after();
      if (kb1 <= kb0 && p.partial) {  // an empty K slice still owns a tile of the fixed-order reduction
        float4* zrow = reinterpret_cast<float4*>(p.partial + ((size_t)u * 128 + q * 32 + lane) * 384);
        for (int i = 0; i < 96; ++i) zrow[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (kb1 > kb0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        float* crow = p.C + (size_t)(m_tile * 128 + q * 32 + lane) * p.ldc + n_tile * 128;
        float4* prow = p.partial ? reinterpret_cast<float4*>(p.partial + ((size_t)u * 128 + q * 32 + lane) * 384) : nullptr;
#pragma unroll 1
        for (int cc = 0; cc < 12; ++cc) {  // column = half*192 + kw*64 + c
          uint32_t v[32];
          tmem_ld_x32(taddr + cc * 32, v);
          tmem_wait_ld();
          if (prow) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              prow[cc * 8 + i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                             __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            continue;
          }
          const int col = cc * 32, half = col / 192, kw = (col % 192) / 64, c0 = col % 64;
          float* dst = crow + (kh * 3 + kw) * p.tap_stride + half * 64 + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dst + i, __uint_as_float(v[i]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty);
      if (p.probe && threadIdx.x == 128) p.probe[blockIdx.x * 4 + 3] = (it == 0 ? 0.f : p.probe[blockIdx.x * 4 + 3]) + (float)(clock64() - e0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// Fixed-order sum of the split-K tiles of gemm_tn9_kernel: C[...] += sum_slice partial[slice][kh][m_tile][n_tile][row][col].
__global__ void __launch_bounds__(256) gemm_tn9_reduce_kernel(const float* __restrict__ partial, int k_split, int m_tiles,
                                                              int n_tiles, float* __restrict__ C, int ldc, int tap_stride) {
  const size_t tile = (size_t)128 * 384;
  const size_t per_slice = (size_t)3 * m_tiles * n_tiles * tile;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_slice) return;
  const int col = (int)(idx % 384), row = (int)((idx / 384) % 128);
  const int un = (int)(idx / tile);  // (kh * m_tiles + m_tile) * n_tiles + n_tile
  const int n_tile = un % n_tiles, m_tile = (un / n_tiles) % m_tiles, kh = un / (n_tiles * m_tiles);
  float acc = 0.f;
#pragma unroll 8
  for (int sl = 0; sl < k_split; ++sl) acc += partial[(size_t)sl * per_slice + idx];  // loads in flight, fixed order
  const int half = col / 192, kw = (col % 192) / 64, c = col % 64;
  float* dst = C + (size_t)(m_tile * 128 + row) * ldc + (kh * 3 + kw) * tap_stride + n_tile * 128 + half * 64 + c;
  *dst += acc;
}

// --------------------------------------------------------------------------
// probe: how fast can TMA refill shared memory from L2?  Each CTA streams
// 16 KB boxes (128 rows x 128 B) of an L2-resident [n_rows][64] bf16 buffer.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
probe_tma_l2_kernel(const __grid_constant__ CUtensorMap map, int n_rows, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int NS = 8;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * 16384);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int n_tiles = n_rows / 128;
    int tile = (blockIdx.x * 37) % n_tiles;
    // keep NS loads in flight; a completed stage is immediately re-armed
    for (int i = 0; i < NS; ++i) {
      mbar_arrive_expect_tx(&full[i], 16384);
      tma_load_2d(smem + i * 16384, &map, &full[i], 0, tile * 128);
      tile = (tile + 1) % n_tiles;
    }
    for (int it = 0; it < iters; ++it) {
      const int st = it % NS;
      mbar_wait(&full[st], (it / NS) & 1);
      mbar_arrive_expect_tx(&full[st], 16384);
      tma_load_2d(smem + st * 16384, &map, &full[st], 0, tile * 128);
      tile = (tile + 1) % n_tiles;
    }
    // drain
    for (int i = 0; i < NS; ++i) {
      const int it = iters + i;
      mbar_wait(&full[it % NS], (it / NS) & 1);
    }
  }
}

}  // namespace cdm
