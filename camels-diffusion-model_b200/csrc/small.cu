// Memory-bound kernels of the ContextUnet / DDPM hot path (everything that is
// not a dense contraction): first/last convolution, EmbedFC, GroupNorm, pooling,
// the sampler tail, the training noise perturbation and the per-sample MSE.
// All of them are single-pass, vectorised (16 B per thread access) and coalesced.
#include <math.h>

#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace cdm {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum (blockDim.x multiple of 32, <= 1024); result broadcast to all threads.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];  // fixed order: deterministic
  return t;
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// ------------------------------------------------------------------ conv_in
// Conv2d(1, C, 3, 1, 1) + per-channel scale/shift (folded BatchNorm) + ReLU, fp32 [n][H][W] -> bf16 NHWC.
// 1152 FMAs per pixel on the CUDA cores made this the slowest memory-bound kernel of the step; it is a
// [pixels x 9] x [9 x C] GEMM, so it runs on mma.sync with the fp32 operands split into bf16 hi + lo parts:
// w' = scale*w = wh + wl, x = xh + xl, K = 29 of 32: xh*wh (9) + xl*wh (9) + xh*wl (9) + 1*shift_hi + 1*shift_lo,
// fp32 accumulate.  The dropped xl*wl term is 2^-16 of a product, i.e. the result keeps fp32-level accuracy
// before the one bf16 rounding of the output, and the whole epilogue is "max(.,0) + round (one F2FP.RELU), store".
// Every WARP is an independent worker (no block-wide barrier after the set-up): it stages the zero-bordered
// (4+2)-row halo of its unit in a warp-private shared-memory buffer (row pitch = 12 mod 32: the three tap rows
// land in disjoint banks), and the NEXT unit's halo is copied into the warp's second buffer (cp.async) under the
// current unit's MMAs, so the only global-load latency a warp ever waits for is its first.  (Round 1 staged per CTA behind a
// __syncthreads: ncu attributed 24 % of the warp-cycles to the wait for that load and 13 % to the barrier, and a
// register spill reload — the 64 weight-fragment registers — another 12 %; the fragments now live in shared
// memory, read as one 16-byte load per two MMAs.)  One unit = 16/parts groups of 16 consecutive pixels x 128
// channels (32 MMAs each) of a 4-row tile; `parts` > 1 spreads small launches over the machine.  The N (channel)
// order of the MMA is permuted so that a thread owns 8 CONSECUTIVE channels per 16-byte store and a warp-level
// store covers 8 pixels x 64 contiguous bytes.  Bound by the 256 B/pixel write.
__device__ __forceinline__ void mma_bf16_m16n8k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bf16_hi(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

constexpr int kCiRows = 4;     // output rows per tile
constexpr int kCiMaxW = 64;
constexpr int kCiPitch = kCiMaxW + 12;                         // image column c <-> index c + 4 (16-byte aligned rows)
constexpr int kCiHalo = (kCiRows + 2) * kCiPitch;              // floats per warp-private halo buffer
constexpr int kCiLoads = (kCiRows + 2) * (kCiMaxW / 4) / 32;   // 16-byte halo chunks per lane (W = 64)
constexpr int kCiCtasPerSm = 3;

template <bool RELU>
__global__ void __launch_bounds__(256, kCiCtasPerSm)
conv_in_mma_kernel(const float* __restrict__ x, int n_img, int H, int W, int C, const float* __restrict__ wgt,
                   const float* __restrict__ scale, const float* __restrict__ shift, bf16* __restrict__ out, int parts) {
  __shared__ uint4 s_bw[4][4][32];     // B fragments [s4][q][lane] = {k-step 0: b0, b1; k-step 1: b0, b1}
  __shared__ __align__(16) float s_x[8][2][kCiHalo];  // warp-private, double-buffered; index 3 <-> image column -1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int slice = blockIdx.y;  // 128-channel slice of the output
  float* sx = s_x[warp][0];
  for (int i = lane; i < 2 * kCiHalo; i += 32) sx[i] = 0.f;
  // this thread's 8 A-fragment slots: slot i <-> k = 16*(i>>2) + 8*((i>>1)&1) + 2t + (i&1);
  // k in [0,9): xh, [9,18): xl, [18,27): xh (times wl), 27/28: the constant 1 (times shift hi / lo), 29..31: 0
  int aoff[8], akind[8];  // shared-memory offset of the tap relative to the pixel; 0 = hi, 1 = lo, 2 = one, 3 = zero
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = 16 * (i >> 2) + 8 * ((i >> 1) & 1) + 2 * t + (i & 1);
    const int tap = k % 9;
    aoff[i] = k < 27 ? (tap / 3) * kCiPitch + tap % 3 + 3 : 0;  // halo row 0 <-> image row y-1, index 3 <-> column -1
    akind[i] = k < 27 ? (k / 9 == 1 ? 1 : 0) : (k < 29 ? 2 : 3);
  }
  // B fragments: n-tile jj = 4*s4 + q (0..15), column g <-> channel 128*slice + 32*s4 + 8*(g>>1) + 2*q + (g&1);
  // warp w prepares n-tiles 2w and 2w+1
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int jj = 2 * warp + u;
    const int ch = 128 * slice + 32 * (jj >> 2) + 8 * (g >> 1) + 2 * (jj & 3) + (g & 1);
    const float sc = __ldg(scale + ch), sh = __ldg(shift + ch);
    uint32_t bw[2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 16 * ks + 8 * r + 2 * t + e;
          v[e] = 0.f;
          if (k < 27) {
            const float w = __ldg(wgt + (k % 9) * C + ch) * sc;
            v[e] = k < 18 ? bf16_hi(w) : w - bf16_hi(w);
          } else if (k == 27) {
            v[e] = bf16_hi(sh);
          } else if (k == 28) {
            v[e] = sh - bf16_hi(sh);
          }
        }
        bw[ks][r] = pack_bf16x2(v[0], v[1]);
      }
    s_bw[jj >> 2][jj & 3][lane] = make_uint4(bw[0][0], bw[0][1], bw[1][0], bw[1][1]);
  }
  __syncthreads();  // the only block-wide barrier: the weight fragments
  const int mtx = W >> 4, tiles_y = H / kCiRows;
  const int n_units = n_img * tiles_y * parts, groups = kCiRows * mtx / parts;
  const int csh = 29 - __clz(W);  // log2(W / 4): 16-byte chunks per row (W = 16 / 32 / 64)
  const int halo_chunks = (kCiRows + 2) << csh;
  const int stride = gridDim.x * 8;
  // the unit's halo rows -> shared memory in 16-byte cp.async chunks (rows outside the image are zero-filled by a
  // zero source size; the border columns keep the zeros written above)
  auto fetch = [&](int unit, float* dst) {
    const int tile = unit / parts, n = tile / tiles_y, y0 = (tile - n * tiles_y) * kCiRows;
    const float* img = x + (size_t)n * H * W;
#pragma unroll
    for (int j = 0; j < kCiLoads; ++j) {
      const int i = lane + 32 * j, rr = i >> csh, cc = (i - (rr << csh)) * 4, yy = y0 - 1 + rr;
      if (i < halo_chunks) {
        const bool in = yy >= 0 && yy < H;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + rr * kCiPitch + cc + 4)),
                     "l"(img + (in ? yy * W + cc : 0)), "r"(in ? 16 : 0)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int unit = blockIdx.x * 8 + warp, buf = 0;
  if (unit < n_units) fetch(unit, sx);
  for (; unit < n_units; unit += stride, buf ^= 1) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();  // this unit's halo is visible to every lane, and every lane has left the previous unit's buffer
    if (unit + stride < n_units) fetch(unit + stride, sx + (buf ^ 1) * kCiHalo);  // in flight under the MMAs below
    const float* sxb = sx + buf * kCiHalo;
    const int tile = unit / parts, part = unit - tile * parts;
    const int n = tile / tiles_y, y0 = (tile - n * tiles_y) * kCiRows;
    for (int tk = part * groups; tk < (part + 1) * groups; ++tk) {
      const int r_loc = tk / mtx, x0 = (tk - r_loc * mtx) * 16 + g;
      const float* px = sxb + r_loc * kCiPitch + x0;
      float av[2][8];  // [pixel row g / g+8][slot]
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float v = px[aoff[i] + 8 * r];
          av[r][i] = akind[i] == 0 ? v : (akind[i] == 1 ? v - bf16_hi(v) : (akind[i] == 2 ? 1.f : 0.f));
        }
      uint32_t a[2][4];  // a0: (g, 2t..), a1: (g+8, 2t..), a2: (g, 2t+8..), a3: (g+8, 2t+8..)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = pack_bf16x2(av[0][4 * ks + 0], av[0][4 * ks + 1]);
        a[ks][1] = pack_bf16x2(av[1][4 * ks + 0], av[1][4 * ks + 1]);
        a[ks][2] = pack_bf16x2(av[0][4 * ks + 2], av[0][4 * ks + 3]);
        a[ks][3] = pack_bf16x2(av[1][4 * ks + 2], av[1][4 * ks + 3]);
      }
      bf16* o0 = out + (((size_t)n * H + y0 + r_loc) * W + x0) * C + 128 * slice + 8 * t;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        float acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 b = s_bw[s4][q][lane];
          acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
          mma_bf16_m16n8k16(acc[q], a[0][0], a[0][1], a[0][2], a[0][3], b.x, b.y);
          mma_bf16_m16n8k16(acc[q], a[1][0], a[1][1], a[1][2], a[1][3], b.z, b.w);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          uint4 v;
          if (RELU) {
            v.x = pack_bf16x2_relu(acc[0][2 * r], acc[0][2 * r + 1]);
            v.y = pack_bf16x2_relu(acc[1][2 * r], acc[1][2 * r + 1]);
            v.z = pack_bf16x2_relu(acc[2][2 * r], acc[2][2 * r + 1]);
            v.w = pack_bf16x2_relu(acc[3][2 * r], acc[3][2 * r + 1]);
          } else {
            v.x = pack_bf16x2(acc[0][2 * r], acc[0][2 * r + 1]);
            v.y = pack_bf16x2(acc[1][2 * r], acc[1][2 * r + 1]);
            v.z = pack_bf16x2(acc[2][2 * r], acc[2][2 * r + 1]);
            v.w = pack_bf16x2(acc[3][2 * r], acc[3][2 * r + 1]);
          }
          *reinterpret_cast<uint4*>(o0 + (size_t)(8 * r) * C + 32 * s4) = v;
        }
      }
    }
  }
}

// ----------------------------------------------------------------- conv_out
// out.1-out.3: GroupNorm(8,128) + ReLU on load, then Conv2d(128, 1, 3, 1, 1).  N = 1 would waste a
// tcgen05 tile, but the nine taps are a perfectly good N: per INPUT pixel p the kernel computes the nine
// dot products d[p][tap] = sum_c relu(gn(x[p][c])) * w[tap][c] as one [pixels x 128] x [128 x 24] bf16
// mma.sync GEMM (weights as bf16 hi + lo: 18 of 24 columns used), parks them in shared memory as nine fp32 planes, and the output is
// the 9-point stencil out(y,x) = bias + sum_{kh,kw} d[(y+kh-1, x+kw-1)][kh*3+kw] (zero outside the image =
// nn.Conv2d's padding, applied after GroupNorm+ReLU).  The activations go global -> registers -> tensor core:
// every element is used once, so there is no shared-memory staging of the 1 MiB/image input.  The K
// (channel) order of an MMA is free as long as A and B agree, so a thread loads 8 CONSECUTIVE channels
// (one 16-byte LDG) of pixel rows g and g+8 and hands them to two k-steps: a warp-level load covers
// 8 pixels x 64 contiguous bytes.  HBM-bound: 1 MiB in + 16 KiB out per image.
constexpr int kCoC = 128;

// relu(a * x + b) on 8 bf16 channels, re-packed to bf16x2 (the one rounding of this layer's input); the ReLU rides
// on the conversion (F2FP.RELU), which takes 8 of the 28 instructions per 16 bytes out of an issue-limited loop.
__device__ __forceinline__ uint4 gn_relu8(const uint4& v, const float4& a0, const float4& a1, const float4& b0,
                                          const float4& b1) {
  uint4 r;
  r.x = pack_bf16x2_relu(fmaf(__uint_as_float(v.x << 16), a0.x, b0.x), fmaf(__uint_as_float(v.x & 0xffff0000u), a0.y, b0.y));
  r.y = pack_bf16x2_relu(fmaf(__uint_as_float(v.y << 16), a0.z, b0.z), fmaf(__uint_as_float(v.y & 0xffff0000u), a0.w, b0.w));
  r.z = pack_bf16x2_relu(fmaf(__uint_as_float(v.z << 16), a1.x, b1.x), fmaf(__uint_as_float(v.z & 0xffff0000u), a1.y, b1.y));
  r.w = pack_bf16x2_relu(fmaf(__uint_as_float(v.w << 16), a1.z, b1.z), fmaf(__uint_as_float(v.w & 0xffff0000u), a1.w, b1.w));
  return r;
}

// Plane stride (floats) of the d[tap] planes: (R+2) x (W+2) padded so that stride % 16 == 4, which makes the
// accumulator stores of a warp (4 planes x 8 consecutive pixels) hit 32 different banks.
static __host__ __device__ inline int co_plane_stride(int R, int W) {
  int ps = (R + 2) * (W + 2);
  while ((ps & 15) != 4) ++ps;
  return ps;
}

template <int R>
__global__ void __launch_bounds__(256, R <= 16 ? 3 : 2) conv_out_mma_kernel(const bf16* __restrict__ src, int n_img, int H, int W,
                                                              const float* __restrict__ mean_rstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ wgt,
                                                              const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ __align__(16) float co_smem[];
  const int PW = W + 2, PS = co_plane_stride(R, W);
  float* s_d = co_smem;          // [9][PS]: plane row 0 <-> image row ty*R-1, column 0 <-> image column -1
  float* s_a = s_d + 9 * PS;     // [128] rstd * gamma
  float* s_b = s_a + kCoC;       // [128] beta - mean * rstd * gamma
  const int tiles_y = H / R;
  const int n = blockIdx.x / tiles_y, ty = blockIdx.x % tiles_y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  if (threadIdx.x < kCoC) {
    const int c = threadIdx.x, grp = c >> 4;
    const float mean = __ldg(mean_rstd + ((size_t)n * 8 + grp) * 2), rstd = __ldg(mean_rstd + ((size_t)n * 8 + grp) * 2 + 1);
    const float a = rstd * __ldg(gamma + c);
    s_a[c] = a;
    s_b[c] = __ldg(beta + c) - mean * a;
  }
  for (int i = threadIdx.x; i < 9 * (R + 2) * 2; i += blockDim.x) {  // the two x-halo columns of every plane row
    const int plane = i / ((R + 2) * 2), rem = i - plane * (R + 2) * 2;
    s_d[plane * PS + (rem >> 1) * PW + ((rem & 1) ? W + 1 : 0)] = 0.f;
  }
  // B fragments (weights) in the permuted channel order: k-step 2q+h, fragment columns {2t,2t+1} and {2t+8,2t+9}
  // <-> channels 32q + 8t + 4h + {0,1} and + {2,3}.  The fp32 weights are split into bf16 hi + lo parts so that the
  // only rounding of this layer is the activation's: n-tile 0 = hi of taps 0..7, n-tile 1 = lo of taps 0..7 (both
  // accumulate into the same registers), n-tile 2 = {hi, lo} of tap 8 in columns 0 and 1.  Kept in shared memory
  // ([k-step][n-tile][lane], 8-byte entries: conflict-free) to leave the registers to the activation loads.
  uint2* s_bw = reinterpret_cast<uint2*>(s_b + kCoC);
  if (warp == 0) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int ch = 32 * (ks >> 1) + 8 * t + 4 * (ks & 1);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wgt + g * kCoC + ch));
      const float h0 = bf16_hi(w0.x), h1 = bf16_hi(w0.y), h2 = bf16_hi(w0.z), h3 = bf16_hi(w0.w);
      s_bw[(ks * 3 + 0) * 32 + lane] = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
      s_bw[(ks * 3 + 1) * 32 + lane] = make_uint2(pack_bf16x2(w0.x - h0, w0.y - h1), pack_bf16x2(w0.z - h2, w0.w - h3));
      float4 w8 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < 2) {
        w8 = __ldg(reinterpret_cast<const float4*>(wgt + 8 * kCoC + ch));
        const float k0 = bf16_hi(w8.x), k1 = bf16_hi(w8.y), k2 = bf16_hi(w8.z), k3 = bf16_hi(w8.w);
        w8 = g == 0 ? make_float4(k0, k1, k2, k3) : make_float4(w8.x - k0, w8.y - k1, w8.z - k2, w8.w - k3);
      }
      s_bw[(ks * 3 + 2) * 32 + lane] = make_uint2(pack_bf16x2(w8.x, w8.y), pack_bf16x2(w8.z, w8.w));
    }
  }
  __syncthreads();
  const int mtx = W >> 4, n_mt = (R + 2) * mtx;
  const float4* sa4 = reinterpret_cast<const float4*>(s_a);
  const float4* sb4 = reinterpret_cast<const float4*>(s_b);
  for (int mt = warp; mt < n_mt; mt += 8) {
    const int pr = mt / mtx, mx = mt - pr * mtx;
    const int ih = ty * R - 1 + pr;
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    if (ih >= 0 && ih < H) {  // warp-uniform; rows outside the image contribute zeros (the conv padding)
      const bf16* p0 = src + ((((size_t)n * H + ih) * W + mx * 16 + g) * kCoC + 8 * t);
      uint4 v0[4], v1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v0[q] = __ldg(reinterpret_cast<const uint4*>(p0 + 32 * q));
        v1[q] = __ldg(reinterpret_cast<const uint4*>(p0 + 8 * kCoC + 32 * q));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a0 = sa4[8 * q + 2 * t], a1 = sa4[8 * q + 2 * t + 1];
        const float4 b0 = sb4[8 * q + 2 * t], b1 = sb4[8 * q + 2 * t + 1];
        const uint4 r0 = gn_relu8(v0[q], a0, a1, b0, b1), r1 = gn_relu8(v1[q], a0, a1, b0, b1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t x0 = h ? r0.z : r0.x, x1 = h ? r1.z : r1.x, x2 = h ? r0.w : r0.y, x3 = h ? r1.w : r1.y;
          const uint2 bh = s_bw[((2 * q + h) * 3 + 0) * 32 + lane], bl = s_bw[((2 * q + h) * 3 + 1) * 32 + lane];
          const uint2 b8 = s_bw[((2 * q + h) * 3 + 2) * 32 + lane];
          mma_bf16_m16n8k16(acc0, x0, x1, x2, x3, bh.x, bh.y);
          mma_bf16_m16n8k16(acc0, x0, x1, x2, x3, bl.x, bl.y);
          mma_bf16_m16n8k16(acc1, x0, x1, x2, x3, b8.x, b8.y);
        }
      }
    }
    float* d0 = s_d + pr * PW + mx * 16 + g + 1;
    d0[(2 * t) * PS] = acc0[0];
    d0[(2 * t + 1) * PS] = acc0[1];
    d0[(2 * t) * PS + 8] = acc0[2];
    d0[(2 * t + 1) * PS + 8] = acc0[3];
    if (t == 0) {  // columns 0 and 1 of n-tile 2: hi and lo part of tap 8
      d0[8 * PS] = acc1[0] + acc1[1];
      d0[8 * PS + 8] = acc1[2] + acc1[3];
    }
  }
  __syncthreads();
  const float bias0 = __ldg(bias);
  for (int o = threadIdx.x; o < R * W; o += blockDim.x) {
    const int r = o / W, x = o - r * W;
    const float* d = s_d + r * PW + x;
    float s[3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
      s[kh] = (d[(kh * 3) * PS + kh * PW] + d[(kh * 3 + 1) * PS + kh * PW + 1]) + d[(kh * 3 + 2) * PS + kh * PW + 2];
    out[((size_t)n * H + ty * R + r) * W + x] = ((s[0] + s[1]) + s[2]) + bias0;
  }
}

// ------------------------------------------------------------------ embed_fc
// EmbedFC: out = W2 * gelu(W1 * v + b1) + b2, one block per input row.
__global__ void __launch_bounds__(256) embed_fc_kernel(const float* __restrict__ in, int din,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       int emb, float* __restrict__ out) {
  extern __shared__ float s_h[];  // [emb]
  const int row = blockIdx.x;
  for (int j = threadIdx.x; j < emb; j += blockDim.x) {
    float s = b1[j];
    for (int i = 0; i < din; ++i) s = fmaf(w1[j * din + i], in[(size_t)row * din + i], s);
    s_h[j] = gelu_erf(s);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < emb; j += nw) {
    float s = 0.f;
    for (int k = lane; k < emb; k += 32) s = fmaf(w2[(size_t)j * emb + k], s_h[k], s);
    s = warp_sum(s);
    if (lane == 0) out[(size_t)row * emb + j] = s + b2[j];
  }
}

// ------------------------------------------------------------ avgpool + gelu
// to_vec: AvgPool2d over all P pixels + GELU; src bf16 [n][P][C] -> out bf16 [n][C].
__global__ void __launch_bounds__(256) avgpool_gelu_kernel(const bf16* __restrict__ src, int P, int C,
                                                           bf16* __restrict__ out) {
  const size_t n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += __bfloat162float(src[(n * P + p) * C + c]);
    out[n * C + c] = __float2bfloat16(gelu_erf(s / (float)P));
  }
}

// Vectorised form for C % 8 == 0 and C / 8 a divisor of 256 (C = 128, 256): a thread owns 8 consecutive channels
// (one 16-byte load per pixel) of the pixels p = j (mod 32) for JPT of the 32 residues j; each residue's pixels are
// added in increasing order into ITS OWN partial sum, and the 32 partials are then added in the order j = 0..31
// through shared memory.  The arithmetic is therefore the same for every JPT: the launch picks 32 / JPT * (C / 8)
// threads per image — 256 threads (JPT = 4 at C = 256, four loads in flight) when there are images enough to fill the
// machine, 1024 (JPT = 1, all of a thread's pixels in flight at once) for small batches, where one block per image
// walking 128 KB in eight dependent round trips took 21 us of the 245 us batch-1 sampling step.
template <int JPT>
__global__ void __launch_bounds__(1024 / JPT) avgpool_gelu_vec_kernel(const bf16* __restrict__ src, int P, int C,
                                                                      bf16* __restrict__ out) {
  extern __shared__ float av_red[];  // [32][C]
  const size_t n = blockIdx.x;
  const int vecs = C >> 3, rows = 32 / JPT;
  const int v = threadIdx.x % vecs, r = threadIdx.x / vecs;
  const bf16* base = src + n * P * C + v * 8;
  float s[JPT][8];
#pragma unroll
  for (int u = 0; u < JPT; ++u)
#pragma unroll
    for (int j = 0; j < 8; ++j) s[u][j] = 0.f;
  constexpr int kDepth = JPT == 1 ? 8 : 1;  // pixels of one residue in flight together
  for (int p0 = r; p0 < P; p0 += 32 * kDepth) {
    uint4 raw[kDepth][JPT];
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
#pragma unroll
      for (int u = 0; u < JPT; ++u) {
        const int p = p0 + 32 * d + u * rows;
        raw[d][u] = p < P ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)p * C)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
#pragma unroll
      for (int u = 0; u < JPT; ++u) {
        float f[8];
        unpack8(raw[d][u], f);
        if (p0 + 32 * d + u * rows < P) {  // (an absent pixel must not even add +0: -0 + +0 would flip a sign bit)
#pragma unroll
          for (int j = 0; j < 8; ++j) s[u][j] += f[j];
        }
      }
  }
#pragma unroll
  for (int u = 0; u < JPT; ++u)
#pragma unroll
    for (int j = 0; j < 8; ++j) av_red[((r + u * rows) * vecs + v) * 8 + j] = s[u][j];
  __syncthreads();
  if (r == 0) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
      for (int q = 0; q < 32; ++q) t += av_red[(q * vecs + v) * 8 + j];
      f[j] = gelu_erf(t / (float)P);
    }
    *reinterpret_cast<uint4*>(out + n * C + v * 8) = pack8(f);
  }
}

// ----------------------------------------------- GroupNorm + ReLU + FiLM (up0)
// One block per (image, group): statistics over P pixels x cpg channels, then
// y = film_scale * relu(gn(x)) + film_shift.  src/out bf16 [n][P][C].
// VPT > 0: the group is exactly VPT 16-byte vectors per thread (up0: 256 px x 32 ch = 4 x 256 vectors) and is
// read from global memory ONCE, all VPT loads in flight together, and kept in registers for the two statistics
// passes and the apply pass.  VPT == 0: any size, three passes over global memory (the later ones hit L1/L2).
// Both walk a thread's elements in the same order, so their results are bit-identical.
template <int VPT>
__global__ void __launch_bounds__(256, 3) gn_relu_film_kernel(const bf16* __restrict__ src, int P, int C, int groups,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps,
                                                           const float* __restrict__ film_scale,
                                                           const float* __restrict__ film_shift, int film_rows,
                                                           const int* __restrict__ step_ptr, bf16* __restrict__ out,
                                                           float* __restrict__ mean_rstd_out) {
  __shared__ float red[32];
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = C / groups;      // multiple of 8
  const int vec_per_px = cpg / 8;  // uint4 per pixel
  const int n_vec = P * vec_per_px;
  const bf16* base = src + (size_t)n * P * C + g * cpg;
  constexpr int kHeld = VPT > 0 ? VPT : 1;
  uint4 held[kHeld];
  if (VPT > 0) {
#pragma unroll
    for (int k = 0; k < kHeld; ++k) {
      const int i = threadIdx.x + k * 256;
      held[k] = *reinterpret_cast<const uint4*>(base + (size_t)(i / vec_per_px) * C + (i % vec_per_px) * 8);
    }
  }
  float s = 0.f;
  if (VPT > 0) {
#pragma unroll
    for (int k = 0; k < kHeld; ++k) {
      float f[8];
      unpack8(held[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
  } else {
    for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
      const int p = i / vec_per_px, v = i % vec_per_px;
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(base + (size_t)p * C + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
  }
  const float cnt = (float)(P * cpg);
  const float mean = block_sum(s, red) / cnt;
  float q = 0.f;
  if (VPT > 0) {
#pragma unroll
    for (int k = 0; k < kHeld; ++k) {
      float f[8];
      unpack8(held[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) q = fmaf(f[j] - mean, f[j] - mean, q);
    }
  } else {
    for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
      const int p = i / vec_per_px, v = i % vec_per_px;
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(base + (size_t)p * C + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) q = fmaf(f[j] - mean, f[j] - mean, q);
    }
  }
  const float rstd = rsqrtf(block_sum(q, red) / cnt + eps);
  if (mean_rstd_out && threadIdx.x == 0) {
    mean_rstd_out[((size_t)n * groups + g) * 2] = mean;
    mean_rstd_out[((size_t)n * groups + g) * 2 + 1] = rstd;
  }
  const int step = step_ptr ? *step_ptr : 0;
  const float* fs = film_scale ? film_scale + (size_t)n * C + g * cpg : nullptr;
  const float* fb =
      film_shift ? film_shift + ((size_t)step * film_rows + (film_rows == 1 ? 0 : n)) * C + g * cpg : nullptr;
  auto norm1 = [&](float x, int c_in_group) {  // one element: GroupNorm affine, ReLU, FiLM
    const int c = g * cpg + c_in_group;
    float y = fmaxf(fmaf((x - mean) * rstd, gamma[c], beta[c]), 0.f);
    if (fs) y = fmaf(fs[c_in_group], y, fb[c_in_group]);
    return y;
  };
  if (VPT > 0) {
    // the launch guarantees 256 % vec_per_px == 0: all of a thread's vectors cover the same 8 channels, so the
    // coefficients of a channel pair are fetched once and applied to the VPT held vectors (in place)
    const int v = threadIdx.x % vec_per_px;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int c0 = v * 8 + w * 2;
      const int c = g * cpg + c0;
      const float g0 = gamma[c], g1 = gamma[c + 1], b0 = beta[c], b1 = beta[c + 1];
      const float s0 = fs ? fs[c0] : 0.f, s1 = fs ? fs[c0 + 1] : 0.f;
      const float t0 = fs ? fb[c0] : 0.f, t1 = fs ? fb[c0 + 1] : 0.f;
#pragma unroll
      for (int k = 0; k < kHeld; ++k) {
        uint32_t& word = w == 0 ? held[k].x : w == 1 ? held[k].y : w == 2 ? held[k].z : held[k].w;
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&word);
        float y0 = fmaxf(fmaf((__low2float(h) - mean) * rstd, g0, b0), 0.f);
        float y1 = fmaxf(fmaf((__high2float(h) - mean) * rstd, g1, b1), 0.f);
        if (fs) {
          y0 = fmaf(s0, y0, t0);
          y1 = fmaf(s1, y1, t1);
        }
        word = pack_bf16x2(y0, y1);
      }
    }
#pragma unroll
    for (int k = 0; k < kHeld; ++k) {
      const int i = threadIdx.x + k * 256;
      *reinterpret_cast<uint4*>(out + (size_t)n * P * C + (size_t)(i / vec_per_px) * C + g * cpg + v * 8) = held[k];
    }
  } else {
    for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
      const int p = i / vec_per_px, v = i % vec_per_px;
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(base + (size_t)p * C + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = norm1(f[j], v * 8 + j);
      *reinterpret_cast<uint4*>(out + (size_t)n * P * C + (size_t)p * C + g * cpg + v * 8) = pack8(f);
    }
  }
}

// ------------------------------------------------------------- gn_finalize
// Reduce the per-tile partial sums the out.0 convolution emitted, in a fixed order
// (deterministic), into mean / rstd per (image, group).  One warp per (image, group).
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int n_img, int slots, float count, float eps,
                                   float* __restrict__ mean_rstd) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_img * 8) return;
  const int n = wid >> 3, g = wid & 7;
  float s = 0.f, q = 0.f;
  for (int k = lane; k < slots; k += 32) {
    const float* p = partial + (((size_t)n * slots + k) * 8 + g) * 2;
    s += p[0];
    q += p[1];
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    const float mean = s / count;
    const float var = fmaxf(q / count - mean * mean, 0.f);
    mean_rstd[(size_t)wid * 2] = mean;
    mean_rstd[(size_t)wid * 2 + 1] = rsqrtf(var + eps);
  }
}

// ---------------------------------------------------------------- RNG (Philox)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// idx = GLOBAL index of the 4-element vector (sample_offset * hw / 4 + local index): the draw of a sample does not
// depend on which rank holds it or where it sits in the rank's shard
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long idx, uint32_t stream) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, stream, 0x1234567u, (uint32_t)(idx >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 5.9604644775390625e-8f;  // 2^-24
  const float u0 = ((r.x >> 8) + 0.5f) * k, u1 = ((r.y >> 8) + 0.5f) * k;
  const float u2 = ((r.z >> 8) + 0.5f) * k, u3 = ((r.w >> 8) + 0.5f) * k;
  const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
  float s0, c0, s1, c1;
  sincosf(6.283185307179586f * u1, &s0, &c0);
  sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// --------------------------------------------------------------- ddpm step
// eps = eps_u + w (eps_c - eps_u)  (iff reps == 2);  x <- (x - eps*k2)/sqrt(a) + sqrt(b) z.
// Every fp32 operation is rounded separately (no FMA contraction, IEEE division) so the
// trajectory matches the reference's chain of torch elementwise ops bit for bit.
struct DdpmKParams {
  float* x;
  const float* eps;
  int n, hw, reps;
  float guide_w;
  const float* coef;
  const int* step_ptr;
  int step, timesteps;
  const float* z;
  long long z_iter_stride;
  unsigned long long seed;
  float* snap;
  const int* snap_slot;
  unsigned long long vec_offset;  // sample_offset * hw / 4
};
__global__ void __launch_bounds__(256) ddpm_step_kernel(const DdpmKParams p) {
  const int i = p.step_ptr ? *p.step_ptr : p.step;
  const float k2 = p.coef[i * 4 + 0], sa = p.coef[i * 4 + 1], sb = p.coef[i * 4 + 2];
  const size_t total4 = (size_t)p.n * p.hw / 4;
  const size_t half = (size_t)p.n * p.hw;
  const float* zbase = p.z ? p.z + (size_t)(p.timesteps - i) * p.z_iter_stride : nullptr;
  const int slot = p.snap_slot ? p.snap_slot[i] : -1;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total4; v += (size_t)gridDim.x * blockDim.x) {
    float4 x4 = reinterpret_cast<float4*>(p.x)[v];
    float4 e4 = reinterpret_cast<const float4*>(p.eps)[v];
    if (p.reps == 2 && p.guide_w > 0.f) {
      const float4 u4 = reinterpret_cast<const float4*>(p.eps + half)[v];
      e4.x = __fadd_rn(u4.x, __fmul_rn(p.guide_w, __fsub_rn(e4.x, u4.x)));
      e4.y = __fadd_rn(u4.y, __fmul_rn(p.guide_w, __fsub_rn(e4.y, u4.y)));
      e4.z = __fadd_rn(u4.z, __fmul_rn(p.guide_w, __fsub_rn(e4.z, u4.z)));
      e4.w = __fadd_rn(u4.w, __fmul_rn(p.guide_w, __fsub_rn(e4.w, u4.w)));
    }
    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i > 1) z4 = zbase ? reinterpret_cast<const float4*>(zbase)[v] : philox_normal4(p.seed, p.vec_offset + v, (uint32_t)i);
    x4.x = __fadd_rn(__fdiv_rn(__fsub_rn(x4.x, __fmul_rn(e4.x, k2)), sa), __fmul_rn(sb, z4.x));
    x4.y = __fadd_rn(__fdiv_rn(__fsub_rn(x4.y, __fmul_rn(e4.y, k2)), sa), __fmul_rn(sb, z4.y));
    x4.z = __fadd_rn(__fdiv_rn(__fsub_rn(x4.z, __fmul_rn(e4.z, k2)), sa), __fmul_rn(sb, z4.z));
    x4.w = __fadd_rn(__fdiv_rn(__fsub_rn(x4.w, __fmul_rn(e4.w, k2)), sa), __fmul_rn(sb, z4.w));
    reinterpret_cast<float4*>(p.x)[v] = x4;
    if (slot >= 0) reinterpret_cast<float4*>(p.snap + (size_t)slot * half)[v] = x4;
  }
}
__global__ void step_advance_kernel(int* step_ptr, int delta) { *step_ptr += delta; }

// ----------------------------------------------------------------- perturb
// x_t = ca[t] * x + cb[t] * noise (ca = sqrt(ab_t); cb = 1 - ab_t in the training / NLL form,
// sqrt(1 - ab_t) in the dataloader-ELBO form).  t per sample (t_idx) or shared (t_shared).
__global__ void __launch_bounds__(256) perturb_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                      float* __restrict__ out, int n, int hw,
                                                      const float* __restrict__ ca, const float* __restrict__ cb,
                                                      const long long* __restrict__ t_idx, int t_shared,
                                                      const int* __restrict__ step_ptr, unsigned long long seed,
                                                      uint32_t stream, float* __restrict__ noise_out,
                                                      unsigned long long vec_offset) {
  const int hw4 = hw / 4;
  if (step_ptr) {
    t_shared = *step_ptr;
    stream += (uint32_t)t_shared;
  }
  const size_t total4 = (size_t)n * hw4;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total4; v += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(v / hw4);
    const int t = t_idx ? (int)t_idx[s] : t_shared;
    const float a = ca[t], b = cb[t];
    const float4 x4 = reinterpret_cast<const float4*>(x)[v];
    float4 n4;
    if (noise) {
      n4 = reinterpret_cast<const float4*>(noise)[v];
    } else {
      n4 = philox_normal4(seed, vec_offset + v, stream);
      reinterpret_cast<float4*>(noise_out)[v] = n4;
    }
    float4 o;
    o.x = __fadd_rn(__fmul_rn(a, x4.x), __fmul_rn(b, n4.x));
    o.y = __fadd_rn(__fmul_rn(a, x4.y), __fmul_rn(b, n4.y));
    o.z = __fadd_rn(__fmul_rn(a, x4.z), __fmul_rn(b, n4.z));
    o.w = __fadd_rn(__fmul_rn(a, x4.w), __fmul_rn(b, n4.w));
    reinterpret_cast<float4*>(out)[v] = o;
  }
}

// --------------------------------------------------------------- mse_accum
// mse[s] = mean((pred - target)^2) per sample; optionally acc[s] += weight[t] * mse[s].
__global__ void __launch_bounds__(256) mse_accum_kernel(const float* __restrict__ pred,
                                                        const float* __restrict__ target, int hw,
                                                        const float* __restrict__ weight_tab,
                                                        const long long* __restrict__ t_idx, int t_shared,
                                                        const int* __restrict__ step_ptr,
                                                        float* __restrict__ mse_out, float* __restrict__ acc,
                                                        const float* __restrict__ weight_tab2,
                                                        float* __restrict__ acc2) {
  __shared__ float red[32];
  if (step_ptr) t_shared = *step_ptr;
  const size_t s = blockIdx.x;
  const float4* p4 = reinterpret_cast<const float4*>(pred + s * hw);
  const float4* t4 = reinterpret_cast<const float4*>(target + s * hw);
  float q = 0.f;
  for (int v = threadIdx.x; v < hw / 4; v += blockDim.x) {
    const float4 a = p4[v], b = t4[v];
    const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
    q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  const float tot = block_sum(q, red);
  if (threadIdx.x == 0) {
    const float mse = tot / (float)hw;
    if (mse_out) mse_out[s] = mse;
    if (acc) {
      const int t = t_idx ? (int)t_idx[s] : t_shared;
      acc[s] += (weight_tab ? weight_tab[t] : 1.f) * mse;
    }
    if (acc2) {  // second accumulator of the same sweep (NLL and ELBO weights over one set of forwards)
      const int t = t_idx ? (int)t_idx[s] : t_shared;
      acc2[s] += (weight_tab2 ? weight_tab2[t] : 1.f) * mse;
    }
  }
}

static int grid_for(size_t work_items, int block) {
  size_t g = (work_items + block - 1) / block;
  const size_t cap = num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace cdm

using namespace cdm;

extern "C" int cdm_conv_in(const cdm_conv_in_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->x && a->weight && a->scale && a->shift && a->out);
  CDM_CHECK_ARG(a->n_img > 0 && a->H > 0 && a->H % kCiRows == 0 && (a->W == 16 || a->W == 32 || a->W == kCiMaxW));
  CDM_CHECK_ARG(((uintptr_t)a->x & 15) == 0);  // the halo rows are copied in 16-byte chunks
  CDM_CHECK_ARG(a->cout > 0 && a->cout % 128 == 0 && a->cout <= 512);
  int rc = check_device();
  if (rc) return rc;
  // Work unit = 1/parts of a 4-row tile, one warp each.  Large launches take whole tiles (the halo is staged once per
  // tile); `parts` doubles while the units would not give every resident warp about six turns (tail quantisation)
  // and for small launches (batch-1 sampling: 16 tiles) until the units spread over the machine.
  const int tiles = a->n_img * (a->H / kCiRows), groups = kCiRows * (a->W / 16);
  const int warps = num_sms() * kCiCtasPerSm * 8;
  int parts = 1;
  while (groups % (parts * 2) == 0 && (long long)tiles * parts < 6LL * warps) parts *= 2;
  const int units = tiles * parts, ctas = (units + 7) / 8;
  const dim3 grid(ctas < num_sms() * kCiCtasPerSm ? ctas : num_sms() * kCiCtasPerSm, a->cout / 128);
  if (a->relu)
    conv_in_mma_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a->x, a->n_img, a->H, a->W, a->cout, a->weight,
                                                                     a->scale, a->shift, (bf16*)a->out, parts);
  else
    conv_in_mma_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a->x, a->n_img, a->H, a->W, a->cout, a->weight,
                                                                      a->scale, a->shift, (bf16*)a->out, parts);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

template <int R>
static int launch_conv_out(const cdm_conv_out_args* a, cudaStream_t stream) {
  static bool attr_set[64];  // per device: cudaFuncSetAttribute is not process-wide
  const int smem = (9 * co_plane_stride(R, a->W) + 2 * kCoC) * (int)sizeof(float) + 8 * 3 * 32 * 8;
  int dev = 0;
  CDM_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    CDM_CHECK_CUDA(cudaFuncSetAttribute(conv_out_mma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  conv_out_mma_kernel<R><<<a->n_img * (a->H / R), 256, smem, stream>>>(
      (const bf16*)a->src, a->n_img, a->H, a->W, a->mean_rstd, a->gamma, a->beta, a->weight, a->bias, a->out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_conv_out(const cdm_conv_out_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->src && a->mean_rstd && a->gamma && a->beta && a->weight && a->bias && a->out);
  CDM_CHECK_ARG(a->n_img > 0 && a->H > 0 && a->W > 0 && a->W % 16 == 0 && a->W <= 64 && a->H % 8 == 0 && a->C == kCoC);
  int rc = check_device();
  if (rc) return rc;
  // 32-row tiles (6 % halo re-read) once they fill the machine twice over, 8-row tiles for small batches
  // (16-row tiles at three blocks per SM were measured too: 0.475 vs 0.484 ms at 2048 images, within noise — the pass
  // is paced by the DRAM pipe, not by occupancy)
  if (a->H % 32 == 0 && a->n_img * (a->H / 32) >= 2 * num_sms()) return launch_conv_out<32>(a, (cudaStream_t)stream);
  return launch_conv_out<8>(a, (cudaStream_t)stream);
}

extern "C" int cdm_embed_fc(const float* in, int rows, int din, const float* w1, const float* b1, const float* w2,
                            const float* b2, int emb, float* out, void* stream) {
  CDM_CHECK_ARG(in && w1 && b1 && w2 && b2 && out && rows > 0 && din > 0 && emb > 0 && emb <= 4096);
  int rc = check_device();
  if (rc) return rc;
  embed_fc_kernel<<<rows, 256, emb * sizeof(float), (cudaStream_t)stream>>>(in, din, w1, b1, w2, b2, emb, out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_avgpool_gelu(const void* src, int n_img, int P, int C, void* out, void* stream) {
  CDM_CHECK_ARG(src && out && n_img > 0 && P > 0 && C > 0);
  int rc = check_device();
  if (rc) return rc;
  if (C % 8 == 0 && C / 8 <= 32 && 32 % (C / 8) == 0) {  // C = 8 .. 256: 32 residues x C / 8 vectors <= 1024 threads
    const int vecs = C / 8;
    const size_t smem = (size_t)32 * C * sizeof(float);
    if (n_img > 2 * num_sms() && (8 * vecs) % 32 == 0)
      avgpool_gelu_vec_kernel<4><<<n_img, 8 * vecs, smem, (cudaStream_t)stream>>>((const bf16*)src, P, C, (bf16*)out);
    else
      avgpool_gelu_vec_kernel<1><<<n_img, 32 * vecs, smem, (cudaStream_t)stream>>>((const bf16*)src, P, C, (bf16*)out);
  }
  else
    avgpool_gelu_kernel<<<n_img, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, P, C, (bf16*)out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_gn_relu_film(const cdm_gn_relu_film_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->src && a->gamma && a->beta && a->out);
  CDM_CHECK_ARG(a->n_img > 0 && a->P > 0 && a->groups > 0 && a->C % a->groups == 0 && (a->C / a->groups) % 8 == 0);
  CDM_CHECK_ARG((a->film_scale == nullptr) == (a->film_shift == nullptr));
  CDM_CHECK_ARG(!a->film_scale || a->film_rows >= 1);
  int rc = check_device();
  if (rc) return rc;
  const int n_vec = a->P * (a->C / a->groups / 8);
  auto kern = (n_vec == 4 * 256 && 256 % (a->C / a->groups / 8) == 0) ? gn_relu_film_kernel<4> : gn_relu_film_kernel<0>;  // up0: 256 px x 32 ch
  kern<<<a->n_img * a->groups, 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)a->src, a->P, a->C, a->groups, a->gamma, a->beta, a->eps, a->film_scale, a->film_shift,
      a->film_rows, a->step_ptr, (bf16*)a->out, a->mean_rstd_out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_gn_finalize(const float* partial, int n_img, int slots, float count, float eps, float* mean_rstd,
                               void* stream) {
  CDM_CHECK_ARG(partial && mean_rstd && n_img > 0 && slots > 0 && count > 0);
  int rc = check_device();
  if (rc) return rc;
  const int warps = n_img * 8;
  gn_finalize_kernel<<<(warps * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partial, n_img, slots, count, eps,
                                                                                 mean_rstd);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_ddpm_step(const cdm_ddpm_step_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->x && a->eps && a->coef);
  CDM_CHECK_ARG(a->n > 0 && a->hw > 0 && a->hw % 4 == 0 && (a->reps == 1 || a->reps == 2) && a->timesteps > 0);
  CDM_CHECK_ARG(a->step_ptr || (a->step >= 1 && a->step <= a->timesteps));
  CDM_CHECK_ARG((a->snap == nullptr) == (a->snap_slot == nullptr));
  int rc = check_device();
  if (rc) return rc;
  DdpmKParams p;
  p.x = a->x;
  p.eps = a->eps;
  p.n = a->n;
  p.hw = a->hw;
  p.reps = a->reps;
  p.guide_w = a->guide_w;
  p.coef = a->coef;
  p.step_ptr = a->step_ptr;
  p.step = a->step;
  p.timesteps = a->timesteps;
  p.z = a->z;
  p.z_iter_stride = a->z_iter_stride;
  p.seed = a->seed;
  p.snap = a->snap;
  p.snap_slot = a->snap_slot;
  CDM_CHECK_ARG(a->sample_offset >= 0);
  p.vec_offset = (unsigned long long)a->sample_offset * (unsigned long long)(a->hw / 4);
  ddpm_step_kernel<<<grid_for((size_t)a->n * a->hw / 4, 256), 256, 0, (cudaStream_t)stream>>>(p);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_step_advance(int* step_ptr, int delta, void* stream) {
  CDM_CHECK_ARG(step_ptr);
  int rc = check_device();
  if (rc) return rc;
  step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_ptr, delta);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_perturb(const cdm_perturb_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->x && a->out && a->ca && a->cb && a->n > 0 && a->hw > 0 && a->hw % 4 == 0);
  CDM_CHECK_ARG(a->noise || a->noise_out);
  CDM_CHECK_ARG(a->sample_offset >= 0);
  int rc = check_device();
  if (rc) return rc;
  perturb_kernel<<<grid_for((size_t)a->n * a->hw / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      a->x, a->noise, a->out, a->n, a->hw, a->ca, a->cb, a->t_idx, a->t_shared, a->step_ptr, a->seed, a->stream_id,
      a->noise_out, (unsigned long long)a->sample_offset * (unsigned long long)(a->hw / 4));
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_mse_accum(const cdm_mse_accum_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->pred && a->target && a->n > 0 && a->hw > 0 && a->hw % 4 == 0 && (a->mse_out || a->acc));
  int rc = check_device();
  if (rc) return rc;
  mse_accum_kernel<<<a->n, 256, 0, (cudaStream_t)stream>>>(a->pred, a->target, a->hw, a->weight_tab, a->t_idx,
                                                           a->t_shared, a->step_ptr, a->mse_out, a->acc,
                                                           a->weight_tab2, a->acc2);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
