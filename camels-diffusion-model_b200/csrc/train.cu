// Training-path kernels that are not dense contractions: train-mode BatchNorm (statistics,
// normalise, backward), MaxPool forward/backward, FiLM / GroupNorm / to_vec / EmbedFC backward,
// the K=9 / N=1 convolutions' weight gradients, space-to-depth for the transposed-conv backward,
// the MSE loss gradient and a fused multi-tensor Adam.  Activations and their gradients are
// NHWC bf16 (pixel stride `ld`, so channel slices of wider tensors can be used in place);
// every statistic, reduction and parameter gradient is fp32.  Reductions are two-stage with a
// fixed order (deterministic); nothing here uses atomics on global memory.
#include <math.h>

#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace cdm {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float t_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float t_block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = t_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ void t_unpack8(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16);
  f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16);
  f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 t_pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}
__device__ __forceinline__ uint4 ld8(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}

// --------------------------------------------------------------------------
// pack_transpose: dst[b][c][r] (bf16) = src[b][r][c] (fp32), batched, both innermost dims contiguous, through a 64 x 64
// shared-memory tile: the two bf16 layouts of up0.0.weight (78 % of all parameters) are such transposes
// ([ci][co][khw] -> [ci][khw][co] and -> [khw][co][ci]).  In the table-driven gather above their innermost output
// dimension walks the source with a stride of 256 / 65536 floats: one 32-byte sector per 4-byte element, 0.185 ms per
// step; as tiled transposes they read and write whole lines.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_transpose_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                             int R, int Cc, long long sb, long long sr, long long db,
                                                             long long dc) {
  __shared__ float tile[64][65];
  const int tiles_c = Cc >> 6, tiles_r = R >> 6;
  const int b = blockIdx.x / (tiles_r * tiles_c), tt = blockIdx.x % (tiles_r * tiles_c);
  const int r0 = (tt / tiles_c) * 64, c0 = (tt % tiles_c) * 64;
  const float* sp = src + (long long)b * sb + (long long)r0 * sr + c0;
  {
    const int cq = (threadIdx.x & 15) * 4, rr = threadIdx.x >> 4;  // 16 threads x float4 = one 64-column row
#pragma unroll
    for (int pss = 0; pss < 4; ++pss) {
      const int r = rr + 16 * pss;
      const float4 v = __ldg(reinterpret_cast<const float4*>(sp + (long long)r * sr + cq));
      tile[r][cq] = v.x, tile[r][cq + 1] = v.y, tile[r][cq + 2] = v.z, tile[r][cq + 3] = v.w;
    }
  }
  __syncthreads();
  bf16* dp = dst + (long long)b * db + (long long)c0 * dc + r0;
  {
    const int rq = (threadIdx.x & 7) * 8, cc = threadIdx.x >> 3;  // 8 threads x 8 bf16 = one 64-row output line
#pragma unroll
    for (int pss = 0; pss < 2; ++pss) {
      const int c = cc + 32 * pss;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = tile[rq + j][c];
      *reinterpret_cast<uint4*>(dp + (long long)c * dc + rq) = t_pack8(f);
    }
  }
}

// ---------------------------------------------------------------- channel reductions
// mode 0: s0 = sum z, s1 = sum z^2                      (BatchNorm batch statistics)
// mode 1: g = dy * [z*scale+shift > 0 if relu]; xhat = (z-mean)*rstd; s0 = sum g, s1 = sum g*xhat
// mode 2: s0 = sum a, s1 = 0                            (bias gradients)
// partial[block][2][C]; chan_reduce_final sums the blocks in order.
struct ChanReduceP {
  const bf16* a;   // z (mode 0), dy (mode 1, 2)
  int lda;
  const bf16* z;   // mode 1
  int ldz;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* rstd;
  int relu, mode;
  long long P;
  int C;
  float* partial;
};
// Occupancy is what bounds these passes: a thread keeps per-channel constants for its 8 channels in registers, and at
// 110 registers only two 256-thread blocks fit an SM (ncu: 23 % warps active, 4.2 TB/s).  So the kernel is
// specialised per mode, mode 1 accumulates sum g and sum g*z and applies (z - mean) * rstd ONCE at the end
// (sum g*xhat = rstd * (sum g*z - mean * sum g): two constant vectors fewer in the loop), and __launch_bounds__ asks
// for three blocks per SM: 24 warps x 8 independent 16-byte loads in flight.
template <int MODE>
__global__ void __launch_bounds__(256, 3) chan_reduce_kernel(const ChanReduceP p) {
  __shared__ float red[2][256][8];
  const int groups = p.C >> 3, rows = 256 / groups;
  const int cg = threadIdx.x % groups, pr = threadIdx.x / groups;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
  float sc[8], sh[8];
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = p.scale[cg * 8 + j];
      sh[j] = p.shift[cg * 8 + j];
    }
  }
  // 4 rows per iteration with all loads issued first (the loop is latency bound otherwise)
  const long long stride = (long long)gridDim.x * rows;
  for (long long r0 = (long long)blockIdx.x * rows + pr; r0 < p.P; r0 += 4 * stride) {
    uint4 ra[4], rz[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u * stride;
      ra[u] = make_uint4(0u, 0u, 0u, 0u);
      rz[u] = make_uint4(0u, 0u, 0u, 0u);
      if (r < p.P) {
        ra[u] = ld8(p.a + r * p.lda + cg * 8);
        if (MODE == 1) rz[u] = ld8(p.z + r * p.ldz + cg * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r0 + u * stride >= p.P) break;
      float a[8];
      t_unpack8(ra[u], a);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s0[j] += a[j];
          s1[j] = fmaf(a[j], a[j], s1[j]);
        }
      } else if (MODE == 1) {
        float z[8];
        t_unpack8(rz[u], z);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g = (!p.relu || fmaf(z[j], sc[j], sh[j]) > 0.f) ? a[j] : 0.f;
          s0[j] += g;
          s1[j] = fmaf(g, z[j], s1[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s0[j] += a[j];
      }
    }
  }
  if (MODE == 1) {  // sum g*xhat = rstd * (sum g*z - mean * sum g), per thread (linear, so the partials add up)
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = (s1[j] - p.mean[cg * 8 + j] * s0[j]) * p.rstd[cg * 8 + j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x][j] = s0[j];
    red[1][threadIdx.x][j] = s1[j];
  }
  __syncthreads();
  if (pr == 0) {
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = 0.f;
        for (int q = 0; q < rows; ++q) t += red[k][q * groups + cg][j];
        p.partial[((size_t)blockIdx.x * 2 + k) * p.C + cg * 8 + j] = t;
      }
  }
}
// One warp per output: lane l sums partials l, l+32, ... in order, then a fixed-order shuffle tree (deterministic).
__global__ void chan_reduce_final_kernel(const float* __restrict__ partial, int n_blocks, int C2,
                                         float* __restrict__ out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= C2) return;
  float t = 0.f;
  for (int b = lane; b < n_blocks; b += 32) t += partial[(size_t)b * C2 + i];
  t = t_warp_sum(t);
  if (lane == 0) out[i] = t;
}

// --------------------------------------------------------------------------
// pack_bf16: every bf16 operand layout the training step needs (forward and data-gradient forms of the 19
// 3x3 convolutions, out.0, the three transposed convolutions — 46 strided 4-D permutations, flips expressed as
// negative strides) in ONE launch.  The weights change every optimizer step, so this runs once per step; as
// ~150 separate torch permute / flip / cast kernels it cost 0.7 ms, a fixed 13 % of the step at 32 images per GPU.
// A table row describes one tensor: out[i0][i1][i2][i3] = (bf16) src[off + i0 s0 + i1 s1 + i2 s2 + i3 s3]; a
// thread produces 8 consecutive outputs (one 16-byte store).
// --------------------------------------------------------------------------
struct PackRow {
  const float* src;
  bf16* dst;
  long long d1, d2, d3;      // output extents of dims 1..3 (dim 0 follows from the vector count)
  long long s0, s1, s2, s3;  // source strides in elements
  long long off;             // source offset in elements (flipped dims start at their last element)
  long long vec_start;       // first 8-element output vector of this tensor in the launch-wide numbering
  long long pad;             // rows are 12 x int64
};
static_assert(sizeof(PackRow) == 96, "cdm_pack_bf16 table rows are 12 x int64");
__global__ void __launch_bounds__(256) pack_bf16_kernel(const PackRow* __restrict__ rows, int n_rows, long long total_vec) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total_vec; v += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_rows - 1;  // last row with vec_start <= v
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (rows[mid].vec_start <= v) lo = mid; else hi = mid - 1;
    }
    const PackRow r = rows[lo];
    const long long e = (v - r.vec_start) * 8;  // linear output index of the first of 8 elements (d3 % 8 == 0)
    const long long i3 = e % r.d3, t2 = e / r.d3, i2 = t2 % r.d2, t1 = t2 / r.d2, i1 = t1 % r.d1, i0 = t1 / r.d1;
    const float* sp = r.src + r.off + i0 * r.s0 + i1 * r.s1 + i2 * r.s2 + i3 * r.s3;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __ldg(sp + j * r.s3);
    *reinterpret_cast<uint4*>(r.dst + e) = t_pack8(f);
  }
}

// --------------------------------------------------------------------------
// xrank_sum: the fixed-order final pass of a two-stage reduction FUSED with its cross-rank exchange over
// NVLink peer memory (data-parallel BatchNorm statistics; replaces chan_reduce_final + an NCCL all-reduce
// of [2C] floats, 36 times per training step).  One warp per output sums the per-CTA partials in order and
// stores the result into slot[rank] of EVERY rank's symmetric buffer (plain P2P stores).  The last block to
// finish (ticket) publishes a sequence number to every peer's flag word (st.release.sys), waits for the
// peers' (ld.acquire.sys, bounded spin -> trap instead of a hang), and adds the slots in rank order — so
// every rank obtains the bit-identical global sum, with no host involvement: the launch is graph-capturable
// and the sequence counter lives on the device.  Slots are double-buffered on the parity of the sequence
// number: a rank can be at most one exchange ahead of a peer.  world == 1 degenerates to the final pass.
// --------------------------------------------------------------------------
constexpr int kXrMaxN = 512;
struct XrankP {
  const float* partial;
  int n_blocks, n;
  float* out;
  int rank, world;
  const unsigned long long* peer_slots;  // [world] device pointers: fp32 [2][world][kXrMaxN] on each rank
  const unsigned long long* peer_flags;  // [world] device pointers: int32 [world] on each rank
  int* seq;
  unsigned int* ticket;
  unsigned long long timeout_ns;
};
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(1024) xrank_sum_kernel(const XrankP p) {
  __shared__ bool last;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int s = p.world > 1 ? *p.seq + 1 : 0;  // read before the ticket: only the last block advances it
  const int par = s & 1;
  if (i < p.n) {
    // four independent chains so that the (up to 19) strided loads of a lane are all in flight at once; the order of
    // the additions is still fixed
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    int b = lane;
    for (; b + 96 < p.n_blocks; b += 128) {
      t0 += p.partial[(size_t)b * p.n + i];
      t1 += p.partial[(size_t)(b + 32) * p.n + i];
      t2 += p.partial[(size_t)(b + 64) * p.n + i];
      t3 += p.partial[(size_t)(b + 96) * p.n + i];
    }
    for (; b < p.n_blocks; b += 32) t0 += p.partial[(size_t)b * p.n + i];
    float t = t_warp_sum((t0 + t1) + (t2 + t3));
    if (p.world == 1) {
      if (lane == 0) p.out[i] = t;
    } else if (lane < p.world) {  // lane q delivers to rank q (its own slot included)
      float* dst = reinterpret_cast<float*>(p.peer_slots[lane]);
      dst[((size_t)par * p.world + p.rank) * kXrMaxN + i] = t;
    }
  }
  if (p.world == 1) return;
  // one system-scope fence per block: bar.sync orders the block's peer stores before thread 0's fence, the fence
  // is cumulative, and the ticket (then the flag) is written after it
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < p.world && threadIdx.x != p.rank) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<int*>(p.peer_flags[threadIdx.x]) + p.rank, s);
    const int* mine = reinterpret_cast<const int*>(p.peer_flags[p.rank]) + threadIdx.x;
    // A peer may legitimately be late by a long time (rank 0 writing a checkpoint, a slow dataloader, lazy graph
    // capture): the bound is wall-clock time (cdm_xrank_set_timeout, default 600 s), where an NCCL all-reduce would
    // simply wait; it only exists so that a dead peer surfaces as a CUDA error instead of a silent hang.
    unsigned int spins = 0;
    unsigned long long t_start = 0;
    while (ld_acquire_sys(mine) - s < 0) {
      if ((++spins & 0xFFFu) == 0) {
        const unsigned long long now = global_timer_ns();
        if (t_start == 0) t_start = now;
        if (now - t_start > p.timeout_ns) {
          printf("cdm: cross-rank exchange timed out after %llu s (rank %d waiting for rank %d, seq %d)\n",
                 p.timeout_ns / 1000000000ull, p.rank, (int)threadIdx.x, s);
          __trap();
        }
        __nanosleep(200);  // back off: the spinning warp shares its SM with nothing useful, but spare the fabric
      }
    }
  }
  __syncthreads();
  const float* slots = reinterpret_cast<const float*>(p.peer_slots[p.rank]) + (size_t)par * p.world * kXrMaxN;
  for (int k = threadIdx.x; k < p.n; k += blockDim.x) {
    float acc = 0.f;
    for (int q = 0; q < p.world; ++q) acc += ld_relaxed_sys(slots + (size_t)q * kXrMaxN + k);
    p.out[k] = acc;
  }
  if (threadIdx.x == 0) {
    *p.seq = s;
    *p.ticket = 0;
  }
}
static unsigned long long g_xrank_timeout_ns = 0;
static unsigned long long xrank_timeout_ns() {
  if (g_xrank_timeout_ns == 0) {
    const char* e = getenv("CDM_XRANK_TIMEOUT_S");
    const double s = e ? atof(e) : 600.0;
    g_xrank_timeout_ns = (unsigned long long)((s > 0.001 ? s : 600.0) * 1e9);
  }
  return g_xrank_timeout_ns;
}
int launch_xrank_sum(const float* partial, int n_blocks, int n, float* out, const cdm_xrank* xr, cudaStream_t st) {
  XrankP p{partial, n_blocks, n, out, 0, 1, nullptr, nullptr, nullptr, nullptr, xrank_timeout_ns()};
  if (xr && xr->world > 1) {
    p.rank = xr->rank;
    p.world = xr->world;
    p.peer_slots = xr->peer_slots;
    p.peer_flags = xr->peer_flags;
    p.seq = xr->seq;
    p.ticket = xr->ticket;
  }
  xrank_sum_kernel<<<(n * 32 + 1023) / 1024, 1024, 0, st>>>(p);
  return 0;
}

// BatchNorm2d train-mode finalisation from (possibly all-reduced) sums over `count` elements.
__global__ void bn_finalize_kernel(const float* __restrict__ sums, int C, float count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = sums[c] / count;
  const float var = fmaxf(sums[C + c] / count - mean * mean, 0.f);  // biased, used for normalisation
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (count / fmaxf(count - 1.f, 1.f));
  }
}

// y = act(z*scale+shift) [+ w_c*x + b_c] ; optional second output yf = fs[n]*y + fb[n|0].
struct BnApplyP {
  const bf16* z;
  long long P;
  int C, relu;
  const float* scale;
  const float* shift;
  bf16* y;
  const float* sc_x;  // fp32 [P] or null
  const float* sc_w;
  const float* sc_b;
  const float* fs;  // film scale [n][C] or null
  const float* fb;  // film shift [rows][C]
  int fb_rows, px_per_img;
  bf16* yf;
  // inline BatchNorm finalisation (sums != null): scale / shift come from the batch sums, block 0 publishes them
  // (+ mean, rstd for the backward pass) and updates the running statistics — what cdm_bn_finalize would have done
  const float* sums;
  const float* gamma;
  const float* beta;
  float count, eps, momentum;
  float* running_mean;
  float* running_var;
  float* scale_out;
  float* shift_out;
  float* mean_out;
  float* rstd_out;
};
// Four rows per thread and iteration, all loads issued before the first use (one 16-byte load in flight per thread left
// the pass at 55-66 % of the copy bandwidth); the shortcut / FiLM extras are compile-time variants so that the plain
// Conv-BN-ReLU instance stays small enough for four blocks per SM.
template <bool EXTRAS>
__global__ void __launch_bounds__(256, 4) bn_apply_kernel(const BnApplyP p) {
  const int groups = p.C >> 3;  // 16 or 32: divides blockDim.x, so a thread keeps the same 8 channels
  const int cg = (int)(threadIdx.x % groups);
  float sc[8], sh[8];
  if (p.sums) {  // the arithmetic of bn_finalize_kernel, per thread for its own 8 channels (bit-identical results)
    const bool publish = blockIdx.x == 0 && threadIdx.x < groups;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      const float mean = p.sums[c] / p.count;
      const float var = fmaxf(p.sums[p.C + c] / p.count - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      sc[j] = p.gamma[c] * rstd;
      sh[j] = p.beta[c] - mean * sc[j];
      if (publish) {
        p.scale_out[c] = sc[j];
        p.shift_out[c] = sh[j];
        p.mean_out[c] = mean;
        p.rstd_out[c] = rstd;
        if (p.running_mean) {
          p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
          p.running_var[c] =
              (1.f - p.momentum) * p.running_var[c] + p.momentum * var * (p.count / fmaxf(p.count - 1.f, 1.f));
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = p.scale[cg * 8 + j];
      sh[j] = p.shift[cg * 8 + j];
    }
  }
  // groups divides blockDim.x, so a thread's row advances by a constant: no 64-bit division in the loop
  const long long r_step = (long long)gridDim.x * (blockDim.x / groups);
  for (long long r0 = (long long)blockIdx.x * (blockDim.x / groups) + threadIdx.x / groups; r0 < p.P; r0 += 4 * r_step) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u * r_step;
      raw[u] = r < p.P ? ld8(p.z + r * p.C + cg * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u * r_step;
      if (r >= p.P) break;
      float f[8];
      t_unpack8(raw[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(f[j], sc[j], sh[j]);
        f[j] = p.relu ? fmaxf(y, 0.f) : y;
      }
      if (EXTRAS && p.sc_x) {
        const float xv = p.sc_x[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += fmaf(__ldg(p.sc_w + cg * 8 + j), xv, __ldg(p.sc_b + cg * 8 + j));
      }
      *reinterpret_cast<uint4*>(p.y + r * p.C + cg * 8) = t_pack8(f);
      if (EXTRAS && p.fs) {
        const long long n = r / p.px_per_img;
        const float* fs = p.fs + n * p.C + cg * 8;
        const float* fb = p.fb + (p.fb_rows == 1 ? 0 : n) * p.C + cg * 8;
        // FiLM acts on the bf16-rounded y the next layer's backward sees
        float g[8];
        t_unpack8(t_pack8(f), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = fmaf(fs[j], g[j], fb[j]);
        *reinterpret_cast<uint4*>(p.yf + r * p.C + cg * 8) = t_pack8(g);
      }
    }
  }
}

// dz = scale * (g - S0/N - xhat * S1/N),  g = dy * relu-mask  (scale = gamma * rstd).
struct BnBwdP {
  const bf16* dy;
  int lddy;
  const bf16* z;
  long long P;
  int C, relu;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* rstd;
  const float* sums;  // [2][C] (all-reduced)
  float count;
  bf16* dz;
};
// dz = scale*g - A*z - B with A = scale*rstd*S1/N and B = scale*S0/N - A*mean (the expression above, expanded so that a
// thread keeps four constant vectors instead of six: 90 -> <= 85 registers, three blocks per SM instead of two);
// four rows per iteration, all eight loads issued first.
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_kernel(const BnBwdP p) {
  const int groups = p.C >> 3;  // divides blockDim.x: a thread keeps the same 8 channels
  const float inv = 1.f / p.count;
  const int cg = (int)(threadIdx.x % groups);
  float sc[8], sh[8], ca[8], cb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    sc[j] = p.scale[c];
    sh[j] = p.shift[c];
    ca[j] = sc[j] * p.rstd[c] * (p.sums[p.C + c] * inv);
    cb[j] = sc[j] * (p.sums[c] * inv) - ca[j] * p.mean[c];
  }
  const long long r_step = (long long)gridDim.x * (blockDim.x / groups);
  for (long long r0 = (long long)blockIdx.x * (blockDim.x / groups) + threadIdx.x / groups; r0 < p.P; r0 += 4 * r_step) {
    uint4 rd[4], rz[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u * r_step;
      rd[u] = rz[u] = make_uint4(0u, 0u, 0u, 0u);
      if (r < p.P) {
        rd[u] = ld8(p.dy + r * p.lddy + cg * 8);
        rz[u] = ld8(p.z + r * p.C + cg * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u * r_step;
      if (r >= p.P) break;
      float d[8], z[8];
      t_unpack8(rd[u], d);
      t_unpack8(rz[u], z);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = (!p.relu || fmaf(z[j], sc[j], sh[j]) > 0.f) ? d[j] : 0.f;
        d[j] = fmaf(sc[j], g, -fmaf(ca[j], z[j], cb[j]));
      }
      *reinterpret_cast<uint4*>(p.dz + r * p.C + cg * 8) = t_pack8(d);
    }
  }
}

// ---------------------------------------------------------------- MaxPool2d(2)
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const bf16* __restrict__ y, int n_img, int H, int W, int C,
                                                           bf16* __restrict__ out) {
  const int groups = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)n_img * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    long long r = i / groups;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const long long n = r / Ho;
    const bf16* b = y + (((n * H + 2 * ho) * W) + 2 * wo) * C + cg * 8;
    const uint4 v0 = ld8(b), v1 = ld8(b + C), v2 = ld8(b + (size_t)W * C), v3 = ld8(b + (size_t)W * C + C);
    uint4 m;
    m.x = bf16x2_max(bf16x2_max(v0.x, v1.x), bf16x2_max(v2.x, v3.x));
    m.y = bf16x2_max(bf16x2_max(v0.y, v1.y), bf16x2_max(v2.y, v3.y));
    m.z = bf16x2_max(bf16x2_max(v0.z, v1.z), bf16x2_max(v2.z, v3.z));
    m.w = bf16x2_max(bf16x2_max(v0.w, v1.w), bf16x2_max(v2.w, v3.w));
    *reinterpret_cast<uint4*>(out + (((n * Ho + ho) * Wo) + wo) * C + cg * 8) = m;
  }
}
// dy[px] = dpool if px is the FIRST maximum of its 2x2 window (torch's tie rule), else 0.
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const bf16* __restrict__ dpool, int lddp,
                                                           const bf16* __restrict__ y, int n_img, int H, int W, int C,
                                                           bf16* __restrict__ dy) {
  const int groups = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)n_img * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    long long r = i / groups;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const long long n = r / Ho;
    const size_t base = ((((size_t)n * H + 2 * ho) * W) + 2 * wo) * C + cg * 8;
    const size_t off[4] = {0, (size_t)C, (size_t)W * C, (size_t)W * C + C};
    float v[4][8], g[8], o[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) t_unpack8(ld8(y + base + off[q]), v[q]);
    t_unpack8(ld8(dpool + (((size_t)n * Ho + ho) * Wo + wo) * lddp + cg * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = fmaxf(fmaxf(v[0][j], v[1][j]), fmaxf(v[2][j], v[3][j]));
      bool taken = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool hit = !taken && v[q][j] == m;
        o[q][j] = hit ? g[j] : 0.f;
        taken = taken || hit;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dy + base + off[q]) = t_pack8(o[q]);
  }
}

// a[r][c] += b[r][c]   (gradient accumulation at the skip connections)
__global__ void __launch_bounds__(256) add_bf16_kernel(bf16* __restrict__ a, int lda, const bf16* __restrict__ b,
                                                       int ldb, long long P, int C) {
  const int groups = C >> 3;
  const long long total = P * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    const long long r = i / groups;
    float x[8], y[8];
    t_unpack8(ld8(a + r * lda + cg * 8), x);
    t_unpack8(ld8(b + r * ldb + cg * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    *reinterpret_cast<uint4*>(a + r * lda + cg * 8) = t_pack8(x);
  }
}

// dv [n][2H][2W][C] -> s2d [n][H][W][(kh,kw,c)]  (A operand of the transposed-conv dgrad / wgrad GEMMs)
__global__ void __launch_bounds__(256) space_to_depth_kernel(const bf16* __restrict__ dv, int n_img, int H, int W,
                                                             int C, bf16* __restrict__ out) {
  const int groups = C >> 3;
  const long long total = (long long)n_img * H * W * 4 * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    long long r = i / groups;
    const int k = (int)(r & 3);
    r >>= 2;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const long long n = r / H;
    const uint4 v = ld8(dv + (((n * 2 * H + 2 * h + (k >> 1)) * (2 * W)) + 2 * w + (k & 1)) * C + cg * 8);
    *reinterpret_cast<uint4*>(out + (((n * H + h) * W + w) * 4 + k) * C + cg * 8) = v;
  }
}

// ---------------------------------------------------------------- FiLM backward
// yf = fs[n][c]*y + fb: dy = fs*dyf; dfs[n][c] = sum_px dyf*y; dfb[n][c] = sum_px dyf.  One block per image.
__global__ void __launch_bounds__(256) film_bwd_kernel(const bf16* __restrict__ dyf, int lddyf,
                                                       const bf16* __restrict__ y, int px, int C,
                                                       const float* __restrict__ fs, bf16* __restrict__ dy,
                                                       float* __restrict__ dfs, float* __restrict__ dfb) {
  __shared__ float red[2][256][8];
  const size_t n = blockIdx.x;
  const int groups = C >> 3, rows = 256 / groups;
  const int cg = threadIdx.x % groups, pr = threadIdx.x / groups;
  float a0[8], a1[8], s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a0[j] = a1[j] = 0.f;
    s[j] = fs[n * C + cg * 8 + j];
  }
  // four rows per iteration, all eight loads issued first (one block per image: at 32 images per GPU the kernel is a
  // latency chain, 42 us with one row in flight); the order of the additions per thread is unchanged
  for (int r0 = pr; r0 < px; r0 += 4 * rows) {
    uint4 rd[4], rv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * rows;
      rd[u] = rv[u] = make_uint4(0u, 0u, 0u, 0u);
      if (r < px) {
        rd[u] = ld8(dyf + (n * px + r) * lddyf + cg * 8);
        rv[u] = ld8(y + (n * px + r) * C + cg * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * rows;
      if (r >= px) break;
      float d[8], v[8];
      t_unpack8(rd[u], d);
      t_unpack8(rv[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0[j] = fmaf(d[j], v[j], a0[j]);
        a1[j] += d[j];
        d[j] *= s[j];
      }
      *reinterpret_cast<uint4*>(dy + (n * px + r) * C + cg * 8) = t_pack8(d);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x][j] = a0[j];
    red[1][threadIdx.x][j] = a1[j];
  }
  __syncthreads();
  if (pr == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t0 = 0.f, t1 = 0.f;
      for (int q = 0; q < rows; ++q) {
        t0 += red[0][q * groups + cg][j];
        t1 += red[1][q * groups + cg][j];
      }
      dfs[n * C + cg * 8 + j] = t0;
      dfb[n * C + cg * 8 + j] = t1;
    }
  }
}

// ---------------------------------------------------------------- GroupNorm (+ReLU, +FiLM) backward
// forward: h = (x-mean)*rstd; y = relu(h*gamma+beta); yf = fs*y + fb (FiLM optional).
// One block per (image, group).  Outputs dx (bf16) and per-(image, channel) dgamma/dbeta/dfs/dfb.
struct GnBwdP {
  const bf16* x;
  const bf16* dyf;
  int lddyf;
  int P, C, groups;
  const float* mean_rstd;
  const float* gamma;
  const float* beta;
  const float* fs;
  bf16* dx;
  float* dgamma_nc;
  float* dbeta_nc;
  float* dfs;
  float* dfb;
};
__global__ void __launch_bounds__(256) gn_bwd_kernel(const GnBwdP p) {
  __shared__ float red[32];
  __shared__ float chan[4][32];  // dgamma, dbeta, dfs, dfb per channel of the group (cpg <= 32)
  const int n = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
  const int cpg = p.C / p.groups, vpp = cpg / 8, n_vec = p.P * vpp;
  const float mean = p.mean_rstd[((size_t)n * p.groups + g) * 2], rstd = p.mean_rstd[((size_t)n * p.groups + g) * 2 + 1];
  if (threadIdx.x < 128) chan[threadIdx.x >> 5][threadIdx.x & 31] = 0.f;
  __syncthreads();
  const bf16* xb = p.x + (size_t)n * p.P * p.C + g * cpg;
  const bf16* db = p.dyf + (size_t)n * p.P * p.lddyf + g * cpg;
  // a thread always visits the same 8 channels (blockDim.x is a multiple of vpp)
  const int v = threadIdx.x % vpp;
  float ga[8], be[8], fsv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = p.gamma[g * cpg + v * 8 + j];
    be[j] = p.beta[g * cpg + v * 8 + j];
    fsv[j] = p.fs ? p.fs[(size_t)n * p.C + g * cpg + v * 8 + j] : 1.f;
  }
  float S1 = 0.f, S2 = 0.f, dg[8], dbt[8], dfs[8], dfb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dg[j] = dbt[j] = dfs[j] = dfb[j] = 0.f;
  for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
    const int px = i / vpp;
    float x[8], d[8];
    t_unpack8(ld8(xb + (size_t)px * p.C + v * 8), x);
    t_unpack8(ld8(db + (size_t)px * p.lddyf + v * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float h = (x[j] - mean) * rstd;
      const float pre = fmaf(h, ga[j], be[j]);
      const float y = fmaxf(pre, 0.f);
      dfs[j] = fmaf(d[j], y, dfs[j]);
      dfb[j] += d[j];
      const float dyv = pre > 0.f ? d[j] * fsv[j] : 0.f;
      dg[j] = fmaf(dyv, h, dg[j]);
      dbt[j] += dyv;
      const float dh = dyv * ga[j];
      S1 += dh;
      S2 = fmaf(dh, h, S2);
    }
  }
  // per-channel sums over the threads that own the channel (thread t owns vector t % vpp), added in thread order:
  // deterministic (shared-memory float atomics are not)
  __shared__ float fold[256][8];
#pragma unroll 1
  for (int qn = 0; qn < 4; ++qn) {
    const float* src = qn == 0 ? dg : (qn == 1 ? dbt : (qn == 2 ? dfs : dfb));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) fold[threadIdx.x][j] = src[j];
    __syncthreads();
    if (threadIdx.x < cpg) {
      const int vv = threadIdx.x >> 3, jj = threadIdx.x & 7;
      float acc = 0.f;
      for (int t = vv; t < (int)blockDim.x; t += vpp) acc += fold[t][jj];
      chan[qn][threadIdx.x] = acc;
    }
  }
  __syncthreads();
  const float M = (float)(p.P * cpg);
  const float s1 = t_block_sum(S1, red) / M;
  const float s2 = t_block_sum(S2, red) / M;
  if (threadIdx.x < cpg) {
    const size_t o = (size_t)n * p.C + g * cpg + threadIdx.x;
    p.dgamma_nc[o] = chan[0][threadIdx.x];
    p.dbeta_nc[o] = chan[1][threadIdx.x];
    if (p.dfs) {
      p.dfs[o] = chan[2][threadIdx.x];
      p.dfb[o] = chan[3][threadIdx.x];
    }
  }
  bf16* ob = p.dx + (size_t)n * p.P * p.C + g * cpg;
  for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
    const int px = i / vpp;
    float x[8], d[8];
    t_unpack8(ld8(xb + (size_t)px * p.C + v * 8), x);
    t_unpack8(ld8(db + (size_t)px * p.lddyf + v * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float h = (x[j] - mean) * rstd;
      const float dh = fmaf(h, ga[j], be[j]) > 0.f ? d[j] * fsv[j] * ga[j] : 0.f;
      d[j] = rstd * (dh - s1 - h * s2);
    }
    *reinterpret_cast<uint4*>(ob + (size_t)px * p.C + v * 8) = t_pack8(d);
  }
}

// Image-major variant for batches that fill the machine with one block per image: a warp reads 512 contiguous bytes
// (all groups of one or two pixels) instead of 32-byte pieces of 16 pixels, which is what held the (image, group)
// mapping at ~1.9 TB/s.  Same math; the per-group / per-channel sums are folded in thread order (deterministic).
__global__ void __launch_bounds__(256) gn_bwd_img_kernel(const GnBwdP p) {
  __shared__ float fold[256][8];
  __shared__ float gsum[256][2];
  __shared__ float s12[8][2];
  const int n = blockIdx.x;
  const int vpa = p.C >> 3;               // 16-byte vectors per pixel (16 or 32)
  const int lanes = blockDim.x / vpa;     // pixel lanes (16 or 8)
  const int vec = threadIdx.x % vpa, pl = threadIdx.x / vpa;
  const int cpg = p.C / p.groups, g = (vec * 8) / cpg;
  const float mean = p.mean_rstd[((size_t)n * p.groups + g) * 2], rstd = p.mean_rstd[((size_t)n * p.groups + g) * 2 + 1];
  const bf16* xb = p.x + (size_t)n * p.P * p.C + vec * 8;
  const bf16* db = p.dyf + (size_t)n * p.P * p.lddyf + vec * 8;
  float ga[8], be[8], fsv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = p.gamma[vec * 8 + j];
    be[j] = p.beta[vec * 8 + j];
    fsv[j] = p.fs ? p.fs[(size_t)n * p.C + vec * 8 + j] : 1.f;
  }
  float S1 = 0.f, S2 = 0.f, dg[8], dbt[8], dfs[8], dfb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dg[j] = dbt[j] = dfs[j] = dfb[j] = 0.f;
  for (int px = pl; px < p.P; px += 2 * lanes) {
    const int px2 = px + lanes;
    const bool two = px2 < p.P;
    const uint4 rx0 = ld8(xb + (size_t)px * p.C), rd0 = ld8(db + (size_t)px * p.lddyf);
    uint4 rx1 = make_uint4(0u, 0u, 0u, 0u), rd1 = rx1;
    if (two) {
      rx1 = ld8(xb + (size_t)px2 * p.C);
      rd1 = ld8(db + (size_t)px2 * p.lddyf);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float x[8], d[8];
      t_unpack8(u ? rx1 : rx0, x);
      t_unpack8(u ? rd1 : rd0, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = (x[j] - mean) * rstd;
        const float pre = fmaf(h, ga[j], be[j]);
        const float y = fmaxf(pre, 0.f);
        dfs[j] = fmaf(d[j], y, dfs[j]);
        dfb[j] += d[j];
        const float dyv = pre > 0.f ? d[j] * fsv[j] : 0.f;
        dg[j] = fmaf(dyv, h, dg[j]);
        dbt[j] += dyv;
        const float dh = dyv * ga[j];
        S1 += dh;
        S2 = fmaf(dh, h, S2);
      }
    }
  }
  // per-channel sums over the pixel lanes that share a vector
#pragma unroll 1
  for (int qn = 0; qn < 4; ++qn) {
    const float* src = qn == 0 ? dg : (qn == 1 ? dbt : (qn == 2 ? dfs : dfb));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) fold[threadIdx.x][j] = src[j];
    __syncthreads();
    if (threadIdx.x < p.C && (qn < 2 || p.dfs)) {
      const int vv = threadIdx.x >> 3, jj = threadIdx.x & 7;
      float acc = 0.f;
      for (int t = vv; t < (int)blockDim.x; t += vpa) acc += fold[t][jj];
      float* dst = qn == 0 ? p.dgamma_nc : (qn == 1 ? p.dbeta_nc : (qn == 2 ? p.dfs : p.dfb));
      dst[(size_t)n * p.C + threadIdx.x] = acc;
    }
  }
  // per-group sums S1, S2 over the threads of the group
  __syncthreads();
  gsum[threadIdx.x][0] = S1;
  gsum[threadIdx.x][1] = S2;
  __syncthreads();
  if (threadIdx.x < p.groups) {
    const int vpg = cpg >> 3;  // vectors per group
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l)
      for (int vv = 0; vv < vpg; ++vv) {
        const int t = l * vpa + threadIdx.x * vpg + vv;
        a += gsum[t][0];
        b += gsum[t][1];
      }
    const float M = (float)p.P * (float)cpg;
    s12[threadIdx.x][0] = a / M;
    s12[threadIdx.x][1] = b / M;
  }
  __syncthreads();
  const float s1 = s12[g][0], s2 = s12[g][1];
  bf16* ob = p.dx + (size_t)n * p.P * p.C + vec * 8;
  for (int px = pl; px < p.P; px += 2 * lanes) {
    const int px2 = px + lanes;
    const bool two = px2 < p.P;
    const uint4 rx0 = ld8(xb + (size_t)px * p.C), rd0 = ld8(db + (size_t)px * p.lddyf);
    uint4 rx1 = make_uint4(0u, 0u, 0u, 0u), rd1 = rx1;
    if (two) {
      rx1 = ld8(xb + (size_t)px2 * p.C);
      rd1 = ld8(db + (size_t)px2 * p.lddyf);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float x[8], d[8];
      t_unpack8(u ? rx1 : rx0, x);
      t_unpack8(u ? rd1 : rd0, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = (x[j] - mean) * rstd;
        const float dh = fmaf(h, ga[j], be[j]) > 0.f ? d[j] * fsv[j] * ga[j] : 0.f;
        d[j] = rstd * (dh - s1 - h * s2);
      }
      *reinterpret_cast<uint4*>(ob + (size_t)(u ? px2 : px) * p.C) = t_pack8(d);
    }
  }
}

// out[c] = sum_n in[n][c]  (sum the per-image partials over the batch; tiny)
__global__ void rows_sum_kernel(const float* __restrict__ in, int rows, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int r = 0; r < rows; ++r) t += in[(size_t)r * C + c];
  out[c] = t;
}

// d_d2[n][px][c] += gelu'(pre[n][c]) * dh[n][c] / P    (to_vec backward)
__global__ void __launch_bounds__(256) avgpool_gelu_bwd_kernel(const float* __restrict__ pre,
                                                               const float* __restrict__ dh, int P, int C,
                                                               bf16* __restrict__ dx) {
  const size_t n = blockIdx.x;
  const int groups = C >> 3;
  // grid.y slices the image's pixels: at 32 images per GPU one block per image left 116 SMs idle for 44 us
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P * groups; i += gridDim.y * blockDim.x) {
    const int cg = i % groups, px = i / groups;
    float d[8];
    t_unpack8(ld8(dx + (n * P + px) * C + cg * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      d[j] += gelu_grad(pre[n * C + c]) * dh[n * C + c] / (float)P;
    }
    *reinterpret_cast<uint4*>(dx + (n * P + px) * C + cg * 8) = t_pack8(d);
  }
}
__global__ void __launch_bounds__(256) avgpool_pre_kernel(const bf16* __restrict__ src, int P, int C,
                                                          float* __restrict__ pre, bf16* __restrict__ out) {
  const size_t n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int px = 0; px < P; ++px) s += __bfloat162float(src[(n * P + px) * C + c]);
    s /= (float)P;
    pre[n * C + c] = s;
    out[n * C + c] = __float2bfloat16(gelu_f(s));
  }
}

// ---------------------------------------------------------------- K=9 / N=1 convolution weight gradients
// dW[tap][c] = sum_px s[px + d(tap)] * v[px][c],  d(tap) = (kh-1, kw-1) (flip = 0) or its negative (flip = 1).
//  * conv_in  wgrad: s = x (network input),   v = dz1,                      flip = 0
//  * conv_out wgrad: s = d eps,               v = relu(GroupNorm(o)) on load, flip = 1
// partial[block][9][C] -> chan_reduce_final.
struct OuterWgradP {
  const float* s;
  const bf16* v;
  int n_img, H, W, C, flip;
  const float* mean_rstd;  // non-null: apply GroupNorm(8)+ReLU to v on load
  const float* gamma;
  const float* beta;
  float* partial;
};
__global__ void __launch_bounds__(256) outer_wgrad_kernel(const OuterWgradP p) {
  __shared__ float red[256][8];
  const int groups = p.C >> 3, rows = 256 / groups;
  const int cg = threadIdx.x % groups, pr = threadIdx.x / groups;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int cpg = p.C / 8;
  const int n_rows = p.n_img * p.H;
  // one image row per block iteration, four pixels per thread with all activation loads issued first (one 16-byte
  // load in flight per thread left the kernel latency bound at ~1.2 TB/s); 32-bit index arithmetic
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int h = row % p.H, n = row / p.H;
    const bf16* vrow = p.v + (size_t)row * p.W * p.C + cg * 8;
    const float* srow = p.s + (size_t)row * p.W;
    float mean = 0.f, rstd = 1.f;
    if (p.mean_rstd) {
      const int g = (cg * 8) / cpg;
      mean = p.mean_rstd[((size_t)n * 8 + g) * 2];
      rstd = p.mean_rstd[((size_t)n * 8 + g) * 2 + 1];
    }
    for (int w0 = pr; w0 < p.W; w0 += 4 * rows) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int w = w0 + u * rows;
        raw[u] = w < p.W ? ld8(vrow + (size_t)w * p.C) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int w = w0 + u * rows;
        if (w >= p.W) break;
        float v[8];
        t_unpack8(raw[u], v);
        if (p.mean_rstd) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[j] = fmaxf(fmaf((v[j] - mean) * rstd, p.gamma[cg * 8 + j], p.beta[cg * 8 + j]), 0.f);
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int dh = p.flip ? 1 - kh : kh - 1, dw = p.flip ? 1 - kw : kw - 1;
            const int hh = h + dh, ww = w + dw;
            const float sv = (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) ? __ldg(srow + dh * p.W + ww) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[kh * 3 + kw][j] = fmaf(sv, v[j], acc[kh * 3 + kw][j]);
          }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[t][j];
    __syncthreads();
    if (pr == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
        for (int q = 0; q < rows; ++q) s += red[q * groups + cg][j];
        p.partial[((size_t)blockIdx.x * 9 + t) * p.C + cg * 8 + j] = s;
      }
    }
  }
}

// ---------------------------------------------------------------- EmbedFC backward (tiny)
// forward: pre = W1 v + b1, h = gelu(pre), out = W2 h + b2.
__global__ void __launch_bounds__(256) embed_hidden_kernel(const float* __restrict__ in, int din,
                                                           const float* __restrict__ w1, const float* __restrict__ b1,
                                                           int emb, float* __restrict__ pre, float* __restrict__ h) {
  const size_t r = blockIdx.x;
  for (int j = threadIdx.x; j < emb; j += blockDim.x) {
    float s = b1[j];
    for (int i = 0; i < din; ++i) s = fmaf(w1[j * din + i], in[r * din + i], s);
    pre[r * emb + j] = s;
    h[r * emb + j] = gelu_f(s);
  }
}
// block j: dW2[j][k] = sum_r dout[r][j] h[r][k]; db2[j] = sum_r dout[r][j]
__global__ void __launch_bounds__(256) embed_bwd_w2_kernel(const float* __restrict__ dout, const float* __restrict__ h,
                                                           int rows, int emb, float* __restrict__ dw2,
                                                           float* __restrict__ db2) {
  const int j = blockIdx.x;
  for (int k = threadIdx.x; k < emb; k += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < rows; ++r) s = fmaf(__ldg(dout + (size_t)r * emb + j), __ldg(h + (size_t)r * emb + k), s);
    dw2[(size_t)j * emb + k] = s;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < rows; ++r) s += __ldg(dout + (size_t)r * emb + j);
    db2[j] = s;
  }
}
// block r: dpre[r][k] = gelu'(pre[r][k]) * sum_j dout[r][j] W2[j][k]
__global__ void __launch_bounds__(256) embed_bwd_hidden_kernel(const float* __restrict__ dout,
                                                               const float* __restrict__ w2,
                                                               const float* __restrict__ pre, int emb,
                                                               float* __restrict__ dpre) {
  extern __shared__ float s_d[];
  const size_t r = blockIdx.x;
  for (int j = threadIdx.x; j < emb; j += blockDim.x) s_d[j] = dout[r * emb + j];
  __syncthreads();
  for (int k = threadIdx.x; k < emb; k += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int j = 0; j < emb; ++j) s = fmaf(s_d[j], __ldg(w2 + (size_t)j * emb + k), s);
    dpre[r * emb + k] = s * gelu_grad(pre[r * emb + k]);
  }
}
// thread (k): dW1[k][i] = sum_r dpre[r][k] in[r][i]; db1[k] = sum_r dpre[r][k]
__global__ void embed_bwd_w1_kernel(const float* __restrict__ dpre, const float* __restrict__ in, int rows, int din,
                                    int emb, float* __restrict__ dw1, float* __restrict__ db1) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= emb) return;
  float sb = 0.f;
#pragma unroll 8
  for (int r = 0; r < rows; ++r) sb += __ldg(dpre + (size_t)r * emb + k);
  db1[k] = sb;
  for (int i = 0; i < din; ++i) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < rows; ++r) s = fmaf(__ldg(dpre + (size_t)r * emb + k), __ldg(in + (size_t)r * din + i), s);
    dw1[(size_t)k * din + i] = s;
  }
}

// ---------------------------------------------------------------- loss + Adam
// F.mse_loss(pred, noise) (mean over all elements): partial sums of (pred-noise)^2 per block and
// d pred = 2 (pred - noise) * inv_count (inv_count = 1 / GLOBAL element count in data-parallel runs).
__global__ void __launch_bounds__(256) mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                       long long n4, float inv_count, float* __restrict__ dpred,
                                                       float* __restrict__ partial) {
  __shared__ float red[32];
  float q = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(pred)[v], b = reinterpret_cast<const float4*>(tgt)[v];
    float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    q += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    const float s = 2.f * inv_count;
    reinterpret_cast<float4*>(dpred)[v] = make_float4(d.x * s, d.y * s, d.z * s, d.w * s);
  }
  const float t = t_block_sum(q, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// torch.optim.Adam defaults (no weight decay, no amsgrad), all tensors in one launch.
struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};
// Persistent 1-D grid over 4096-element chunks of ALL tensors (a block finds its tensor in a chunk prefix table it builds
// in shared memory), four float4 per thread and operand with all sixteen loads issued first.  The first version ran
// one grid row per tensor: up0.0.weight (78 % of the 21.6 M parameters) was left to 512 blocks of scalar loads and the
// other ~50 k blocks exited at once — 0.226 ms for 605 MB (2.7 TB/s), a fixed 6 % of the 32-images-per-GPU step.
constexpr int kAdamChunk = 256 * 16;  // elements per block iteration
constexpr int kAdamMaxTensors = 512;
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float beta1, float beta2, float eps,
                                            float step_size, float sqrt_bc2) {
  m = beta1 * m + (1.f - beta1) * g;
  v = beta2 * v + (1.f - beta2) * g * g;
  // torch: denom = sqrt(v)/sqrt(bc2) + eps; p -= (lr/bc1) * m / denom
  const float denom = sqrtf(v) / sqrt_bc2 + eps;
  p -= step_size * (m / denom);
}
__global__ void __launch_bounds__(256) adam_kernel(const AdamTensor* __restrict__ tab, int n_tensors, float lr,
                                                   float beta1, float beta2, float eps, float bc1, float bc2,
                                                   const float* __restrict__ lr_dev,
                                                   const int* __restrict__ step_dev) {
  __shared__ long long pre[kAdamMaxTensors + 1];  // first chunk of every tensor
  if (lr_dev) lr = *lr_dev;  // device-resident hyper-parameters: the launch can be replayed from a CUDA graph
  if (step_dev) {
    const float st = (float)*step_dev;
    bc1 = 1.f - powf(beta1, st);
    bc2 = 1.f - powf(beta2, st);
  }
  // lr / bc1 and sqrt(bc2) are loop invariants of torch's expression (same values, same roundings)
  const float step_size = lr / bc1, sqrt_bc2 = sqrtf(bc2);
  for (int t = threadIdx.x; t < n_tensors; t += blockDim.x) pre[t] = (tab[t].n + kAdamChunk - 1) / kAdamChunk;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long acc = 0;
    for (int t = 0; t < n_tensors; ++t) {
      const long long c = pre[t];
      pre[t] = acc;
      acc += c;
    }
    pre[n_tensors] = acc;
  }
  __syncthreads();
  const long long total = pre[n_tensors];
  for (long long c = blockIdx.x; c < total; c += gridDim.x) {
    int lo = 0, hi = n_tensors - 1;  // last t with pre[t] <= c
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (pre[mid] <= c) lo = mid; else hi = mid - 1;
    }
    const AdamTensor t = tab[lo];
    const long long base = (c - pre[lo]) * kAdamChunk;
    const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                       reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0 &&
                     base + kAdamChunk <= t.n;
    if (vec) {
      float4 g4[4], m4[4], v4[4], p4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = base + (long long)(u * 256 + threadIdx.x) * 4;
        g4[u] = *reinterpret_cast<const float4*>(t.g + i);
        m4[u] = *reinterpret_cast<const float4*>(t.m + i);
        v4[u] = *reinterpret_cast<const float4*>(t.v + i);
        p4[u] = *reinterpret_cast<const float4*>(t.p + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = base + (long long)(u * 256 + threadIdx.x) * 4;
        adam_update(p4[u].x, g4[u].x, m4[u].x, v4[u].x, beta1, beta2, eps, step_size, sqrt_bc2);
        adam_update(p4[u].y, g4[u].y, m4[u].y, v4[u].y, beta1, beta2, eps, step_size, sqrt_bc2);
        adam_update(p4[u].z, g4[u].z, m4[u].z, v4[u].z, beta1, beta2, eps, step_size, sqrt_bc2);
        adam_update(p4[u].w, g4[u].w, m4[u].w, v4[u].w, beta1, beta2, eps, step_size, sqrt_bc2);
        *reinterpret_cast<float4*>(t.m + i) = m4[u];
        *reinterpret_cast<float4*>(t.v + i) = v4[u];
        *reinterpret_cast<float4*>(t.p + i) = p4[u];
      }
    } else {
      const long long end = base + kAdamChunk < t.n ? base + kAdamChunk : t.n;
      for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
        float p = t.p[i], m = t.m[i], v = t.v[i];
        adam_update(p, t.g[i], m, v, beta1, beta2, eps, step_size, sqrt_bc2);
        t.m[i] = m;
        t.v[i] = v;
        t.p[i] = p;
      }
    }
  }
}

static int grid1d(long long work, int block = 256, int cap = num_sms() * 8) {
  long long g = (work + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace cdm

using namespace cdm;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int cdm_chan_reduce(const cdm_chan_reduce_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->a && a->out && a->workspace && a->P > 0 && a->C > 0 && a->C % 8 == 0 && 256 % (a->C / 8) == 0);
  CDM_CHECK_ARG(a->mode >= 0 && a->mode <= 2 && a->lda >= a->C);
  if (a->mode == 1) CDM_CHECK_ARG(a->z && a->scale && a->shift && a->mean && a->rstd && a->ldz >= a->C);
  CDM_CHECK_ARG(2 * a->C <= kXrMaxN);
  int rc = check_device();
  if (rc) return rc;
  const int rows = 256 / (a->C / 8);
  int blocks = (int)((a->P + rows * 8 - 1) / (rows * 8));
  if (blocks > num_sms() * 3) blocks = num_sms() * 3;  // three resident blocks per SM = one wave; the final pass walks these partials
  if (blocks > a->workspace_blocks) blocks = a->workspace_blocks;
  if (blocks < 1) blocks = 1;
  ChanReduceP p{(const bf16*)a->a, a->lda, (const bf16*)a->z, a->ldz, a->scale, a->shift, a->mean, a->rstd,
                a->relu, a->mode, a->P, a->C, a->workspace};
  if (a->mode == 0)
    chan_reduce_kernel<0><<<blocks, 256, 0, ST(stream)>>>(p);
  else if (a->mode == 1)
    chan_reduce_kernel<1><<<blocks, 256, 0, ST(stream)>>>(p);
  else
    chan_reduce_kernel<2><<<blocks, 256, 0, ST(stream)>>>(p);
  CDM_CHECK_LAUNCH();
  // fixed-order final pass, fused with the cross-rank exchange when a->xr describes a peer group
  launch_xrank_sum(a->workspace, blocks, 2 * a->C, a->out, a->xr, ST(stream));
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_pack_bf16(const void* table, int n_rows, long long total_vec, void* stream) {
  CDM_CHECK_ARG(table && n_rows > 0 && total_vec > 0);
  int rc = check_device();
  if (rc) return rc;
  long long blocks = (total_vec + 255) / 256;
  if (blocks > num_sms() * 32) blocks = num_sms() * 32;
  pack_bf16_kernel<<<(int)blocks, 256, 0, ST(stream)>>>(reinterpret_cast<const PackRow*>(table), n_rows, total_vec);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_pack_transpose_bf16(const float* src, void* dst, int batches, int R, int Cc, long long src_batch_stride,
                                       long long src_row_stride, long long dst_batch_stride, long long dst_col_stride,
                                       void* stream) {
  CDM_CHECK_ARG(src && dst && batches > 0 && R > 0 && Cc > 0 && R % 64 == 0 && Cc % 64 == 0);
  CDM_CHECK_ARG(src_row_stride % 4 == 0 && src_batch_stride % 4 == 0 && dst_col_stride % 8 == 0 &&
                dst_batch_stride % 8 == 0);
  CDM_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  int rc = check_device();
  if (rc) return rc;
  const long long blocks = (long long)batches * (R / 64) * (Cc / 64);
  CDM_CHECK_ARG(blocks < (1ll << 31));
  pack_transpose_kernel<<<(unsigned)blocks, 256, 0, ST(stream)>>>(src, reinterpret_cast<bf16*>(dst), R, Cc,
                                                                   src_batch_stride, src_row_stride, dst_batch_stride,
                                                                   dst_col_stride);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_xrank_set_timeout(double seconds) {
  CDM_CHECK_ARG(seconds > 0.001 && seconds < 1e7);
  g_xrank_timeout_ns = (unsigned long long)(seconds * 1e9);
  return CDM_OK;
}

extern "C" int cdm_xrank_sum(const float* partial, int n_blocks, int n, float* out, const cdm_xrank* xr, void* stream) {
  CDM_CHECK_ARG(partial && out && n_blocks > 0 && n > 0 && n <= kXrMaxN);
  if (xr && xr->world > 1)
    CDM_CHECK_ARG(xr->world <= 32 && xr->rank >= 0 && xr->rank < xr->world && xr->peer_slots && xr->peer_flags &&
                  xr->seq && xr->ticket);
  int rc = check_device();
  if (rc) return rc;
  launch_xrank_sum(partial, n_blocks, n, out, xr, ST(stream));
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_bn_finalize(const float* sums, int C, float count, const float* gamma, const float* beta, float eps,
                               float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                               float* mean, float* rstd, void* stream) {
  CDM_CHECK_ARG(sums && gamma && beta && scale && shift && mean && rstd && C > 0 && count > 0);
  CDM_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr));
  int rc = check_device();
  if (rc) return rc;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(sums, C, count, gamma, beta, eps, momentum, running_mean,
                                                             running_var, scale, shift, mean, rstd);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_bn_apply(const cdm_bn_apply_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->z && a->y && a->P > 0 && a->C % 8 == 0 && 256 % (a->C / 8) == 0);
  CDM_CHECK_ARG((a->scale && a->shift) || a->sums);
  if (a->sums)
    CDM_CHECK_ARG(a->gamma && a->beta && a->count > 0 && a->scale_out && a->shift_out && a->mean_out && a->rstd_out &&
                  (a->running_mean == nullptr) == (a->running_var == nullptr));
  CDM_CHECK_ARG(!a->sc_x || (a->sc_w && a->sc_b));
  CDM_CHECK_ARG(!a->film_scale || (a->film_shift && a->yf && a->px_per_img > 0 && a->film_rows >= 1));
  int rc = check_device();
  if (rc) return rc;
  BnApplyP p{(const bf16*)a->z, a->P, a->C, a->relu, a->scale, a->shift, (bf16*)a->y, a->sc_x, a->sc_w, a->sc_b,
             a->film_scale, a->film_shift, a->film_rows, a->px_per_img, (bf16*)a->yf,
             a->sums, a->gamma, a->beta, a->count, a->eps, a->momentum, a->running_mean, a->running_var,
             a->scale_out, a->shift_out, a->mean_out, a->rstd_out};
  const int g_apply = grid1d(a->P * (a->C / 8) / 4, 256, num_sms() * 4);  // four rows per thread, four blocks per SM
  if (a->sc_x || a->film_scale)
    bn_apply_kernel<true><<<g_apply, 256, 0, ST(stream)>>>(p);
  else
    bn_apply_kernel<false><<<g_apply, 256, 0, ST(stream)>>>(p);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_bn_bwd_apply(const cdm_bn_bwd_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->dy && a->z && a->scale && a->shift && a->mean && a->rstd && a->sums && a->dz);
  CDM_CHECK_ARG(a->P > 0 && a->C % 8 == 0 && 256 % (a->C / 8) == 0 && a->lddy >= a->C && a->count > 0);
  int rc = check_device();
  if (rc) return rc;
  BnBwdP p{(const bf16*)a->dy, a->lddy, (const bf16*)a->z, a->P, a->C, a->relu, a->scale, a->shift, a->mean, a->rstd,
           a->sums, a->count, (bf16*)a->dz};
  bn_bwd_apply_kernel<<<grid1d(a->P * (a->C / 8) / 4, 256, num_sms() * 3), 256, 0, ST(stream)>>>(p);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_maxpool2_fwd(const void* y, int n_img, int H, int W, int C, void* out, void* stream) {
  CDM_CHECK_ARG(y && out && n_img > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0);
  int rc = check_device();
  if (rc) return rc;
  maxpool2_fwd_kernel<<<grid1d((long long)n_img * (H / 2) * (W / 2) * (C / 8)), 256, 0, ST(stream)>>>(
      (const bf16*)y, n_img, H, W, C, (bf16*)out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_maxpool2_bwd(const void* dpool, int lddp, const void* y, int n_img, int H, int W, int C, void* dy,
                                void* stream) {
  CDM_CHECK_ARG(dpool && y && dy && n_img > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && lddp >= C);
  int rc = check_device();
  if (rc) return rc;
  maxpool2_bwd_kernel<<<grid1d((long long)n_img * (H / 2) * (W / 2) * (C / 8)), 256, 0, ST(stream)>>>(
      (const bf16*)dpool, lddp, (const bf16*)y, n_img, H, W, C, (bf16*)dy);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_add_bf16(void* a, int lda, const void* b, int ldb, long long P, int C, void* stream) {
  CDM_CHECK_ARG(a && b && P > 0 && C % 8 == 0 && lda >= C && ldb >= C);
  int rc = check_device();
  if (rc) return rc;
  add_bf16_kernel<<<grid1d(P * (C / 8)), 256, 0, ST(stream)>>>((bf16*)a, lda, (const bf16*)b, ldb, P, C);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_space_to_depth(const void* dv, int n_img, int H, int W, int C, void* out, void* stream) {
  CDM_CHECK_ARG(dv && out && n_img > 0 && H > 0 && W > 0 && C % 8 == 0);
  int rc = check_device();
  if (rc) return rc;
  space_to_depth_kernel<<<grid1d((long long)n_img * H * W * 4 * (C / 8)), 256, 0, ST(stream)>>>(
      (const bf16*)dv, n_img, H, W, C, (bf16*)out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_film_bwd(const void* dyf, int lddyf, const void* y, int n_img, int px, int C, const float* fs,
                            void* dy, float* dfs, float* dfb, void* stream) {
  CDM_CHECK_ARG(dyf && y && fs && dy && dfs && dfb && n_img > 0 && px > 0 && C % 8 == 0 && 256 % (C / 8) == 0);
  int rc = check_device();
  if (rc) return rc;
  film_bwd_kernel<<<n_img, 256, 0, ST(stream)>>>((const bf16*)dyf, lddyf, (const bf16*)y, px, C, fs, (bf16*)dy, dfs,
                                                 dfb);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_gn_bwd(const cdm_gn_bwd_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->x && a->dyf && a->mean_rstd && a->gamma && a->beta && a->dx && a->dgamma_nc && a->dbeta_nc);
  CDM_CHECK_ARG(a->n_img > 0 && a->P > 0 && a->groups > 0 && a->C % a->groups == 0);
  const int cpg = a->C / a->groups;
  CDM_CHECK_ARG(cpg % 8 == 0 && cpg <= 32 && 256 % (cpg / 8) == 0 && a->lddyf >= a->C);
  CDM_CHECK_ARG((a->film_scale == nullptr) == (a->dfs == nullptr) && (a->dfs == nullptr) == (a->dfb == nullptr));
  int rc = check_device();
  if (rc) return rc;
  GnBwdP p{(const bf16*)a->x, (const bf16*)a->dyf, a->lddyf, a->P, a->C, a->groups, a->mean_rstd, a->gamma, a->beta,
           a->film_scale, (bf16*)a->dx, a->dgamma_nc, a->dbeta_nc, a->dfs, a->dfb};
  if (a->n_img >= 128 && (a->C == 128 || a->C == 256) && a->groups == 8)
    gn_bwd_img_kernel<<<a->n_img, 256, 0, ST(stream)>>>(p);  // one block per image once that fills the machine
  else
    gn_bwd_kernel<<<a->n_img * a->groups, 256, 0, ST(stream)>>>(p);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_rows_sum(const float* in, int rows, int C, float* out, void* stream) {
  CDM_CHECK_ARG(in && out && rows > 0 && C > 0);
  int rc = check_device();
  if (rc) return rc;
  rows_sum_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(in, rows, C, out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_avgpool_gelu_train(const void* src, int n_img, int P, int C, float* pre, void* out, void* stream) {
  CDM_CHECK_ARG(src && pre && out && n_img > 0 && P > 0 && C > 0);
  int rc = check_device();
  if (rc) return rc;
  avgpool_pre_kernel<<<n_img, 256, 0, ST(stream)>>>((const bf16*)src, P, C, pre, (bf16*)out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_avgpool_gelu_bwd(const float* pre, const float* dh, int n_img, int P, int C, void* dx,
                                    void* stream) {
  CDM_CHECK_ARG(pre && dh && dx && n_img > 0 && P > 0 && C % 8 == 0);
  int rc = check_device();
  if (rc) return rc;
  int slices = (num_sms() * 4 + n_img - 1) / n_img;  // ~4 blocks per SM in total
  const int max_slices = (P * (C / 8) + 255) / 256;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  avgpool_gelu_bwd_kernel<<<dim3(n_img, slices), 256, 0, ST(stream)>>>(pre, dh, P, C, (bf16*)dx);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_outer_wgrad(const cdm_outer_wgrad_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->s && a->v && a->out && a->workspace && a->n_img > 0 && a->H > 0 && a->W > 0);
  CDM_CHECK_ARG(a->C % 8 == 0 && 256 % (a->C / 8) == 0);
  CDM_CHECK_ARG(!a->mean_rstd || (a->gamma && a->beta));
  int rc = check_device();
  if (rc) return rc;
  int blocks = a->n_img * a->H;  // one image row per block iteration
  if (blocks > num_sms() * 2) blocks = num_sms() * 2;
  if (blocks > a->workspace_blocks) blocks = a->workspace_blocks;
  if (blocks < 1) blocks = 1;
  OuterWgradP p{a->s, (const bf16*)a->v, a->n_img, a->H, a->W, a->C, a->flip, a->mean_rstd, a->gamma, a->beta,
                a->workspace};
  outer_wgrad_kernel<<<blocks, 256, 0, ST(stream)>>>(p);
  CDM_CHECK_LAUNCH();
  chan_reduce_final_kernel<<<(9 * a->C * 32 + 255) / 256, 256, 0, ST(stream)>>>(a->workspace, blocks, 9 * a->C, a->out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_embed_bwd(const cdm_embed_bwd_args* a, void* stream) {
  CDM_CHECK_ARG(a && a->in && a->w1 && a->b1 && a->w2 && a->dout && a->pre && a->h && a->dpre);
  CDM_CHECK_ARG(a->dw1 && a->db1 && a->dw2 && a->db2 && a->rows > 0 && a->din > 0 && a->emb > 0 && a->emb <= 4096);
  int rc = check_device();
  if (rc) return rc;
  embed_hidden_kernel<<<a->rows, 256, 0, ST(stream)>>>(a->in, a->din, a->w1, a->b1, a->emb, a->pre, a->h);
  CDM_CHECK_LAUNCH();
  embed_bwd_w2_kernel<<<a->emb, 256, 0, ST(stream)>>>(a->dout, a->h, a->rows, a->emb, a->dw2, a->db2);
  CDM_CHECK_LAUNCH();
  embed_bwd_hidden_kernel<<<a->rows, 256, a->emb * sizeof(float), ST(stream)>>>(a->dout, a->w2, a->pre, a->emb,
                                                                               a->dpre);
  CDM_CHECK_LAUNCH();
  embed_bwd_w1_kernel<<<(a->emb + 127) / 128, 128, 0, ST(stream)>>>(a->dpre, a->in, a->rows, a->din, a->emb, a->dw1,
                                                                   a->db1);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_mse_grad(const float* pred, const float* target, long long n, float inv_count, float* dpred,
                            float* partial, int partial_blocks, float* loss_sum, void* stream) {
  CDM_CHECK_ARG(pred && target && dpred && partial && loss_sum && n > 0 && n % 4 == 0 && partial_blocks > 0);
  int rc = check_device();
  if (rc) return rc;
  int blocks = grid1d(n / 4);
  if (blocks > partial_blocks) blocks = partial_blocks;
  mse_grad_kernel<<<blocks, 256, 0, ST(stream)>>>(pred, target, n / 4, inv_count, dpred, partial);
  CDM_CHECK_LAUNCH();
  chan_reduce_final_kernel<<<1, 32, 0, ST(stream)>>>(partial, blocks, 1, loss_sum);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
extern "C" int cdm_adam_step(const void* table, int n_tensors, long long max_numel, float lr, float beta1, float beta2,
                             float eps, int step, const float* lr_dev, const int* step_dev, void* stream) {
  CDM_CHECK_ARG(table && n_tensors > 0 && max_numel > 0 && (step >= 1 || step_dev));
  int rc = check_device();
  if (rc) return rc;
  const float bc1 = 1.f - powf(beta1, (float)(step < 1 ? 1 : step)), bc2 = 1.f - powf(beta2, (float)(step < 1 ? 1 : step));
  CDM_CHECK_ARG(n_tensors <= kAdamMaxTensors);
  adam_kernel<<<num_sms() * 8, 256, 0, ST(stream)>>>((const AdamTensor*)table, n_tensors, lr, beta1, beta2, eps, bc1,
                                                     bc2, lr_dev, step_dev);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
