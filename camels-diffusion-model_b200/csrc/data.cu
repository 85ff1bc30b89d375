// Data preparation of the training / evaluation scripts on the device:
// code/train_diffusion_paper.py:232-262 — parameter min-max normalisation and the map pipeline
//   shift to positive (if min <= 0) -> / max -> log10 -> min-max to [0,1] -> F.interpolate(size=64, bilinear).
// Every step of the map pipeline is monotonic, so the extrema after each step are the images of the raw
// extrema: ONE min/max reduction over the raw maps gives every normalisation constant, and the second pass
// reads only the input pixels the bilinear taps touch (for 256 -> 64: the central 2x2 of each 4x4 block),
// applies the scalar chain to them and writes the resized map.  All fp32, each operation rounded separately
// in the reference's order (numpy float32 semantics); log10f differs from numpy's by <= 1 ulp.
#include <float.h>
#include <math.h>

#include "common.h"

namespace cdm {

__device__ __forceinline__ void warp_minmax(float& lo, float& hi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
}
__device__ __forceinline__ void block_minmax(float& lo, float& hi, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  warp_minmax(lo, hi);
  __syncthreads();
  if (lane == 0) {
    red[2 * warp] = lo;
    red[2 * warp + 1] = hi;
  }
  __syncthreads();
  lo = FLT_MAX;
  hi = -FLT_MAX;
  for (int i = 0; i < nw; ++i) {
    lo = fminf(lo, red[2 * i]);
    hi = fmaxf(hi, red[2 * i + 1]);
  }
}

// stage 1: per-block extrema (16-byte loads); stage 2 (last block, via a ticket) folds them into out[0..1].
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, long long n, float* __restrict__ partial,
                                                     unsigned int* __restrict__ ticket, float* __restrict__ out) {
  __shared__ float red[64];
  __shared__ bool last;
  float lo = FLT_MAX, hi = -FLT_MAX;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
    hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    lo = fminf(lo, x[i]);
    hi = fmaxf(hi, x[i]);
  }
  block_minmax(lo, hi, red);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = lo;
    partial[2 * blockIdx.x + 1] = hi;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  lo = FLT_MAX;
  hi = -FLT_MAX;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    lo = fminf(lo, partial[2 * i]);
    hi = fmaxf(hi, partial[2 * i + 1]);
  }
  block_minmax(lo, hi, red);
  if (threadIdx.x == 0) {
    out[0] = lo;
    out[1] = hi;
    *ticket = 0;  // re-armed for the next launch (graph replay)
  }
}

// The scalar chain of train_diffusion_paper.py:254-260 applied to one value.
struct MapNorm {
  float mn, mx1, lmin, lrange;
  bool shift;
  __device__ __forceinline__ float pos(float v) const { return shift ? __fadd_rn(__fsub_rn(v, mn), 1e-8f) : v; }
  __device__ __forceinline__ float operator()(float v) const {
    const float l = log10f(__fdiv_rn(pos(v), mx1));
    return __fdiv_rn(__fsub_rn(l, lmin), lrange);
  }
  __device__ MapNorm(float raw_min, float raw_max) {
    mn = raw_min;
    shift = raw_min <= 0.f;
    mx1 = pos(raw_max);
    lmin = log10f(__fdiv_rn(pos(raw_min), mx1));
    const float lmax = log10f(__fdiv_rn(mx1, mx1));
    lrange = __fsub_rn(lmax, lmin);
  }
};

// out[n][Ho][Wo] = bilinear(align_corners=False) of norm(in[n][Hi][Wi]); torch's source-index rule.
__global__ void __launch_bounds__(256) preprocess_maps_kernel(const float* __restrict__ in, int n, int Hi, int Wi,
                                                              const float* __restrict__ raw_minmax, int Ho, int Wo,
                                                              float* __restrict__ out) {
  const MapNorm f(raw_minmax[0], raw_minmax[1]);
  const float sy = (float)Hi / (float)Ho, sx = (float)Wi / (float)Wo;
  const long long total = (long long)n * Ho * Wo;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(o % Wo), oy = (int)((o / Wo) % Ho);
    const long long img = o / ((long long)Wo * Ho);
    const float fy = fmaxf(sy * ((float)oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)ox + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* p = in + img * Hi * Wi;
    const float a = f(p[(long long)y0 * Wi + x0]), b = f(p[(long long)y0 * Wi + x1]);
    const float c = f(p[(long long)y1 * Wi + x0]), d = f(p[(long long)y1 * Wi + x1]);
    out[o] = __fadd_rn(__fmul_rn(ly0, __fadd_rn(__fmul_rn(lx0, a), __fmul_rn(lx1, b))),
                       __fmul_rn(ly1, __fadd_rn(__fmul_rn(lx0, c), __fmul_rn(lx1, d))));
  }
}

// Parameter table: per-column min / max over the rows, out = (x - min) / (max - min + 1e-8), columns cut or
// zero-padded to out_cols, every input row written `repeat` times (np.repeat(param_data, 15, axis=0)).
__global__ void __launch_bounds__(256) normalize_params_kernel(const float* __restrict__ x, int rows, int cols,
                                                               int repeat, int out_cols, float* __restrict__ out,
                                                               float* __restrict__ col_min, float* __restrict__ col_max) {
  __shared__ float red[64];
  const int c = blockIdx.x;  // one block per OUTPUT column
  if (c >= cols) {
    for (long long r = threadIdx.x; r < (long long)rows * repeat; r += blockDim.x) out[r * out_cols + c] = 0.f;
    return;
  }
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float v = x[(long long)r * cols + c];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  block_minmax(lo, hi, red);
  if (threadIdx.x == 0) {
    col_min[c] = lo;
    col_max[c] = hi;
  }
  if (c >= out_cols) return;
  const float den = __fadd_rn(__fsub_rn(hi, lo), 1e-8f);
  for (long long r = threadIdx.x; r < (long long)rows * repeat; r += blockDim.x)
    out[r * out_cols + c] = __fdiv_rn(__fsub_rn(x[(r / repeat) * cols + c], lo), den);
}

}  // namespace cdm

using namespace cdm;

extern "C" int cdm_minmax(const float* x, long long n, float* workspace, int workspace_floats, float* out, void* stream) {
  CDM_CHECK_ARG(x && workspace && out && n > 0 && workspace_floats >= 2 * 8 + 1);
  CDM_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  int rc = check_device();
  if (rc) return rc;
  long long want = (n / 4 + 255) / 256;
  int blocks = (workspace_floats - 1) / 2;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (want < blocks) blocks = (int)(want < 1 ? 1 : want);
  // the last workspace float is the ticket counter: the caller zero-fills the workspace once
  minmax_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      x, n, workspace, reinterpret_cast<unsigned int*>(workspace + workspace_floats - 1), out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_preprocess_maps(const float* in, int n, int Hi, int Wi, const float* raw_minmax, int Ho, int Wo,
                                   float* out, void* stream) {
  CDM_CHECK_ARG(in && raw_minmax && out && n > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0);
  int rc = check_device();
  if (rc) return rc;
  const long long total = (long long)n * Ho * Wo;
  long long blocks = (total + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  preprocess_maps_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, n, Hi, Wi, raw_minmax, Ho, Wo, out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_normalize_params(const float* x, int rows, int cols, int repeat, int out_cols, float* out,
                                    float* col_min, float* col_max, void* stream) {
  CDM_CHECK_ARG(x && out && col_min && col_max && rows > 0 && cols > 0 && repeat > 0 && out_cols > 0);
  int rc = check_device();
  if (rc) return rc;
  normalize_params_kernel<<<cols > out_cols ? cols : out_cols, 256, 0, (cudaStream_t)stream>>>(
      x, rows, cols, repeat, out_cols, out, col_min, col_max);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
