// Map statistics of the generated samples, computed where the samples live:
//   * radial power spectrum P(k) of every map  — power_spectrum, code/diffusion_utilities.py:302-368,
//     called per image by compare_power_spectra (:370-431) and sample_power_spectra.py;
//   * per-image pixel histograms               — compare_distributions, code/train_diffusion_paper.py:861-876
//     (np.histogram(image.ravel(), bins, density=True) per image).
// The reference loops over pixels in Python (one interpreter iteration per Fourier mode); here one CTA
// handles one map: a 2-D DFT as two passes of an NxN twiddle product in shared memory (N <= 64, fp32
// data as numpy's complex64 FFT, twiddles from sincospi), |F|^2, and a fixed-order fp64 reduction of the
// modes of each radial bin (bin membership comes from the caller as a CSR list, so the binning rule is the
// reference's own `int(round(k/dk))` evaluated on the host, not re-derived here).
#include <math.h>

#include "common.h"

namespace cdm {

constexpr int kPsMaxN = 64;

__global__ void __launch_bounds__(256) power_spectrum_kernel(const float* __restrict__ maps, int N,
                                                             const int* __restrict__ bin_start,
                                                             const int* __restrict__ bin_items, int n_bins,
                                                             double scale, double* __restrict__ pk) {
  extern __shared__ float ps_smem[];
  float* s_x = ps_smem;             // [N][N]   input map, later |F|^2
  float* s_re = s_x + N * N;        // [N][N+1] row-transformed, real
  float* s_im = s_re + N * (N + 1); // [N][N+1] row-transformed, imaginary
  float* s_c = s_im + N * (N + 1);  // [N] cos(2 pi j / N)
  float* s_s = s_c + N;             // [N] sin(2 pi j / N)
  const float* x = maps + (size_t)blockIdx.x * N * N;
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) s_x[i] = x[i];
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float s, c;
    sincospif(2.f * (float)j / (float)N, &s, &c);
    s_c[j] = c;
    s_s[j] = s;
  }
  __syncthreads();
  const int mask = N - 1;  // N is a power of two
  // pass 1: Y[r][k] = sum_c x[r][c] * exp(-2 pi i c k / N)
  for (int o = threadIdx.x; o < N * N; o += blockDim.x) {
    const int r = o / N, k = o - r * N;
    float re = 0.f, im = 0.f;
    for (int c = 0; c < N; ++c) {
      const int j = (c * k) & mask;
      const float v = s_x[r * N + c];
      re = fmaf(v, s_c[j], re);
      im = fmaf(-v, s_s[j], im);
    }
    s_re[r * (N + 1) + k] = re;
    s_im[r * (N + 1) + k] = im;
  }
  __syncthreads();
  // pass 2: F[l][k] = sum_r Y[r][k] * exp(-2 pi i r l / N);  power = |F|^2 / N^2 (norm="ortho")
  const float inv = 1.f / ((float)N * (float)N);
  for (int o = threadIdx.x; o < N * N; o += blockDim.x) {
    const int l = o / N, k = o - l * N;
    float re = 0.f, im = 0.f;
    for (int r = 0; r < N; ++r) {
      const int j = (r * l) & mask;
      const float yr = s_re[r * (N + 1) + k], yi = s_im[r * (N + 1) + k], c = s_c[j], s = s_s[j];
      re = fmaf(yr, c, fmaf(yi, s, re));
      im = fmaf(yi, c, fmaf(-yr, s, im));
    }
    s_x[o] = (re * re + im * im) * inv;
  }
  __syncthreads();
  // radial bins: one warp per bin, fixed order, fp64 (the reference accumulates into np.zeros = float64)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int b = warp; b < n_bins; b += nw) {
    const int i0 = bin_start[b], i1 = bin_start[b + 1];
    double s = 0.0;
    for (int i = i0 + lane; i < i1; i += 32) s += (double)s_x[bin_items[i]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) pk[(size_t)blockIdx.x * n_bins + b] = (i1 > i0 ? s / (double)(i1 - i0) : 0.0) * scale;
  }
}

// np.histogram(values, edges): bin i holds edges[i] <= v < edges[i+1], the last bin is closed on the right;
// comparisons in fp64 (numpy promotes the float32 pixels to the float64 edges).  Integer counts: exact.
__global__ void __launch_bounds__(256) pixel_histogram_kernel(const float* __restrict__ maps, int P,
                                                              const double* __restrict__ edges, int n_bins,
                                                              int* __restrict__ counts) {
  extern __shared__ int hist_smem[];
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x) hist_smem[i] = 0;
  __syncthreads();
  const float* x = maps + (size_t)blockIdx.x * P;
  const double lo = edges[0], hi = edges[n_bins];
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const double v = (double)x[i];
    if (!(v >= lo && v <= hi)) continue;  // out of range (or NaN): not counted, as numpy
    int a = 0, b = n_bins;                // invariant: edges[a] <= v, and (b == n_bins or v < edges[b])
    while (b - a > 1) {
      const int m = (a + b) >> 1;
      if (edges[m] <= v) a = m; else b = m;
    }
    atomicAdd(&hist_smem[a], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x) counts[(size_t)blockIdx.x * n_bins + i] = hist_smem[i];
}

}  // namespace cdm

using namespace cdm;

extern "C" int cdm_power_spectrum(const float* maps, int n_maps, int N, const int* bin_start, const int* bin_items,
                                  int n_bins, double scale, double* pk, void* stream) {
  CDM_CHECK_ARG(maps && bin_start && bin_items && pk && n_maps > 0 && n_bins > 0);
  CDM_CHECK_ARG(N >= 2 && N <= kPsMaxN && (N & (N - 1)) == 0);
  int rc = check_device();
  if (rc) return rc;
  const int smem = (N * N + 2 * N * (N + 1) + 2 * N) * (int)sizeof(float);
  static bool attr_set[64];  // per device: cudaFuncSetAttribute is not process-wide
  int dev = 0;
  CDM_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    CDM_CHECK_CUDA(cudaFuncSetAttribute(power_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  power_spectrum_kernel<<<n_maps, 256, smem, (cudaStream_t)stream>>>(maps, N, bin_start, bin_items, n_bins, scale, pk);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

extern "C" int cdm_pixel_histogram(const float* maps, int n_maps, int P, const double* edges, int n_bins, int* counts,
                                   void* stream) {
  CDM_CHECK_ARG(maps && edges && counts && n_maps > 0 && P > 0 && n_bins > 0 && n_bins <= 12 * 1024);
  int rc = check_device();
  if (rc) return rc;
  pixel_histogram_kernel<<<n_maps, 256, n_bins * sizeof(int), (cudaStream_t)stream>>>(maps, P, edges, n_bins, counts);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
