// Prepared launches of the tcgen05 kernels: everything cdm_conv3x3 / cdm_gemm derive from their arguments (tensor
// maps, kernel variant, grid, kernel parameters), kept so that a cdm_plan can encode them once (api_igemm.cu).
#pragma once
#include "common.h"
#include "kparams.h"

namespace cdm {

enum { kConvSw32 = 3, kConvSw8 = 4 };  // + 2: the TMA-store epilogue variants (5, 6); 0-2: conv3x3_kernel<MODE>

struct ConvLaunch {
  CUtensorMap a0, a1, b, out;
  ConvKParams p;
  int variant, grid;
  int bn_fold;  // CDM_EPI_BNSTATS: fold p.bn_partial (+ the ranks of xr) into bn_sums after the convolution
  float* bn_sums;
  const cdm_xrank* xr;
};
int conv_prepare(const cdm_conv3x3_args* a, ConvLaunch* L);
int conv_launch(const ConvLaunch& L, cudaStream_t st);

struct GemmLaunch {
  CUtensorMap a0, a1, b;
  CUtensorMap out;    // variant 1, out_mode 1: the pixel-shuffled output as [img*H + h][kh][w][kw][128 ch] (TMA store)
  GemmKParams p;      // variant 0: streaming gemm_kernel (+ split-K reduce)
  GemmBresKParams q;  // variant 1: resident-weight gemm_bres_kernel
  int variant, grid, M;
};
int gemm_prepare(const cdm_gemm_args* a, GemmLaunch* G);
int gemm_launch(const GemmLaunch& G, cudaStream_t st);

}  // namespace cdm
