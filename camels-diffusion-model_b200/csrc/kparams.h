// Kernel parameter blocks of the tcgen05 kernels (shared by igemm.cuh and the prepared launches of launch.h).
#pragma once
#include <cuda_bf16.h>

namespace cdm {

typedef __nv_bfloat16 bf16;

struct ConvKParams {
  int H, W, n_img;
  int chunks0, chunks;  // 64-channel K chunks from src0 / in total
  int n_tiles, cout;
  int strips_x, strips_y;
  int n_units;
  int flags;
  const float* scale;
  const float* shift;
  bf16* out;
  const float* sc_x;
  const float* sc_tab;
  int sc_reps;
  const float* film_scale;
  const float* film_shift;
  int film_shift_rows;
  const int* step_ptr;
  float* gn_partial;
  float* bn_partial;  // CDM_EPI_BNSTATS: [gridDim.x][2][cout]
  float res_scale;    // CDM_EPI_RESSCALE
  // CDM_EPI_BNBWD: the launch is the data gradient dy of a Conv-BatchNorm-ReLU layer whose forward z (bf16, same
  // NHWC shape as `out`), scale / shift (ReLU mask: z*scale+shift > 0), batch mean and rstd are given; the epilogue also
  // accumulates sum g and sum g*xhat per channel (g = stored dy under the mask) into bn_partial
  const bf16* bwd_z;
  const float* bwd_scale;
  const float* bwd_shift;
  const float* bwd_mean;
  const float* bwd_rstd;
};

struct GemmKParams {
  int M, N;
  int chunks0, chunks;
  int m_tiles, n_tiles, n_units;
  const float* shift;
  int shift_mod;
  int out_mode, H, W, h_shift, w_shift;
  bf16* out;
  int k_split;     // > 1: unit = (K slice, m_tile, n_tile); each slice stores its fp32 tile to `partial`
  float* partial;  // [k_split][m_tiles*128][N]; gemm_splitk_reduce_kernel adds the slices in order (+ shift -> bf16)
};

struct GemmBresKParams {
  int M, N;
  int chunks0, chunks;
  int m_tiles, n_res, n_groups;
  const float* shift;
  int shift_mod;
  int out_mode, H, W, h_shift, w_shift;
  bf16* out;
};

}  // namespace cdm
