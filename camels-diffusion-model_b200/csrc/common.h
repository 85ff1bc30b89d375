// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cdm_b200.h"

namespace cdm {

void set_error(const char* fmt, ...);
int check_device();  // CDM_OK or CDM_ERR_ARCH / CDM_ERR_CUDA
int num_sms();       // multiprocessor count of the current device (cached per device)

#define CDM_CHECK_ARG(cond)                                             \
  do {                                                                  \
    if (!(cond)) {                                                      \
      cdm::set_error("%s: argument check failed: %s", __func__, #cond); \
      return CDM_ERR_ARG;                                               \
    }                                                                   \
  } while (0)

#define CDM_CHECK_CUDA(expr)                                                               \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      cdm::set_error("%s: %s failed: %s", __func__, #expr, cudaGetErrorString(_e));       \
      return CDM_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define CDM_CHECK_LAUNCH() CDM_CHECK_CUDA(cudaGetLastError())

// Encode a tiled, 128B-swizzled bf16 tensor map (rank 2 or 4). dims/box are
// innermost-first; strides_bytes has rank-1 entries (dim 1 ..).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

// Final pass of a two-stage reduction (+ cross-rank exchange over peer memory when xr->world > 1); train.cu.
int launch_xrank_sum(const float* partial, int n_blocks, int n, float* out, const cdm_xrank* xr, cudaStream_t st);

}  // namespace cdm
