// Measurement / bring-up probe, -DCDM_PROBES build only (libcdm_b200_probes.so, tools/gpu_probe.py tmem_layout):
// which TMEM element each register of `tcgen05.ld.16x256b.x4` receives.  One CTA writes lane*1000 + column into 32
// columns with tcgen05.st.32x32b.x32 and reads them back with the 16-lane shape; the host checks the mapping the
// fragment-shaped epilogue experiments rely on.
#ifdef CDM_PROBES
#include "common.h"
#include "ptx.cuh"

namespace cdm {

__global__ void __launch_bounds__(128) probe_tmem_layout_kernel(float* __restrict__ out) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_ptr + ((uint32_t)(warp * 32) << 16);
  {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (float)((warp * 32 + lane) * 1000 + i);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(base),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]),
        "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]),
        "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // lanes 0-15 and 16-31 of this warp's quadrant
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(base + ((uint32_t)(h * 16) << 16))
        : "memory");
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) out[((warp * 2 + h) * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_ptr, 32);
}

}  // namespace cdm

// out: fp32 [4 warps][2 halves][32 threads][16 registers] = lane * 1000 + column of the element each register got
extern "C" int cdm_probe_tmem_layout(float* out, void* stream) {
  CDM_CHECK_ARG(out != nullptr);
  int rc = cdm::check_device();
  if (rc) return rc;
  cdm::probe_tmem_layout_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}
#endif
