/* A non-Python consumer of the composite C ABI (include/cdm_b200.h): one eval forward of ContextUnet
 * (ContextUnet.py:42-60) from plain C — cudart for device memory, libcdm_b200.so for everything else.
 *
 *   plan_forward <weights.bin> <inputs.bin> <eps_out.bin> <n_cfeat> <batch>
 *
 * weights.bin: the cdm_plan_n_tensors() fp32 tensors in cdm_plan_tensor_name() order (= state_dict order without the
 * num_batches_tracked entries), PyTorch layouts, concatenated.  inputs.bin: x [B][64][64], t [1], c [B][n_cfeat],
 * shortcut [2][128] (the fresh 1x1 conv of this call: w_c then b_c), all fp32.  Writes eps [B][64][64] fp32.
 * Built and run by tests/test_gpu_abi.py, which compares eps with ContextUnet.forward of the Python module bit for bit. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "cdm_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                     \
      return 10;                                                                   \
    }                                                                              \
  } while (0)
#define CDM(x)                                                                     \
  do {                                                                             \
    int rc_ = (x);                                                                 \
    if (rc_ != CDM_OK) {                                                           \
      fprintf(stderr, "%s -> %d: %s\n", #x, rc_, cdm_last_error());                \
      return 11;                                                                   \
    }                                                                              \
  } while (0)

static float* upload(FILE* f, long long n) {
  float* h = (float*)malloc((size_t)n * 4);
  float* d = NULL;
  if (!h || fread(h, 4, (size_t)n, f) != (size_t)n) return NULL;
  if (cudaMalloc((void**)&d, (size_t)n * 4) != cudaSuccess) return NULL;
  if (cudaMemcpy(d, h, (size_t)n * 4, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
  free(h);
  return d;
}

int main(int argc, char** argv) {
  if (argc != 6) return 2;
  const int ncf = atoi(argv[4]), B = atoi(argv[5]);
  FILE* fw = fopen(argv[1], "rb");
  FILE* fi = fopen(argv[2], "rb");
  if (!fw || !fi) return 3;
  CDM(cdm_device_ok());
  const int nt = cdm_plan_n_tensors();
  const float** w = (const float**)malloc(sizeof(float*) * (size_t)nt);
  for (int i = 0; i < nt; ++i) {
    w[i] = upload(fw, cdm_plan_tensor_numel(i, ncf));
    if (!w[i]) {
      fprintf(stderr, "tensor %d (%s): short read or allocation failure\n", i, cdm_plan_tensor_name(i));
      return 4;
    }
  }
  float* x = upload(fi, (long long)B * 64 * 64);
  float* t = upload(fi, 1);
  float* c = upload(fi, (long long)B * ncf);
  float* sc = upload(fi, 2 * 128);
  if (!x || !t || !c || !sc) return 5;

  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cdm_plan_desc d;
  d.n_cfeat = ncf, d.batch = B, d.reps = 1, d.tensors = w, d.conv_mode = 0;
  d.arena_bytes = cdm_plan_arena_bytes(ncf);
  d.workspace_bytes = cdm_plan_workspace_bytes(B, 1);
  CK(cudaMalloc(&d.arena, (size_t)d.arena_bytes));
  CK(cudaMalloc(&d.workspace, (size_t)d.workspace_bytes));
  cdm_plan* plan = NULL;
  CDM(cdm_plan_create(&d, st, &plan));

  float *cemb1, *temb1, *cemb2, *temb2, *eps;
  CK(cudaMalloc((void**)&cemb1, (size_t)B * 256 * 4));
  CK(cudaMalloc((void**)&temb1, 256 * 4));
  CK(cudaMalloc((void**)&cemb2, (size_t)B * 128 * 4));
  CK(cudaMalloc((void**)&temb2, 128 * 4));
  CK(cudaMalloc((void**)&eps, (size_t)B * 64 * 64 * 4));
  CDM(cdm_plan_embed(plan, 0, c, B, cemb1, st)); /* contextembed1(c)  ContextUnet.py:51 */
  CDM(cdm_plan_embed(plan, 1, t, 1, temb1, st)); /* timeembed1(t)     :52 */
  CDM(cdm_plan_embed(plan, 2, c, B, cemb2, st)); /* contextembed2(c)  :53 */
  CDM(cdm_plan_embed(plan, 3, t, 1, temb2, st)); /* timeembed2(t)     :54 */

  cdm_forward_args f;
  f.x = x, f.sc_tab = sc, f.cemb1 = cemb1, f.temb1 = temb1, f.cemb2 = cemb2, f.temb2 = temb2;
  f.temb_rows = 1, f.step_ptr = NULL, f.eps = eps;
  CDM(cdm_forward_eval(plan, &f, st));
  CDM(cdm_forward_eval(plan, &f, st)); /* a plan is reusable: the second call must give the same bits */
  CK(cudaStreamSynchronize(st));

  float* h = (float*)malloc((size_t)B * 64 * 64 * 4);
  CK(cudaMemcpy(h, eps, (size_t)B * 64 * 64 * 4, cudaMemcpyDeviceToHost));
  FILE* fo = fopen(argv[3], "wb");
  if (!fo || fwrite(h, 4, (size_t)B * 64 * 64, fo) != (size_t)B * 64 * 64) return 6;
  fclose(fo);
  cdm_plan_destroy(plan);
  printf("plan_forward ok: %d tensors, %d launches per forward, batch %d, n_cfeat %d\n", nt, cdm_plan_n_launches(), B, ncf);
  return 0;
}
