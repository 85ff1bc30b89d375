// cdm_plan: the eval-mode ContextUnet forward (ContextUnet.py:42-60) and one reverse-diffusion step
// (code/train_diffusion_paper.py:594-618) as SINGLE C calls.
//
// cdm_plan_create packs the model's fp32 PyTorch-layout parameters into the tensor-core layouts once (bf16 K-major
// weights, eval BatchNorm folded to per-channel scale / shift), carves the activation workspace, and encodes every
// layer's tensor maps and kernel parameters (conv_prepare / gemm_prepare).  cdm_forward_eval is then 26 kernel
// launches and nothing else: no descriptor encoding, no allocation, no host synchronisation — graph-capturable.
// The library allocates no DEVICE memory: arena and workspace belong to the caller.
#include <string.h>

#include <string>
#include <vector>

#include "common.h"
#include "launch.h"

namespace cdm {

constexpr int kNF = 128, kH = 64;  // the configuration BASELINE.json names: n_feat = 128, height = 64

// ------------------------------------------------------------------------------------------------ tensor table
struct TensorInfo {
  std::string name;
  long long numel;  // for n_cfeat = 1; + cf_scale * (n_cfeat - 1)
  long long cf_scale;
};

static const std::vector<TensorInfo>& tensor_table() {
  static std::vector<TensorInfo> t;
  if (!t.empty()) return t;
  auto add = [&](const std::string& n, long long numel, long long cf = 0) { t.push_back({n, numel, cf}); };
  auto rcb = [&](const std::string& pre, int cin, int cout) {
    int ci = cin;
    for (int k = 1; k <= 2; ++k) {
      const std::string c = pre + ".conv" + std::to_string(k);
      add(c + ".0.weight", (long long)cout * ci * 9);
      add(c + ".0.bias", cout);
      add(c + ".1.weight", cout);
      add(c + ".1.bias", cout);
      add(c + ".1.running_mean", cout);
      add(c + ".1.running_var", cout);
      ci = cout;
    }
  };
  auto embed = [&](const std::string& pre, int din_fixed, int emb) {  // din_fixed < 0: din = n_cfeat
    add(pre + ".model.0.weight", emb, din_fixed < 0 ? emb : 0);
    add(pre + ".model.0.bias", emb);
    add(pre + ".model.2.weight", (long long)emb * emb);
    add(pre + ".model.2.bias", emb);
  };
  rcb("init_conv", 1, kNF);
  rcb("down1.model.0", kNF, kNF);
  rcb("down1.model.1", kNF, kNF);
  rcb("down2.model.0", kNF, 2 * kNF);
  rcb("down2.model.1", 2 * kNF, 2 * kNF);
  embed("timeembed1", 1, 2 * kNF);
  embed("timeembed2", 1, kNF);
  embed("contextembed1", -1, 2 * kNF);
  embed("contextembed2", -1, kNF);
  add("up0.0.weight", (long long)2 * kNF * 2 * kNF * (kH / 4) * (kH / 4));
  add("up0.0.bias", 2 * kNF);
  add("up0.1.weight", 2 * kNF);
  add("up0.1.bias", 2 * kNF);
  add("up1.model.0.weight", (long long)4 * kNF * kNF * 4);
  add("up1.model.0.bias", kNF);
  rcb("up1.model.1", kNF, kNF);
  rcb("up1.model.2", kNF, kNF);
  add("up2.model.0.weight", (long long)2 * kNF * kNF * 4);
  add("up2.model.0.bias", kNF);
  rcb("up2.model.1", kNF, kNF);
  rcb("up2.model.2", kNF, kNF);
  add("out.0.weight", (long long)kNF * 2 * kNF * 9);
  add("out.0.bias", kNF);
  add("out.1.weight", kNF);
  add("out.1.bias", kNF);
  add("out.3.weight", (long long)kNF * 9);
  add("out.3.bias", 1);
  return t;
}

static int tensor_index(const std::string& name) {
  const auto& t = tensor_table();
  for (size_t i = 0; i < t.size(); ++i)
    if (t[i].name == name) return (int)i;
  return -1;
}

// ------------------------------------------------------------------------------------------------ pack kernels
// dst[i0][i1][i2][i3] (bf16, contiguous) = src[i0*s0 + i1*s1 + i2*s2 + i3*s3] (fp32): every weight permutation of
// the forward (OIHW -> [co][kh][kw][ci], IOHW -> [(kh,kw,co)][ci]) is one instance.
__global__ void __launch_bounds__(256) plan_pack_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                             int d1, int d2, int d3, long long s0, long long s1,
                                                             long long s2, long long s3, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int i3 = (int)(i % d3);
    long long r = i / d3;
    const int i2 = (int)(r % d2);
    r /= d2;
    const int i1 = (int)(r % d1);
    const long long i0 = r / d1;
    dst[i] = __float2bfloat16_rn(src[i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3]);
  }
}
// dst[t][c] = src[c][t] (fp32): the K = 9 / N = 1 convolutions keep fp32 tap-major weights
__global__ void plan_transpose9_kernel(const float* __restrict__ src, float* __restrict__ dst, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 9 * C) dst[i] = src[(i % C) * 9 + i / C];
}
// eval BatchNorm folded into the convolution: scale = gamma / sqrt(var + eps), shift = beta + (bias - mean) * scale,
// every operation rounded separately (what the torch expressions gamma / torch.sqrt(var + eps) etc. give)
__global__ void plan_fold_bn_kernel(const float* __restrict__ bias, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps, int C, float* __restrict__ scale,
                                    float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const float s = __fdiv_rn(gamma[i], __fsqrt_rn(__fadd_rn(var[i], eps)));
  scale[i] = s;
  shift[i] = __fadd_rn(beta[i], __fmul_rn(__fsub_rn(bias[i], mean[i]), s));
}
__global__ void plan_fill_kernel(float* dst, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

}  // namespace cdm

using namespace cdm;

// ------------------------------------------------------------------------------------------------ the plan
struct ConvSlot {
  void* w;       // bf16 [cout][3][3][cin]  (init_conv.conv1: fp32 [9][cout])
  float* scale;  // fp32 [cout]
  float* shift;
  int cin, cout;
};

struct cdm_plan {
  int n_cfeat, batch, reps, n, conv_mode;
  std::vector<const float*> tensors;
  uint8_t* arena;
  uint8_t* ws;
  // packed weights (arena)
  ConvSlot conv[19];  // 18 ResidualConvBlock convolutions in forward order + out.0
  __nv_bfloat16 *up0_w, *up1_w, *up2_w;
  float* out3_w;
  // fp32 vectors used in place (pointers into `tensors`, resolved once)
  const float *up0_b, *up0_g, *up0_beta, *up1_b, *up2_b, *out_g, *out_beta, *out3_b;
  const float* emb[4][4];  // [contextembed1, timeembed1, contextembed2, timeembed2][w1, b1, w2, b2]
  // workspace
  __nv_bfloat16 *x0, *p64, *q64, *d1, *p32w, *q32w, *u1f, *d2, *hidden, *u0raw, *u0f;
  float *gn_partial, *gn_mr, *eps;
  // prepared launches (forward order)
  ConvLaunch cl[19];
  GemmLaunch gl[3];
};

namespace {

constexpr size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

const char* const kConvPrefix[18] = {
    "init_conv.conv1",     "init_conv.conv2",     "down1.model.0.conv1", "down1.model.0.conv2", "down1.model.1.conv1",
    "down1.model.1.conv2", "down2.model.0.conv1", "down2.model.0.conv2", "down2.model.1.conv1", "down2.model.1.conv2",
    "up1.model.1.conv1",   "up1.model.1.conv2",   "up1.model.2.conv1",   "up1.model.2.conv2",   "up2.model.1.conv1",
    "up2.model.1.conv2",   "up2.model.2.conv1",   "up2.model.2.conv2"};
const int kConvCin[19] = {1, 128, 128, 128, 128, 128, 128, 256, 256, 256, 128, 128, 128, 128, 128, 128, 128, 128, 256};
const int kConvCout[19] = {128, 128, 128, 128, 128, 128, 256, 256, 256, 256, 128, 128, 128, 128, 128, 128, 128, 128, 128};

size_t arena_layout(cdm_plan* p, uint8_t* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = base ? base + off : nullptr;
    off += al(bytes);
    return r;
  };
  for (int i = 0; i < 19; ++i) {
    const size_t wbytes = i == 0 ? (size_t)9 * kNF * 4 : (size_t)kConvCout[i] * 9 * kConvCin[i] * 2;
    void* w = take(wbytes);
    float* sc = reinterpret_cast<float*>(take((size_t)kConvCout[i] * 4));
    float* sh = reinterpret_cast<float*>(take((size_t)kConvCout[i] * 4));
    if (p) p->conv[i] = {w, sc, sh, kConvCin[i], kConvCout[i]};
  }
  const int h4 = kH / 4;
  void* a = take((size_t)h4 * h4 * 2 * kNF * 2 * kNF * 2);
  void* b = take((size_t)4 * kNF * 4 * kNF * 2);
  void* c = take((size_t)4 * kNF * 2 * kNF * 2);
  void* d = take((size_t)9 * kNF * 4);
  if (p) {
    p->up0_w = reinterpret_cast<__nv_bfloat16*>(a);
    p->up1_w = reinterpret_cast<__nv_bfloat16*>(b);
    p->up2_w = reinterpret_cast<__nv_bfloat16*>(c);
    p->out3_w = reinterpret_cast<float*>(d);
  }
  return off;
}

size_t ws_layout(cdm_plan* p, uint8_t* base, int batch, int reps) {
  const size_t n = (size_t)batch * reps, h = kH, h2 = kH / 2, h4 = kH / 4, nf = kNF;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = base ? base + off : nullptr;
    off += al(bytes);
    return r;
  };
  void* x0 = take(n * h * h * nf * 2);
  void* p64 = take(n * h * h * nf * 2);
  void* q64 = take(n * h * h * nf * 2);
  void* d1 = take(n * h2 * h2 * nf * 2);
  void* p32w = take(n * h2 * h2 * 2 * nf * 2);
  void* q32w = take(n * h2 * h2 * 2 * nf * 2);
  void* u1f = take(n * h2 * h2 * nf * 2);
  void* d2 = take(n * h4 * h4 * 2 * nf * 2);
  void* hidden = take(n * 2 * nf * 2);
  void* u0raw = take(n * h4 * h4 * 2 * nf * 2);
  void* u0f = take(n * h4 * h4 * 2 * nf * 2);
  void* gnp = take(n * (h / 16) * (h / 16) * 8 * 16 * 4);
  void* gmr = take(n * 16 * 4);
  void* eps = take(n * h * h * 4);
  if (p) {
    p->x0 = (__nv_bfloat16*)x0, p->p64 = (__nv_bfloat16*)p64, p->q64 = (__nv_bfloat16*)q64, p->d1 = (__nv_bfloat16*)d1;
    p->p32w = (__nv_bfloat16*)p32w, p->q32w = (__nv_bfloat16*)q32w, p->u1f = (__nv_bfloat16*)u1f;
    p->d2 = (__nv_bfloat16*)d2, p->hidden = (__nv_bfloat16*)hidden, p->u0raw = (__nv_bfloat16*)u0raw;
    p->u0f = (__nv_bfloat16*)u0f, p->gn_partial = (float*)gnp, p->gn_mr = (float*)gmr, p->eps = (float*)eps;
  }
  return off;
}

const float* T(const cdm_plan* p, const std::string& name) { return p->tensors[tensor_index(name)]; }

int pack_perm(const float* src, void* dst, int d0, int d1, int d2, int d3, long long s0, long long s1, long long s2,
              long long s3, cudaStream_t st) {
  const long long total = (long long)d0 * d1 * d2 * d3;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
  plan_pack_bf16_kernel<<<(int)blocks, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), d1, d2, d3, s0, s1, s2, s3,
                                                     total);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

int pack_weights(cdm_plan* p, cudaStream_t st) {
  int rc;
  for (int i = 0; i < 18; ++i) {
    const std::string pre = kConvPrefix[i];
    const ConvSlot& s = p->conv[i];
    const float* w = T(p, pre + ".0.weight");
    if (i == 0) {
      plan_transpose9_kernel<<<(9 * kNF + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<float*>(s.w), kNF);
      CDM_CHECK_LAUNCH();
    } else {  // OIHW -> [co][kh][kw][ci]
      rc = pack_perm(w, s.w, s.cout, 3, 3, s.cin, (long long)s.cin * 9, 3, 1, 9, st);
      if (rc) return rc;
    }
    plan_fold_bn_kernel<<<(s.cout + 127) / 128, 128, 0, st>>>(T(p, pre + ".0.bias"), T(p, pre + ".1.weight"),
                                                              T(p, pre + ".1.bias"), T(p, pre + ".1.running_mean"),
                                                              T(p, pre + ".1.running_var"), 1e-5f, s.cout, s.scale, s.shift);
    CDM_CHECK_LAUNCH();
  }
  {  // out.0: plain convolution, bias as the shift
    const ConvSlot& s = p->conv[18];
    rc = pack_perm(T(p, "out.0.weight"), s.w, s.cout, 3, 3, s.cin, (long long)s.cin * 9, 3, 1, 9, st);
    if (rc) return rc;
    plan_fill_kernel<<<1, 128, 0, st>>>(s.scale, 1.f, s.cout);
    CDM_CHECK_LAUNCH();
    CDM_CHECK_CUDA(cudaMemcpyAsync(s.shift, T(p, "out.0.bias"), s.cout * 4, cudaMemcpyDeviceToDevice, st));
  }
  // ConvTranspose2d IOHW [ci][co][kh][kw] -> GEMM B operand [(kh,kw,co)][ci]
  const int h4 = kH / 4;
  rc = pack_perm(T(p, "up0.0.weight"), p->up0_w, h4, h4, 2 * kNF, 2 * kNF, h4, 1, (long long)h4 * h4,
                 (long long)2 * kNF * h4 * h4, st);
  if (rc) return rc;
  rc = pack_perm(T(p, "up1.model.0.weight"), p->up1_w, 2, 2, kNF, 4 * kNF, 2, 1, 4, (long long)kNF * 4, st);
  if (rc) return rc;
  rc = pack_perm(T(p, "up2.model.0.weight"), p->up2_w, 2, 2, kNF, 2 * kNF, 2, 1, 4, (long long)kNF * 4, st);
  if (rc) return rc;
  plan_transpose9_kernel<<<(9 * kNF + 255) / 256, 256, 0, st>>>(T(p, "out.3.weight"), p->out3_w, kNF);
  CDM_CHECK_LAUNCH();
  return CDM_OK;
}

int prepare_launches(cdm_plan* p) {
  const int n = p->n, B = p->batch, h = kH, h2 = kH / 2, h4 = kH / 4;
  int rc;
  auto conv = [&](int i, const void* src, int H, void* out, int flags, const void* src1 = nullptr, int c1 = 0,
                  int n_img = -1) {
    cdm_conv3x3_args a;
    memset(&a, 0, sizeof(a));
    const ConvSlot& s = p->conv[i];
    a.src0 = src, a.src1 = src1, a.c0 = s.cin - c1, a.c1 = c1;
    a.n_img = n_img < 0 ? n : n_img, a.H = H, a.W = H;
    a.weight = s.w, a.cout = s.cout, a.scale = s.scale, a.shift = s.shift, a.flags = flags, a.out = out;
    a.mode = p->conv_mode;
    a.sc_reps = 1, a.film_shift_rows = 1;
    if (flags & CDM_EPI_SHORTCUT) {  // per-call pointers are patched in cdm_forward_eval; placeholders pass the checks
      a.sc_x = reinterpret_cast<const float*>(p->eps), a.sc_tab = reinterpret_cast<const float*>(p->eps), a.sc_reps = p->reps;
    }
    if (flags & CDM_EPI_FILM) a.film_scale = a.film_shift = reinterpret_cast<const float*>(p->eps);
    if (flags & CDM_EPI_GNSTATS) a.gn_partial = p->gn_partial;
    return conv_prepare(&a, &p->cl[i]);
  };
  const int R = CDM_EPI_RELU;
  // init_conv.conv1 (cl[0] unused: K = 9 runs on cdm_conv_in); its output aliases q64 (dead before q64 is written)
  if ((rc = conv(1, p->q64, h, p->x0, R | CDM_EPI_SHORTCUT, nullptr, 0, B))) return rc;
  if ((rc = conv(2, p->x0, h, p->p64, R))) return rc;
  if ((rc = conv(3, p->p64, h, p->q64, R))) return rc;
  if ((rc = conv(4, p->q64, h, p->p64, R))) return rc;
  if ((rc = conv(5, p->p64, h, p->d1, R | CDM_EPI_POOL))) return rc;
  if ((rc = conv(6, p->d1, h2, p->p32w, R))) return rc;
  if ((rc = conv(7, p->p32w, h2, p->q32w, R))) return rc;
  if ((rc = conv(8, p->q32w, h2, p->p32w, R))) return rc;
  if ((rc = conv(9, p->p32w, h2, p->d2, R | CDM_EPI_POOL))) return rc;
  // the nf-wide h/2 buffers of up1 alias the (dead by then) 2nf-wide ones of down2
  __nv_bfloat16 *p32 = p->p32w, *q32 = p->q32w;
  if ((rc = conv(10, p32, h2, q32, R))) return rc;
  if ((rc = conv(11, q32, h2, p32, R))) return rc;
  if ((rc = conv(12, p32, h2, q32, R))) return rc;
  if ((rc = conv(13, q32, h2, p->u1f, R | CDM_EPI_FILM))) return rc;
  if ((rc = conv(14, p->p64, h, p->q64, R))) return rc;
  if ((rc = conv(15, p->q64, h, p->p64, R))) return rc;
  if ((rc = conv(16, p->p64, h, p->q64, R))) return rc;
  if ((rc = conv(17, p->q64, h, p->p64, R))) return rc;
  if ((rc = conv(18, p->p64, h, p->q64, CDM_EPI_GNSTATS, p->x0, kNF))) return rc;

  cdm_gemm_args g;
  memset(&g, 0, sizeof(g));  // up0: [n,256] x [256,65536]
  g.a0 = p->hidden, g.k0 = 2 * kNF, g.M = n, g.N = h4 * h4 * 2 * kNF, g.bw = p->up0_w;
  g.shift = p->up0_b, g.shift_mod = 2 * kNF, g.out = p->u0raw;
  if ((rc = gemm_prepare(&g, &p->gl[0]))) return rc;
  memset(&g, 0, sizeof(g));  // up1: cat(film(up0), d2) -> ConvTranspose2d(512,128,2,2)
  g.a0 = p->u0f, g.k0 = 2 * kNF, g.a1 = p->d2, g.k1 = 2 * kNF, g.M = n * h4 * h4, g.N = 4 * kNF, g.bw = p->up1_w;
  g.shift = p->up1_b, g.shift_mod = kNF, g.out_mode = 1, g.H = h4, g.W = h4, g.out = p32;
  if ((rc = gemm_prepare(&g, &p->gl[1]))) return rc;
  memset(&g, 0, sizeof(g));  // up2: cat(film(up1), d1) -> ConvTranspose2d(256,128,2,2)
  g.a0 = p->u1f, g.k0 = kNF, g.a1 = p->d1, g.k1 = kNF, g.M = n * h2 * h2, g.N = 4 * kNF, g.bw = p->up2_w;
  g.shift = p->up2_b, g.shift_mod = kNF, g.out_mode = 1, g.H = h2, g.W = h2, g.out = p->p64;
  if ((rc = gemm_prepare(&g, &p->gl[2]))) return rc;
  return CDM_OK;
}

}  // namespace

extern "C" int cdm_plan_n_tensors(void) { return (int)tensor_table().size(); }
extern "C" const char* cdm_plan_tensor_name(int i) {
  const auto& t = tensor_table();
  return (i >= 0 && i < (int)t.size()) ? t[i].name.c_str() : nullptr;
}
extern "C" long long cdm_plan_tensor_numel(int i, int n_cfeat) {
  const auto& t = tensor_table();
  if (i < 0 || i >= (int)t.size() || n_cfeat < 1) return -1;
  return t[i].numel + t[i].cf_scale * (n_cfeat - 1);
}
extern "C" long long cdm_plan_arena_bytes(int n_cfeat) {
  (void)n_cfeat;  // the context width only changes fp32 EmbedFC weights, which are used in place
  return (long long)arena_layout(nullptr, nullptr);
}
extern "C" long long cdm_plan_workspace_bytes(int batch, int reps) {
  if (batch < 1 || reps < 1) return -1;
  return (long long)ws_layout(nullptr, nullptr, batch, reps);
}

extern "C" int cdm_plan_create(const cdm_plan_desc* d, void* stream, cdm_plan** out) {
  CDM_CHECK_ARG(d != nullptr && out != nullptr);
  *out = nullptr;
  CDM_CHECK_ARG(d->n_cfeat >= 1 && d->batch >= 1 && d->reps >= 1 && d->reps <= 2);
  CDM_CHECK_ARG(d->tensors && d->arena && d->workspace);
  CDM_CHECK_ARG(d->arena_bytes >= cdm_plan_arena_bytes(d->n_cfeat));
  CDM_CHECK_ARG(d->workspace_bytes >= cdm_plan_workspace_bytes(d->batch, d->reps));
  CDM_CHECK_ARG(((uintptr_t)d->arena & 255) == 0 && ((uintptr_t)d->workspace & 255) == 0);
  CDM_CHECK_ARG(d->conv_mode >= 0 && d->conv_mode <= 4);
  for (int i = 0; i < cdm_plan_n_tensors(); ++i) {
    if (!d->tensors[i]) {
      set_error("cdm_plan_create: tensor %d (%s) is NULL", i, cdm_plan_tensor_name(i));
      return CDM_ERR_ARG;
    }
  }
  int rc = check_device();
  if (rc) return rc;
  cdm_plan* p = new cdm_plan();
  p->n_cfeat = d->n_cfeat, p->batch = d->batch, p->reps = d->reps, p->n = d->batch * d->reps;
  p->conv_mode = d->conv_mode == 0 ? CDM_CONV_MODE_DEFAULT : d->conv_mode;
  p->tensors.assign(d->tensors, d->tensors + cdm_plan_n_tensors());
  p->arena = reinterpret_cast<uint8_t*>(d->arena);
  p->ws = reinterpret_cast<uint8_t*>(d->workspace);
  arena_layout(p, p->arena);
  ws_layout(p, p->ws, p->batch, p->reps);
  p->up0_b = T(p, "up0.0.bias"), p->up0_g = T(p, "up0.1.weight"), p->up0_beta = T(p, "up0.1.bias");
  p->up1_b = T(p, "up1.model.0.bias"), p->up2_b = T(p, "up2.model.0.bias");
  p->out_g = T(p, "out.1.weight"), p->out_beta = T(p, "out.1.bias"), p->out3_b = T(p, "out.3.bias");
  {
    static const char* const pre[4] = {"contextembed1", "timeembed1", "contextembed2", "timeembed2"};
    static const char* const suf[4] = {".model.0.weight", ".model.0.bias", ".model.2.weight", ".model.2.bias"};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) p->emb[i][j] = T(p, std::string(pre[i]) + suf[j]);
  }
  rc = pack_weights(p, reinterpret_cast<cudaStream_t>(stream));
  if (!rc) rc = prepare_launches(p);
  if (rc) {
    delete p;
    return rc;
  }
  *out = p;
  return CDM_OK;
}

extern "C" int cdm_plan_refresh(cdm_plan* p, void* stream) {
  CDM_CHECK_ARG(p != nullptr);
  return pack_weights(p, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" void cdm_plan_destroy(cdm_plan* p) { delete p; }

extern "C" int cdm_plan_buffer(const cdm_plan* p, const char* name, void** ptr, long long* bytes) {
  CDM_CHECK_ARG(p && name && ptr && bytes);
  const size_t n = p->n, h = kH, h2 = kH / 2, h4 = kH / 4, nf = kNF;
  struct {
    const char* nm;
    void* ptr;
    size_t bytes;
  } tab[] = {{"x0", p->x0, n * h * h * nf * 2},          {"p64", p->p64, n * h * h * nf * 2},
             {"q64", p->q64, n * h * h * nf * 2},        {"d1", p->d1, n * h2 * h2 * nf * 2},
             {"p32w", p->p32w, n * h2 * h2 * 2 * nf * 2}, {"q32w", p->q32w, n * h2 * h2 * 2 * nf * 2},
             {"u1f", p->u1f, n * h2 * h2 * nf * 2},      {"d2", p->d2, n * h4 * h4 * 2 * nf * 2},
             {"hidden", p->hidden, n * 2 * nf * 2},      {"u0raw", p->u0raw, n * h4 * h4 * 2 * nf * 2},
             {"u0f", p->u0f, n * h4 * h4 * 2 * nf * 2},  {"gn_mr", p->gn_mr, n * 16 * 4},
             {"gn_partial", p->gn_partial, n * (h / 16) * (h / 16) * 8 * 16 * 4},
             {"eps", p->eps, n * h * h * 4}};
  for (auto& e : tab)
    if (!strcmp(e.nm, name)) {
      *ptr = e.ptr;
      *bytes = (long long)e.bytes;
      return CDM_OK;
    }
  set_error("cdm_plan_buffer: unknown buffer '%s'", name);
  return CDM_ERR_ARG;
}

extern "C" int cdm_plan_embed(const cdm_plan* p, int which, const float* in, int rows, float* out, void* stream) {
  CDM_CHECK_ARG(p && in && out && rows >= 1 && which >= 0 && which < 4);
  const int din = (which & 1) ? 1 : p->n_cfeat, emb = which < 2 ? 2 * kNF : kNF;
  return cdm_embed_fc(in, rows, din, p->emb[which][0], p->emb[which][1], p->emb[which][2], p->emb[which][3], emb, out,
                      stream);
}

// The 26 launches of one eval forward, in order (cdm_plan_launch_name / cdm_plan_profile index).
static const char* const kLaunchNames[26] = {
    "conv_in init_conv.conv1",  "conv3x3 init_conv.conv2+shortcut", "conv3x3 down1.0.conv1", "conv3x3 down1.0.conv2",
    "conv3x3 down1.1.conv1",    "conv3x3 down1.1.conv2+pool",       "conv3x3 down2.0.conv1", "conv3x3 down2.0.conv2",
    "conv3x3 down2.1.conv1",    "conv3x3 down2.1.conv2+pool",       "avgpool_gelu to_vec",   "gemm up0",
    "gn_relu_film up0",         "gemm up1.convT",                   "conv3x3 up1.1.conv1",   "conv3x3 up1.1.conv2",
    "conv3x3 up1.2.conv1",      "conv3x3 up1.2.conv2+film",         "gemm up2.convT",        "conv3x3 up2.1.conv1",
    "conv3x3 up2.1.conv2",      "conv3x3 up2.2.conv1",              "conv3x3 up2.2.conv2",   "conv3x3 out.0+gnstats",
    "gn_finalize out.1",        "conv_out out.1-3"};

// ev != nullptr: record an event after every launch (ev[0] before the first), for cdm_plan_profile
static int forward_impl(cdm_plan* p, const cdm_forward_args* f, cudaStream_t st, cudaEvent_t* ev) {
  CDM_CHECK_ARG(p != nullptr && f != nullptr);
  CDM_CHECK_ARG(f->x && f->sc_tab && f->cemb1 && f->temb1 && f->cemb2 && f->temb2);
  CDM_CHECK_ARG(f->temb_rows == 1 || f->temb_rows == p->n);
  void* stream = st;
  const int n = p->n, B = p->batch, h = kH, h4 = kH / 4;
  int rc, k = 0;
  if (ev) CDM_CHECK_CUDA(cudaEventRecord(ev[0], st));
#define CDM_STEP(call)                                         \
  do {                                                         \
    if ((rc = (call))) return rc;                              \
    if (ev) CDM_CHECK_CUDA(cudaEventRecord(ev[++k], st));      \
  } while (0)
  {  // init_conv.conv1: Conv2d(1,128) + BN + ReLU on the B shared inputs
    cdm_conv_in_args a;
    memset(&a, 0, sizeof(a));
    const ConvSlot& s = p->conv[0];
    a.x = f->x, a.n_img = B, a.H = h, a.W = h, a.weight = reinterpret_cast<const float*>(s.w), a.cout = kNF;
    a.scale = s.scale, a.shift = s.shift, a.relu = 1, a.out = p->q64;
    CDM_STEP(cdm_conv_in(&a, stream));
  }
  {  // init_conv.conv2 + the fresh 1x1 shortcut, fanned out to the `reps` passes
    ConvLaunch L = p->cl[1];
    L.p.sc_x = f->x, L.p.sc_tab = f->sc_tab, L.p.step_ptr = f->step_ptr;
    CDM_STEP(conv_launch(L, st));
  }
  for (int i = 2; i <= 9; ++i) CDM_STEP(conv_launch(p->cl[i], st));
  CDM_STEP(cdm_avgpool_gelu(p->d2, n, h4 * h4, 2 * kNF, p->hidden, stream));
  CDM_STEP(gemm_launch(p->gl[0], st));
  {
    cdm_gn_relu_film_args a;
    memset(&a, 0, sizeof(a));
    a.src = p->u0raw, a.n_img = n, a.P = h4 * h4, a.C = 2 * kNF, a.groups = 8;
    a.gamma = p->up0_g, a.beta = p->up0_beta, a.eps = 1e-5f;
    a.film_scale = f->cemb1, a.film_shift = f->temb1, a.film_rows = f->temb_rows, a.step_ptr = f->step_ptr;
    a.out = p->u0f;
    CDM_STEP(cdm_gn_relu_film(&a, stream));
  }
  CDM_STEP(gemm_launch(p->gl[1], st));
  for (int i = 10; i <= 12; ++i) CDM_STEP(conv_launch(p->cl[i], st));
  {
    ConvLaunch L = p->cl[13];
    L.p.film_scale = f->cemb2, L.p.film_shift = f->temb2, L.p.film_shift_rows = f->temb_rows, L.p.step_ptr = f->step_ptr;
    CDM_STEP(conv_launch(L, st));
  }
  CDM_STEP(gemm_launch(p->gl[2], st));
  for (int i = 14; i <= 18; ++i) CDM_STEP(conv_launch(p->cl[i], st));
  CDM_STEP(cdm_gn_finalize(p->gn_partial, n, (h / 16) * (h / 16) * 8, (float)((kNF / 8) * h * h), 1e-5f, p->gn_mr, stream));
  {
    cdm_conv_out_args a;
    memset(&a, 0, sizeof(a));
    a.src = p->q64, a.n_img = n, a.H = h, a.W = h, a.C = kNF, a.mean_rstd = p->gn_mr;
    a.gamma = p->out_g, a.beta = p->out_beta, a.weight = p->out3_w, a.bias = p->out3_b;
    a.out = f->eps ? f->eps : p->eps;
    CDM_STEP(cdm_conv_out(&a, stream));
  }
#undef CDM_STEP
  return CDM_OK;
}

extern "C" int cdm_forward_eval(cdm_plan* p, const cdm_forward_args* f, void* stream) {
  return forward_impl(p, f, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

extern "C" int cdm_plan_n_launches(void) { return 26; }
extern "C" const char* cdm_plan_launch_name(int i) { return (i >= 0 && i < 26) ? kLaunchNames[i] : nullptr; }

// One forward with a CUDA event after every launch (recorded on `stream`, the stream the kernels run on);
// synchronises, then ms_host[i] = duration of launch i.  Measurement only (bench.py's roofline figures).
extern "C" int cdm_plan_profile(cdm_plan* p, const cdm_forward_args* f, float* ms_host, void* stream) {
  CDM_CHECK_ARG(ms_host != nullptr);
  cudaEvent_t ev[27];
  for (int i = 0; i < 27; ++i) CDM_CHECK_CUDA(cudaEventCreate(&ev[i]));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = forward_impl(p, f, st, ev);
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("cdm_plan_profile: %s", cudaGetErrorString(e));
      rc = CDM_ERR_CUDA;
    }
  }
  if (!rc)
    for (int i = 0; i < 26; ++i) cudaEventElapsedTime(&ms_host[i], ev[i], ev[i + 1]);
  for (int i = 0; i < 27; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

extern "C" int cdm_sample_step(cdm_plan* p, const cdm_sample_step_args* s, void* stream) {
  CDM_CHECK_ARG(p != nullptr && s != nullptr && s->x && s->coef && s->step_ptr && s->timesteps >= 1);
  cdm_forward_args f = s->fwd;
  f.x = s->x;
  f.step_ptr = s->step_ptr;
  f.eps = nullptr;  // the plan's own eps buffer
  int rc = cdm_forward_eval(p, &f, stream);
  if (rc) return rc;
  cdm_ddpm_step_args a;
  memset(&a, 0, sizeof(a));
  a.x = s->x, a.eps = p->eps, a.n = p->batch, a.hw = kH * kH, a.reps = p->reps, a.guide_w = s->guide_w;
  a.coef = s->coef, a.step_ptr = s->step_ptr, a.timesteps = s->timesteps;
  a.z = s->z, a.z_iter_stride = s->z_iter_stride, a.seed = s->seed, a.sample_offset = s->sample_offset;
  a.snap = s->snap, a.snap_slot = s->snap_slot;
  if ((rc = cdm_ddpm_step(&a, stream))) return rc;
  return cdm_step_advance(s->step_ptr, -1, stream);
}
