// api.cu -- the C ABI of libsga_b200.so (include/sga_b200.h): context, weight binding, and the per-iteration
// orchestration of the SGA graph (reference sga.py:88-174, 207-215, 240-247) on one CUDA stream.
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "kernels.h"
#include "tc_kernels.h"

namespace sga {

static thread_local char g_err[512] = "";
thread_local long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef __nv_bfloat16 bf16;

struct ConvW {
  int kh = 0, kw = 0, cin = 0, cout = 0;
  float* w = nullptr;       // [tap][cin][cout] (reference layout)
  float* w_t = nullptr;     // [tap][cout][cin] (SIMT dgrad)
  bf16* wb_fwd = nullptr;   // [tap][cout][cin] K-major B operand
  bf16* wb_dgrad = nullptr; // [tap][cin][cout]
  float* bias = nullptr;
};

// bits-back variant (bb_sga.py): how z_tilde is produced and which part of the graph is evaluated
struct BbArgs {
  const float* z_mean;
  const float* z_logvar;
  const float* eps;     // injected N(0,1) noise or nullptr -> Philox + Box-Muller
  int rate_only;        // second loop (bb_sga.py:236-256): y_tilde fixed, loss = train_bpp, no synthesis transform
  int y_hard;           // y holds y_tilde (already rounded) instead of y
};

struct Dims {
  int N, H, W, C, mid;
  int Hg[5], Wg[5];   // g_a pyramid: Hg[0]=H ... Hg[4]=Hy
  int Hy, Wy, Hz1, Wz1, Hz, Wz;
  int Hs[5], Ws[5];   // g_s pyramid: Hs[0]=Hy ... Hs[4]=16Hy
  int Hh, Wh;         // h_s output grid (4*Hz)
  long long ny, nz;
};

}  // namespace sga

using namespace sga;

struct sga_ctx {
  sga_config cfg;
  Dims d;
  bool bound = false;
  std::vector<void*> allocs;
  size_t bytes = 0;

  ConvW ga[4], gs[4], ha[3], hs[3];
  float *ga_beta[3], *ga_gamma[3], *gs_beta[3], *gs_gamma[3], *gs_gamma_t[3];
  bf16 *gs_gamma_b[3], *gs_gamma_tb[3];
  float* eb_pack = nullptr;
  float gs3_bias_host[3];
  bf16 *l3_wscatter = nullptr, *l3_wgather = nullptr;

  // shared
  double* acc = nullptr;
  float *dz, *gz_rate, *gz_hyper, *dy, *gy_rate, *gy_dist, *h2;
  // fp32 path
  float *f_zt, *f_h0, *f_h1, *f_yt, *f_gh2, *f_xpre[3], *f_s[3], *f_act[3], *f_xt, *f_gx, *f_tmp, *f_nrm;
  float *f_ga[3], *f_gb[3], *f_gh1, *f_gh0;   // backward (only when cfg.operand == F32)
  float *f_init[2];                            // analysis scratch
  float* bb_ha = nullptr;                      // bits-back variant: h_a output [n,hz,wz,2C] before the mean/logvar split
  // bf16 path
  bf16 *b_zt, *b_h0, *b_h1, *b_yt, *b_gh2, *b_xpre[3], *b_s[3], *b_act[3], *b_gx4, *b_ga[3], *b_gb[3], *b_gh1, *b_gh0;
  TcConv *tc_hs[3] = {0, 0, 0}, *tc_gs[3] = {0, 0, 0}, *tc_hs_d[3] = {0, 0, 0}, *tc_gs_d[3] = {0, 0, 0};
  IgdnTc *ig_f[3] = {0, 0, 0}, *ig_b[3] = {0, 0, 0};
  L3Tc* l3 = nullptr;
  bool fused = false;   // IGDN runs inside the (de)conv epilogues (num_filters in {64,128,192})

  // event-based per-kernel timing (sga_profile_*)
  bool prof = false;
  struct ProfRec { int label; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }

  template <typename T>
  int alloc(T** p, long long count) {
    if (count <= 0) count = 1;
    void* q = nullptr;
    const size_t nbytes = static_cast<size_t>(count) * sizeof(T);
    cudaError_t e = cudaMalloc(&q, nbytes);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", nbytes, cudaGetErrorString(e));
      return SGA_ERR_NOMEM;
    }
    allocs.push_back(q);
    bytes += nbytes;
    *p = static_cast<T*>(q);
    return SGA_OK;
  }
};

namespace sga {

// run `expr` (one launch group) bracketed by events when profiling is on
#define PROF(c, label, st, expr)                                           \
  do {                                                                     \
    if ((c)->prof && (c)->prof_recs.size() < 16384) {                      \
      sga_ctx::ProfRec _r{(label), (c)->prof_event(), (c)->prof_event()};  \
      cudaEventRecord(_r.a, (st));                                         \
      int _s = (expr);                                                     \
      cudaEventRecord(_r.b, (st));                                         \
      (c)->prof_recs.push_back(_r);                                        \
      if (_s != SGA_OK) return _s;                                         \
    } else {                                                               \
      SGA_TRY(expr);                                                       \
    }                                                                      \
  } while (0)

static void compute_dims(const sga_config& c, Dims& d) {
  d.N = c.n_images; d.H = c.height; d.W = c.width; d.C = c.num_filters;
  d.mid = static_cast<int>(c.num_filters * 1.5);   // nn_models.py:157
  d.Hg[0] = d.H; d.Wg[0] = d.W;
  for (int i = 1; i <= 4; ++i) { d.Hg[i] = (d.Hg[i - 1] + 1) / 2; d.Wg[i] = (d.Wg[i - 1] + 1) / 2; }
  d.Hy = d.Hg[4]; d.Wy = d.Wg[4];
  d.Hz1 = (d.Hy + 1) / 2; d.Wz1 = (d.Wy + 1) / 2;
  d.Hz = (d.Hz1 + 1) / 2; d.Wz = (d.Wz1 + 1) / 2;
  d.Hs[0] = d.Hy; d.Ws[0] = d.Wy;
  for (int i = 1; i <= 4; ++i) { d.Hs[i] = 2 * d.Hs[i - 1]; d.Ws[i] = 2 * d.Ws[i - 1]; }
  d.Hh = 4 * d.Hz; d.Wh = 4 * d.Wz;
  d.ny = static_cast<long long>(d.N) * d.Hy * d.Wy * d.C;
  d.nz = static_cast<long long>(d.N) * d.Hz * d.Wz * d.C;
}

static long long gs_elems(const Dims& d, int level) {  // level 1..3: output of g_s layer level-1
  return static_cast<long long>(d.N) * d.Hs[level] * d.Ws[level] * d.C;
}

static int conv_f32(int kind, const ConvW& w, int N, int Hin, int Win, const float* x, bool use_bias, int relu,
                    const float* mask, float* out, cudaStream_t st) {
  ConvPlan plan;
  SGA_TRY(make_conv_plan(kind, w.kh, w.kw, N, Hin, Win, w.cin, w.cout, &plan));
  const float* wp = plan.w_transposed ? w.w_t : w.w;
  return launch_conv_simt(plan, x, wp, use_bias ? w.bias : nullptr, mask, relu, out, st);
}

// 1x1 channel mixing out = x @ M (+ bias): M is [K][Nn]
static int mix_f32(const float* x, const float* M, const float* bias, long long pixels, int C, float* out,
                   cudaStream_t st) {
  ConvPlan plan;
  SGA_CHECK(pixels < (1LL << 31), "too many pixels");
  SGA_TRY(make_conv_plan(kS1Corr, 1, 1, 1, 1, static_cast<int>(pixels), C, C, &plan));
  return launch_conv_simt(plan, x, M, bias, nullptr, 0, out, st);
}

static int gdn_f32(sga_ctx* c, int inverse, const float* x, const float* beta, const float* gamma, long long pixels,
                   int C, float* tmp, float* nrm, float* out, float* s_out, cudaStream_t st) {
  const long long n = pixels * C;
  SGA_TRY(launch_square(x, n, tmp, st));
  SGA_TRY(mix_f32(tmp, gamma, beta, pixels, C, nrm, st));
  return launch_gdn_finish(x, nrm, n, inverse, out, s_out, st);
}

static int igdn_bwd_f32(const float* g, const float* x, const float* s, const float* gamma_t, long long pixels, int C,
                        float* tmp, float* u, float* gx, cudaStream_t st) {
  const long long n = pixels * C;
  SGA_TRY(launch_igdn_bwd_pre(g, x, s, n, tmp, st));
  SGA_TRY(mix_f32(tmp, gamma_t, nullptr, pixels, C, u, st));
  return launch_igdn_bwd_finish(g, x, s, u, n, gx, st);
}

static int bind_conv(sga_ctx* c, ConvW& w, const float* kernel, const float* bias, int kh, int kw, int cin, int cout,
                     bool need_bf16, cudaStream_t st) {
  w.kh = kh; w.kw = kw; w.cin = cin; w.cout = cout;
  const long long n = static_cast<long long>(kh) * kw * cin * cout;
  SGA_TRY(c->alloc(&w.w, n));
  SGA_TRY(c->alloc(&w.w_t, n));
  SGA_CUDA(cudaMemcpyAsync(w.w, kernel, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SGA_TRY(launch_transpose_taps(w.w, kh * kw, cin, cout, w.w_t, st));
  if (bias) {
    SGA_TRY(c->alloc(&w.bias, cout));
    SGA_CUDA(cudaMemcpyAsync(w.bias, bias, cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  if (need_bf16) {
    SGA_TRY(c->alloc(&w.wb_fwd, n));
    SGA_TRY(c->alloc(&w.wb_dgrad, n));
    SGA_TRY(launch_pack_weights_bf16(w.w, kh * kw, cin, cout, false, w.wb_fwd, st));
    SGA_TRY(launch_pack_weights_bf16(w.w, kh * kw, cin, cout, true, w.wb_dgrad, st));
  }
  return SGA_OK;
}

static int make_tc(int kind, const ConvW& w, int N, int Hin, int Win, const bf16* x, TcConv** out, int fuse = 0,
                   const bf16* mix_op = nullptr) {
  ConvPlan plan;
  SGA_TRY(make_conv_plan(kind, w.kh, w.kw, N, Hin, Win, w.cin, w.cout, &plan));
  return tc_conv_create(plan, x, plan.w_transposed ? w.wb_dgrad : w.wb_fwd, fuse, mix_op, out);
}

// ---------------------------------------------------------------------------------------------------
// forward / backward, fp32 path
// ---------------------------------------------------------------------------------------------------
static int forward_f32(sga_ctx* c, const float* x, const float* y, const float* z, float T, int step, const float* ny,
                       const float* nz, int hard, bool want_grad, float* stats, cudaStream_t st,
                       const BbArgs* bb = nullptr) {
  const Dims& d = c->d;
  const sga_config& g = c->cfg;
  const float rate_scale = static_cast<float>(1.0 / (log(2.0) * d.H * static_cast<double>(d.W) * g.n_ref));
  SGA_CUDA(cudaMemsetAsync(c->acc, 0, 5 * d.N * sizeof(double), st));
  if (bb) {
    hard = bb->y_hard;
    SGA_TRY(launch_zbb_path<float>(bb->z_mean, bb->z_logvar, bb->eps, d.nz, d.nz / d.N, d.C, c->eb_pack, g.seed, step,
                                   rate_scale, c->f_zt, want_grad ? c->gz_rate : nullptr, want_grad ? c->dz : nullptr,
                                   c->acc + d.N, c->acc + 4 * d.N, g.first_image * (d.nz / d.N), st));
  } else {
    SGA_TRY(launch_z_path<float>(z, nz, d.nz, d.nz / d.N, d.C, c->eb_pack, T, g.clip_eps, g.seed, step, hard, rate_scale,
                                 c->f_zt, want_grad ? c->gz_rate : nullptr, want_grad ? c->dz : nullptr, c->acc + d.N,
                                 g.first_image * (d.nz / d.N), st));
  }
  SGA_TRY(conv_f32(kUp2, c->hs[0], d.N, d.Hz, d.Wz, c->f_zt, true, 1, nullptr, c->f_h0, st));
  SGA_TRY(conv_f32(kUp2, c->hs[1], d.N, 2 * d.Hz, 2 * d.Wz, c->f_h0, true, 1, nullptr, c->f_h1, st));
  SGA_TRY(conv_f32(kS1True, c->hs[2], d.N, d.Hh, d.Wh, c->f_h1, true, 0, nullptr, c->h2, st));
  SGA_TRY(launch_y_path<float>(y, ny, c->h2, d.N, d.Hy, d.Wy, d.Hh, d.Wh, d.C, T, g.clip_eps, g.seed, step, hard,
                               rate_scale, c->f_yt, want_grad ? c->gy_rate : nullptr, want_grad ? c->dy : nullptr,
                               want_grad ? c->f_gh2 : nullptr, c->acc, g.first_image * (d.ny / d.N), st));
  if (bb && bb->rate_only) {
    if (stats) SGA_TRY(launch_finalize_stats_bb(c->acc, d.N, g.n_ref, d.H, d.W, g.lmbda, 1, stats, st));
    return SGA_OK;
  }
  const float* in = c->f_yt;
  for (int i = 0; i < 3; ++i) {
    SGA_TRY(conv_f32(kUp2, c->gs[i], d.N, d.Hs[i], d.Ws[i], in, true, 0, nullptr, c->f_xpre[i], st));
    const long long pix = static_cast<long long>(d.N) * d.Hs[i + 1] * d.Ws[i + 1];
    SGA_TRY(gdn_f32(c, 1, c->f_xpre[i], c->gs_beta[i], c->gs_gamma[i], pix, d.C, c->f_tmp, c->f_nrm, c->f_act[i],
                    c->f_s[i], st));
    in = c->f_act[i];
  }
  SGA_TRY(conv_f32(kUp2, c->gs[3], d.N, d.Hs[3], d.Ws[3], in, true, 0, nullptr, c->f_xt, st));
  const float grad_scale =
      (g.lmbda > 0.0f) ? static_cast<float>(2.0 * 65025.0 * g.lmbda / (static_cast<double>(g.n_ref) * d.H * d.W * 3.0))
                       : 0.0f;
  SGA_TRY(launch_distortion(x, c->f_xt, d.N, d.H, d.W, d.Hs[4], d.Ws[4], grad_scale, want_grad ? c->f_gx : nullptr,
                            c->acc + 2 * d.N, c->acc + 3 * d.N, st));
  if (stats) {
    if (bb) SGA_TRY(launch_finalize_stats_bb(c->acc, d.N, g.n_ref, d.H, d.W, g.lmbda, 0, stats, st));
    else SGA_TRY(launch_finalize_stats(c->acc, d.N, g.n_ref, d.H, d.W, g.lmbda, stats, st));
  }
  return SGA_OK;
}

static int backward_f32(sga_ctx* c, cudaStream_t st, bool rate_only = false) {
  const Dims& d = c->d;
  SGA_CHECK(c->f_ga[0] != nullptr, "context was created with operand=bf16: fp32 backward buffers are not allocated");
  if (rate_only) {
    // y_tilde is fixed: only the hyper path carries gradient
  } else if (c->cfg.lmbda > 0.0f) {
    SGA_TRY(conv_f32(kUp2Dgrad, c->gs[3], d.N, d.Hs[4], d.Ws[4], c->f_gx, false, 0, nullptr, c->f_ga[2], st));
    for (int i = 2; i >= 0; --i) {
      const long long pix = static_cast<long long>(d.N) * d.Hs[i + 1] * d.Ws[i + 1];
      SGA_TRY(igdn_bwd_f32(c->f_ga[i], c->f_xpre[i], c->f_s[i], c->gs_gamma_t[i], pix, d.C, c->f_tmp, c->f_nrm,
                           c->f_gb[i], st));
      float* out = (i == 0) ? c->gy_dist : c->f_ga[i - 1];
      SGA_TRY(conv_f32(kUp2Dgrad, c->gs[i], d.N, d.Hs[i + 1], d.Ws[i + 1], c->f_gb[i], false, 0, nullptr, out, st));
    }
  } else {
    SGA_CUDA(cudaMemsetAsync(c->gy_dist, 0, d.ny * sizeof(float), st));
  }
  SGA_TRY(conv_f32(kS1TrueDgrad, c->hs[2], d.N, d.Hh, d.Wh, c->f_gh2, false, 0, c->f_h1, c->f_gh1, st));
  SGA_TRY(conv_f32(kUp2Dgrad, c->hs[1], d.N, d.Hh, d.Wh, c->f_gh1, false, 0, c->f_h0, c->f_gh0, st));
  SGA_TRY(conv_f32(kUp2Dgrad, c->hs[0], d.N, 2 * d.Hz, 2 * d.Wz, c->f_gh0, false, 0, nullptr, c->gz_hyper, st));
  return SGA_OK;
}

// ---------------------------------------------------------------------------------------------------
// forward + backward, bf16 tcgen05 path
// ---------------------------------------------------------------------------------------------------
static int forward_backward_bf16(sga_ctx* c, const float* x, const float* y, const float* z, float T, int step,
                                 const float* ny, const float* nz, float* stats, cudaStream_t st,
                                 const BbArgs* bb = nullptr) {
  const Dims& d = c->d;
  const sga_config& g = c->cfg;
  const float rate_scale = static_cast<float>(1.0 / (log(2.0) * d.H * static_cast<double>(d.W) * g.n_ref));
  const bool rate_only = bb && bb->rate_only;
  const int y_hard = bb ? bb->y_hard : 0;
  SGA_CUDA(cudaMemsetAsync(c->acc, 0, 5 * d.N * sizeof(double), st));
  if (bb) {
    PROF(c, SGA_PROF_ELEMENTWISE, st, launch_zbb_path<bf16>(bb->z_mean, bb->z_logvar, bb->eps, d.nz, d.nz / d.N, d.C, c->eb_pack, g.seed,
                                step, rate_scale, c->b_zt, c->gz_rate, c->dz, c->acc + d.N, c->acc + 4 * d.N,
                                g.first_image * (d.nz / d.N), st));
  } else {
  PROF(c, SGA_PROF_ELEMENTWISE, st, launch_z_path<bf16>(z, nz, d.nz, d.nz / d.N, d.C, c->eb_pack, T, g.clip_eps, g.seed, step, 0, rate_scale,
                              c->b_zt, c->gz_rate, c->dz, c->acc + d.N, g.first_image * (d.nz / d.N), st));
  }
  PROF(c, SGA_PROF_CONV_HS + 0, st, tc_conv_launch(c->tc_hs[0], c->hs[0].bias, 1, nullptr, c->b_h0, nullptr, st));
  PROF(c, SGA_PROF_CONV_HS + 1, st, tc_conv_launch(c->tc_hs[1], c->hs[1].bias, 1, nullptr, c->b_h1, nullptr, st));
  PROF(c, SGA_PROF_CONV_HS + 2, st, tc_conv_launch(c->tc_hs[2], c->hs[2].bias, 0, nullptr, nullptr, c->h2, st));
  PROF(c, SGA_PROF_ELEMENTWISE, st, launch_y_path<bf16>(y, ny, c->h2, d.N, d.Hy, d.Wy, d.Hh, d.Wh, d.C, T, g.clip_eps, g.seed, step, y_hard,
                              rate_scale, c->b_yt, c->gy_rate, c->dy, c->b_gh2, c->acc, g.first_image * (d.ny / d.N), st));
  for (int i = 0; i < 3 && !rate_only; ++i) {
    if (c->fused) {
      PROF(c, SGA_PROF_CONV_GS + i, st, tc_conv_launch_igdn_fwd(c->tc_gs[i], c->gs[i].bias, c->gs_beta[i], c->b_s[i],
                                                                c->b_act[i], st));
    } else {
      PROF(c, SGA_PROF_CONV_GS + i, st, tc_conv_launch(c->tc_gs[i], c->gs[i].bias, 0, nullptr, c->b_xpre[i], nullptr, st));
      PROF(c, SGA_PROF_IGDN_FWD, st, igdn_tc_forward(c->ig_f[i], c->b_xpre[i], c->gs_beta[i], c->b_act[i], c->b_s[i], st));
    }
  }
  const float grad_scale =
      (g.lmbda > 0.0f) ? static_cast<float>(2.0 * 65025.0 * g.lmbda / (static_cast<double>(g.n_ref) * d.H * d.W * 3.0))
                       : 0.0f;
  if (!rate_only)
    PROF(c, SGA_PROF_L3_FWD, st, l3_tc_forward(c->l3, c->gs3_bias_host, x, grad_scale, nullptr, c->b_gx4, c->acc + 2 * d.N, c->acc + 3 * d.N,
                          st));
  if (stats) {
    if (bb) PROF(c, SGA_PROF_ELEMENTWISE, st, launch_finalize_stats_bb(c->acc, d.N, g.n_ref, d.H, d.W, g.lmbda, rate_only ? 1 : 0, stats, st));
    else PROF(c, SGA_PROF_ELEMENTWISE, st, launch_finalize_stats(c->acc, d.N, g.n_ref, d.H, d.W, g.lmbda, stats, st));
  }
  // ---- backward
  if (rate_only) {
    // y_tilde is fixed: only the hyper path carries gradient
  } else if (g.lmbda > 0.0f) {
    if (c->fused) {
      // every IGDN backward runs in the epilogue of the dgrad that produces its upstream gradient: IGDN_2 inside the
      // layer-3 dgrad, IGDN_1 / IGDN_0 inside the dgrad convs of layers 2 / 1
      PROF(c, SGA_PROF_L3_DGRAD, st, l3_tc_dgrad_igdn(c->l3, c->b_gx4, c->b_gb[2], st));
      PROF(c, SGA_PROF_CONV_GS + 5, st, tc_conv_launch_igdn_bwd(c->tc_gs_d[2], c->b_act[1], c->b_s[1], c->b_gb[1], st));
      PROF(c, SGA_PROF_CONV_GS + 4, st, tc_conv_launch_igdn_bwd(c->tc_gs_d[1], c->b_act[0], c->b_s[0], c->b_gb[0], st));
      PROF(c, SGA_PROF_CONV_GS + 3, st, tc_conv_launch(c->tc_gs_d[0], nullptr, 0, nullptr, nullptr, c->gy_dist, st));
    } else {
      PROF(c, SGA_PROF_L3_DGRAD, st, l3_tc_dgrad(c->l3, c->b_gx4, c->b_ga[2], st));
      for (int i = 2; i >= 0; --i) {
        PROF(c, SGA_PROF_IGDN_BWD, st, igdn_tc_backward(c->ig_b[i], c->b_ga[i], c->b_xpre[i], c->b_s[i], c->b_gb[i], st));
        if (i == 0) PROF(c, SGA_PROF_CONV_GS + 3, st, tc_conv_launch(c->tc_gs_d[0], nullptr, 0, nullptr, nullptr, c->gy_dist, st));
        else PROF(c, SGA_PROF_CONV_GS + 3 + i, st, tc_conv_launch(c->tc_gs_d[i], nullptr, 0, nullptr, c->b_ga[i - 1], nullptr, st));
      }
    }
  } else {
    SGA_CUDA(cudaMemsetAsync(c->gy_dist, 0, d.ny * sizeof(float), st));
  }
  PROF(c, SGA_PROF_CONV_HS + 5, st, tc_conv_launch(c->tc_hs_d[2], nullptr, 0, c->b_h1, c->b_gh1, nullptr, st));
  PROF(c, SGA_PROF_CONV_HS + 4, st, tc_conv_launch(c->tc_hs_d[1], nullptr, 0, c->b_h0, c->b_gh0, nullptr, st));
  PROF(c, SGA_PROF_CONV_HS + 3, st, tc_conv_launch(c->tc_hs_d[0], nullptr, 0, nullptr, nullptr, c->gz_hyper, st));
  return SGA_OK;
}

static int grad_common(sga_ctx* c, const float* x, const float* y, const float* z, float T, int step, const float* ny,
                       const float* nz, float* stats, cudaStream_t st) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(T > 0.0f, "temperature must be positive");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_SGA, "bits-back contexts are driven through sga_bb_grad / sga_bb_step");
  if (c->cfg.operand == SGA_OPERAND_BF16) return forward_backward_bf16(c, x, y, z, T, step, ny, nz, stats, st);
  SGA_TRY(forward_f32(c, x, y, z, T, step, ny, nz, 0, true, stats, st));
  return backward_f32(c, st);
}

}  // namespace sga

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* sga_last_error(void) { return sga::g_err; }
int sga_version(void) { return 100; }

int sga_create(const sga_config* cfg, sga_ctx** out) {
  if (!cfg || !out) { set_error("null argument"); return SGA_ERR_INVALID; }
  SGA_CHECK(cfg->n_images > 0 && cfg->height > 0 && cfg->width > 0, "bad image shape");
  SGA_CHECK(cfg->num_filters >= 8 && cfg->num_filters % 2 == 0, "num_filters must be even and >= 8");
  SGA_CHECK(cfg->n_ref >= 1, "n_ref must be >= 1");
  SGA_CHECK(cfg->first_image >= 0, "first_image must be >= 0");
  SGA_CHECK(cfg->operand == SGA_OPERAND_F32 || cfg->operand == SGA_OPERAND_BF16, "unknown operand type");
  SGA_CHECK(cfg->variant == SGA_VARIANT_SGA || cfg->variant == SGA_VARIANT_BITS_BACK, "unknown variant %d", cfg->variant);
  if (cfg->operand == SGA_OPERAND_BF16)
    SGA_CHECK(cfg->num_filters % 64 == 0 && cfg->num_filters <= 256,
              "the bf16 tcgen05 path needs num_filters %% 64 == 0 and <= 256 (got %d)", cfg->num_filters);
  sga_ctx* c = new sga_ctx();
  c->cfg = *cfg;
  compute_dims(*cfg, c->d);
  const Dims& d = c->d;
  int st = SGA_OK;
  auto A = [&](auto** p, long long n) { if (st == SGA_OK) st = c->alloc(p, n); };
  const long long nh0 = static_cast<long long>(d.N) * 2 * d.Hz * 2 * d.Wz * d.C;
  const long long nh1 = static_cast<long long>(d.N) * d.Hh * d.Wh * d.mid;
  const long long nh2 = static_cast<long long>(d.N) * d.Hh * d.Wh * 2 * d.C;
  const long long nxt = static_cast<long long>(d.N) * d.Hs[4] * d.Ws[4] * 3;
  const long long e3 = gs_elems(d, 3);
  A(&c->acc, 5 * d.N);
  if (cfg->variant == SGA_VARIANT_BITS_BACK) A(&c->bb_ha, 2 * d.nz);
  A(&c->dz, d.nz); A(&c->gz_rate, d.nz); A(&c->gz_hyper, d.nz);
  A(&c->dy, d.ny); A(&c->gy_rate, d.ny); A(&c->gy_dist, d.ny);
  A(&c->h2, nh2);
  // fp32 forward (init / eval / fp32 steps)
  A(&c->f_zt, d.nz); A(&c->f_h0, nh0); A(&c->f_h1, nh1); A(&c->f_yt, d.ny); A(&c->f_gh2, nh2);
  for (int i = 0; i < 3; ++i) { A(&c->f_xpre[i], gs_elems(d, i + 1)); A(&c->f_s[i], gs_elems(d, i + 1)); A(&c->f_act[i], gs_elems(d, i + 1)); }
  A(&c->f_xt, nxt); A(&c->f_gx, nxt);
  const long long ninit = static_cast<long long>(d.N) * d.Hg[1] * d.Wg[1] * d.C;
  const long long nscratch = e3 > ninit ? e3 : ninit;
  A(&c->f_tmp, nscratch); A(&c->f_nrm, nscratch);
  A(&c->f_init[0], ninit); A(&c->f_init[1], ninit);
  for (int i = 0; i < 3; ++i) { c->f_ga[i] = nullptr; c->f_gb[i] = nullptr; }
  c->f_gh1 = c->f_gh0 = nullptr;
  if (cfg->operand == SGA_OPERAND_F32) {
    for (int i = 0; i < 3; ++i) { A(&c->f_ga[i], gs_elems(d, i + 1)); A(&c->f_gb[i], gs_elems(d, i + 1)); }
    A(&c->f_gh1, nh1); A(&c->f_gh0, nh0);
  } else {
    A(&c->b_zt, d.nz); A(&c->b_h0, nh0); A(&c->b_h1, nh1); A(&c->b_yt, d.ny); A(&c->b_gh2, nh2);
    for (int i = 0; i < 3; ++i) {
      A(&c->b_xpre[i], gs_elems(d, i + 1)); A(&c->b_s[i], gs_elems(d, i + 1)); A(&c->b_act[i], gs_elems(d, i + 1));
      A(&c->b_ga[i], gs_elems(d, i + 1)); A(&c->b_gb[i], gs_elems(d, i + 1));
    }
    A(&c->b_gx4, static_cast<long long>(d.N) * d.Hs[4] * d.Ws[4] * 4);
    A(&c->b_gh1, nh1); A(&c->b_gh0, nh0);
  }
  if (st != SGA_OK) { sga_destroy(c); return st; }
  *out = c;
  return SGA_OK;
}

int sga_destroy(sga_ctx* c) {
  if (!c) return SGA_OK;
  for (int i = 0; i < 3; ++i) {
    tc_conv_destroy(c->tc_hs[i]); tc_conv_destroy(c->tc_gs[i]); tc_conv_destroy(c->tc_hs_d[i]); tc_conv_destroy(c->tc_gs_d[i]);
    igdn_tc_destroy(c->ig_f[i]); igdn_tc_destroy(c->ig_b[i]);
  }
  l3_tc_destroy(c->l3);
  for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  for (void* p : c->allocs) cudaFree(p);
  delete c;
  return SGA_OK;
}

size_t sga_workspace_bytes(const sga_ctx* c) { return c ? c->bytes : 0; }

int sga_latent_shape(const sga_ctx* c, int32_t* hy, int32_t* wy, int32_t* hz, int32_t* wz) {
  if (!c) { set_error("null context"); return SGA_ERR_INVALID; }
  if (hy) *hy = c->d.Hy;
  if (wy) *wy = c->d.Wy;
  if (hz) *hz = c->d.Hz;
  if (wz) *wz = c->d.Wz;
  return SGA_OK;
}

int sga_bind_weights(sga_ctx* c, const sga_weights* w, void* stream) {
  if (!c || !w) { set_error("null argument"); return SGA_ERR_INVALID; }
  SGA_CHECK(!c->bound, "weights already bound to this context");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Dims& d = c->d;
  const int C = d.C;
  const bool bf = (c->cfg.operand == SGA_OPERAND_BF16);
  const int ga_ci[4] = {3, C, C, C};
  for (int i = 0; i < 4; ++i) SGA_TRY(bind_conv(c, c->ga[i], w->ga_kernel[i], w->ga_bias[i], 5, 5, ga_ci[i], C, false, st));
  const int gs_co[4] = {C, C, C, 3};
  for (int i = 0; i < 4; ++i) SGA_TRY(bind_conv(c, c->gs[i], w->gs_kernel[i], w->gs_bias[i], 5, 5, C, gs_co[i], bf && i < 3, st));
  SGA_TRY(bind_conv(c, c->ha[0], w->ha_kernel[0], w->ha_bias[0], 3, 3, C, C, false, st));
  SGA_TRY(bind_conv(c, c->ha[1], w->ha_kernel[1], w->ha_bias[1], 5, 5, C, C, false, st));
  // bits-back variant: h_a emits 2C channels, z_mean | z_logvar (bb_sga.py:69,99)
  SGA_TRY(bind_conv(c, c->ha[2], w->ha_kernel[2], nullptr, 5, 5, C, c->cfg.variant == SGA_VARIANT_BITS_BACK ? 2 * C : C, false, st));
  SGA_TRY(bind_conv(c, c->hs[0], w->hs_kernel[0], w->hs_bias[0], 5, 5, C, C, bf, st));
  SGA_TRY(bind_conv(c, c->hs[1], w->hs_kernel[1], w->hs_bias[1], 5, 5, C, d.mid, bf, st));
  SGA_TRY(bind_conv(c, c->hs[2], w->hs_kernel[2], w->hs_bias[2], 3, 3, d.mid, 2 * C, bf, st));
  for (int i = 0; i < 3; ++i) {
    SGA_TRY(c->alloc(&c->ga_beta[i], C)); SGA_TRY(c->alloc(&c->ga_gamma[i], C * C));
    SGA_TRY(c->alloc(&c->gs_beta[i], C)); SGA_TRY(c->alloc(&c->gs_gamma[i], C * C)); SGA_TRY(c->alloc(&c->gs_gamma_t[i], C * C));
    SGA_CUDA(cudaMemcpyAsync(c->ga_beta[i], w->ga_beta[i], C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SGA_CUDA(cudaMemcpyAsync(c->ga_gamma[i], w->ga_gamma[i], C * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SGA_CUDA(cudaMemcpyAsync(c->gs_beta[i], w->gs_beta[i], C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SGA_CUDA(cudaMemcpyAsync(c->gs_gamma[i], w->gs_gamma[i], C * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SGA_TRY(launch_transpose_taps(c->gs_gamma[i], 1, C, C, c->gs_gamma_t[i], st));
    if (bf) {
      SGA_TRY(c->alloc(&c->gs_gamma_b[i], C * C)); SGA_TRY(c->alloc(&c->gs_gamma_tb[i], C * C));
      SGA_TRY(launch_f32_to_bf16(c->gs_gamma[i], C * C, c->gs_gamma_b[i], st));
      SGA_TRY(launch_f32_to_bf16(c->gs_gamma_t[i], C * C, c->gs_gamma_tb[i], st));
    }
  }
  SGA_TRY(c->alloc(&c->eb_pack, 44 * C));
  SGA_TRY(launch_eb_pack(w->eb_matrix, w->eb_bias, w->eb_factor, C, c->eb_pack, st));
  SGA_CUDA(cudaMemcpyAsync(c->gs3_bias_host, w->gs_bias[3], 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  SGA_CUDA(cudaStreamSynchronize(st));
  if (bf) {
    SGA_TRY(make_tc(kUp2, c->hs[0], d.N, d.Hz, d.Wz, c->b_zt, &c->tc_hs[0]));
    SGA_TRY(make_tc(kUp2, c->hs[1], d.N, 2 * d.Hz, 2 * d.Wz, c->b_h0, &c->tc_hs[1]));
    SGA_TRY(make_tc(kS1True, c->hs[2], d.N, d.Hh, d.Wh, c->b_h1, &c->tc_hs[2]));
    SGA_TRY(make_tc(kS1TrueDgrad, c->hs[2], d.N, d.Hh, d.Wh, c->b_gh2, &c->tc_hs_d[2]));
    SGA_TRY(make_tc(kUp2Dgrad, c->hs[1], d.N, d.Hh, d.Wh, c->b_gh1, &c->tc_hs_d[1]));
    SGA_TRY(make_tc(kUp2Dgrad, c->hs[0], d.N, 2 * d.Hz, 2 * d.Wz, c->b_gh0, &c->tc_hs_d[0]));
    c->fused = tc_conv_can_fuse(C) && getenv("SGA_NO_FUSE") == nullptr;
    for (int i = 0; i < 3; ++i) {
      const bf16* in = (i == 0) ? c->b_yt : c->b_act[i - 1];
      const long long pix = static_cast<long long>(d.N) * d.Hs[i + 1] * d.Ws[i + 1];
      if (c->fused) {
        SGA_TRY(make_tc(kUp2, c->gs[i], d.N, d.Hs[i], d.Ws[i], in, &c->tc_gs[i], 1, c->gs_gamma_tb[i]));
        // the dgrad of layer i produces the upstream gradient of IGDN_{i-1}
        if (i > 0) SGA_TRY(make_tc(kUp2Dgrad, c->gs[i], d.N, d.Hs[i + 1], d.Ws[i + 1], c->b_gb[i], &c->tc_gs_d[i], 2,
                                   c->gs_gamma_b[i - 1]));
        else SGA_TRY(make_tc(kUp2Dgrad, c->gs[i], d.N, d.Hs[i + 1], d.Ws[i + 1], c->b_gb[i], &c->tc_gs_d[i]));
      } else {
        SGA_TRY(make_tc(kUp2, c->gs[i], d.N, d.Hs[i], d.Ws[i], in, &c->tc_gs[i]));
        SGA_TRY(make_tc(kUp2Dgrad, c->gs[i], d.N, d.Hs[i + 1], d.Ws[i + 1], c->b_gb[i], &c->tc_gs_d[i]));
        SGA_TRY(igdn_tc_create(c->gs_gamma_tb[i], C, pix, 0, &c->ig_f[i]));
        SGA_TRY(igdn_tc_create(c->gs_gamma_b[i], C, pix, 1, &c->ig_b[i]));
      }
    }
    SGA_TRY(c->alloc(&c->l3_wscatter, 80 * C)); SGA_TRY(c->alloc(&c->l3_wgather, 80 * C));
    SGA_TRY(l3_pack_weights(c->gs[3].w, C, c->l3_wscatter, c->l3_wgather, st));
    // fused: the IGDN backward recovers x_2 from the saved act_2 and s_2 (the fused forward does not store x)
    SGA_TRY(l3_tc_create(c->b_act[2], c->l3_wscatter, c->l3_wgather, c->fused ? c->gs_gamma_b[2] : nullptr,
                         c->fused ? c->b_act[2] : c->b_xpre[2], c->b_s[2], d.N, d.Hs[3], d.Ws[3], C, d.H, d.W, &c->l3));
    SGA_CUDA(cudaStreamSynchronize(st));
  }
  c->bound = true;
  return SGA_OK;
}

int sga_encode_init(sga_ctx* c, const float* x, float* y_out, float* z_out, void* stream) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Dims& d = c->d;
  // g_a: 4 x (5x5 corr, stride 2) with GDN after the first three (nn_models.py:13-36)
  const float* in = x;
  for (int i = 0; i < 4; ++i) {
    float* conv_out = (i == 3) ? y_out : c->f_init[0];
    SGA_TRY(conv_f32(kDown2Corr, c->ga[i], d.N, d.Hg[i], d.Wg[i], in, true, 0, nullptr, conv_out, st));
    if (i < 3) {
      const long long pix = static_cast<long long>(d.N) * d.Hg[i + 1] * d.Wg[i + 1];
      SGA_TRY(gdn_f32(c, 0, c->f_init[0], c->ga_beta[i], c->ga_gamma[i], pix, d.C, c->f_tmp, c->f_nrm, c->f_init[1],
                      nullptr, st));
      in = c->f_init[1];
    }
  }
  // h_a: 3x3 s1 ReLU, 5x5 s2 ReLU, 5x5 s2 without bias (nn_models.py:84-103)
  if (c->cfg.variant == SGA_VARIANT_BITS_BACK) {
    SGA_CHECK(z_out == nullptr, "bits-back variant: z_mean / z_logvar come from sga_bb_hyper_init (pass z_out = NULL)");
    return SGA_OK;
  }
  SGA_TRY(conv_f32(kS1Corr, c->ha[0], d.N, d.Hy, d.Wy, y_out, true, 1, nullptr, c->f_init[0], st));
  SGA_TRY(conv_f32(kDown2Corr, c->ha[1], d.N, d.Hy, d.Wy, c->f_init[0], true, 1, nullptr, c->f_init[1], st));
  SGA_TRY(conv_f32(kDown2Corr, c->ha[2], d.N, d.Hz1, d.Wz1, c->f_init[1], false, 0, nullptr, z_out, st));
  return SGA_OK;
}

// ---- bits-back variant (reference bb_sga.py) ---------------------------------------------------------------------

int sga_bb_hyper_init(sga_ctx* c, const float* y_tilde, float* z_mean_out, float* z_logvar_out, void* stream) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_BITS_BACK, "context was not created with variant = SGA_VARIANT_BITS_BACK");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Dims& d = c->d;
  SGA_TRY(conv_f32(kS1Corr, c->ha[0], d.N, d.Hy, d.Wy, y_tilde, true, 1, nullptr, c->f_init[0], st));
  SGA_TRY(conv_f32(kDown2Corr, c->ha[1], d.N, d.Hy, d.Wy, c->f_init[0], true, 1, nullptr, c->f_init[1], st));
  SGA_TRY(conv_f32(kDown2Corr, c->ha[2], d.N, d.Hz1, d.Wz1, c->f_init[1], false, 0, nullptr, c->bb_ha, st));
  return launch_split2(c->bb_ha, d.nz / d.C, d.C, z_mean_out, z_logvar_out, st);
}

static int bb_grad_common(sga_ctx* c, const float* x, const float* y, const BbArgs& bb, float T, int step, const float* ny,
                          float* stats, cudaStream_t st) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_BITS_BACK, "context was not created with variant = SGA_VARIANT_BITS_BACK");
  SGA_CHECK(bb.rate_only || T > 0.0f, "temperature must be positive");
  const Dims& d = c->d;
  if (c->cfg.operand == SGA_OPERAND_BF16) {
    SGA_TRY(forward_backward_bf16(c, x, y, nullptr, T, step, ny, nullptr, stats, st, &bb));
  } else {
    SGA_TRY(forward_f32(c, x, y, nullptr, T, step, ny, nullptr, 0, true, stats, st, &bb));
    SGA_TRY(backward_f32(c, st, bb.rate_only != 0));
  }
  // g_mean -> gz_rate, g_logvar -> dz (in place)
  const float half_scale = static_cast<float>(0.5 / (log(2.0) * d.H * static_cast<double>(d.W) * c->cfg.n_ref));
  return launch_bb_grads(c->gz_rate, c->gz_hyper, c->dz, d.nz, half_scale, c->gz_rate, c->dz, st);
}

int sga_bb_grad(sga_ctx* c, const float* x, const float* y, const float* z_mean, const float* z_logvar, float T,
                int32_t step, const float* noise_y, const float* eps_z, int32_t rate_only, float* grad_y,
                float* grad_z_mean, float* grad_z_logvar, float* stats, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BbArgs bb{z_mean, z_logvar, eps_z, rate_only ? 1 : 0, rate_only ? 1 : 0};
  SGA_TRY(bb_grad_common(c, x, y, bb, T, step, noise_y, stats, st));
  const sga_config& g = c->cfg;
  if (!rate_only && grad_y)
    SGA_TRY(launch_adam(nullptr, nullptr, nullptr, c->gy_rate, c->gy_dist, c->dy, c->d.ny, g.lr, g.beta1, g.beta2,
                        g.adam_eps, 0, grad_y, st));
  SGA_CUDA(cudaMemcpyAsync(grad_z_mean, c->gz_rate, c->d.nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SGA_CUDA(cudaMemcpyAsync(grad_z_logvar, c->dz, c->d.nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SGA_OK;
}

int sga_bb_step(sga_ctx* c, const float* x, float* y, float* z_mean, float* z_logvar, float* m_y, float* v_y, float* m_zm,
                float* v_zm, float* m_zl, float* v_zl, float T, int32_t step, int32_t adam_iterations, const float* noise_y,
                const float* eps_z, int32_t rate_only, double lr, float* stats, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BbArgs bb{z_mean, z_logvar, eps_z, rate_only ? 1 : 0, rate_only ? 1 : 0};
  SGA_TRY(bb_grad_common(c, x, y, bb, T, step, noise_y, stats, st));
  const sga_config& g = c->cfg;
  if (!rate_only)
    SGA_TRY(launch_adam(y, m_y, v_y, c->gy_rate, c->gy_dist, c->dy, c->d.ny, lr, g.beta1, g.beta2, g.adam_eps, adam_iterations,
                        nullptr, st));
  SGA_TRY(launch_adam(z_mean, m_zm, v_zm, c->gz_rate, nullptr, nullptr, c->d.nz, lr, g.beta1, g.beta2, g.adam_eps, adam_iterations,
                      nullptr, st));
  SGA_TRY(launch_adam(z_logvar, m_zl, v_zl, c->dz, nullptr, nullptr, c->d.nz, lr, g.beta1, g.beta2, g.adam_eps, adam_iterations,
                      nullptr, st));
  return SGA_OK;
}

int sga_bb_eval(sga_ctx* c, const float* x, const float* y_tilde, const float* z_mean, const float* z_logvar,
                int32_t step, const float* eps_z, float* stats, void* stream) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_BITS_BACK, "context was not created with variant = SGA_VARIANT_BITS_BACK");
  const BbArgs bb{z_mean, z_logvar, eps_z, 0, 1};
  return forward_f32(c, x, y_tilde, nullptr, 1.0f, step, nullptr, nullptr, 1, false, stats, static_cast<cudaStream_t>(stream),
                     &bb);
}

int sga_grad(sga_ctx* c, const float* x, const float* y, const float* z, float T, int32_t step, const float* ny,
             const float* nz, float* grad_y, float* grad_z, float* stats, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SGA_TRY(grad_common(c, x, y, z, T, step, ny, nz, stats, st));
  const sga_config& g = c->cfg;
  SGA_TRY(launch_adam(nullptr, nullptr, nullptr, c->gy_rate, c->gy_dist, c->dy, c->d.ny, g.lr, g.beta1, g.beta2,
                      g.adam_eps, 0, grad_y, st));
  SGA_TRY(launch_adam(nullptr, nullptr, nullptr, c->gz_rate, c->gz_hyper, c->dz, c->d.nz, g.lr, g.beta1, g.beta2,
                      g.adam_eps, 0, grad_z, st));
  return SGA_OK;
}

int sga_step(sga_ctx* c, const float* x, float* y, float* z, float* m_y, float* v_y, float* m_z, float* v_z, float T,
             int32_t step, const float* ny, const float* nz, float* stats, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SGA_TRY(grad_common(c, x, y, z, T, step, ny, nz, stats, st));
  const sga_config& g = c->cfg;
  PROF(c, SGA_PROF_ELEMENTWISE, st, launch_adam(y, m_y, v_y, c->gy_rate, c->gy_dist, c->dy, c->d.ny, g.lr, g.beta1,
                                                g.beta2, g.adam_eps, step, nullptr, st));
  PROF(c, SGA_PROF_ELEMENTWISE, st, launch_adam(z, m_z, v_z, c->gz_rate, c->gz_hyper, c->dz, c->d.nz, g.lr, g.beta1,
                                                g.beta2, g.adam_eps, step, nullptr, st));
  return SGA_OK;
}

int sga_eval_hard(sga_ctx* c, const float* x, const float* y, const float* z, float* stats, void* stream) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_SGA, "bits-back contexts are evaluated through sga_bb_eval");
  return forward_f32(c, x, y, z, 1.0f, 0, nullptr, nullptr, 1, false, stats, static_cast<cudaStream_t>(stream));
}

int sga_reconstruction(sga_ctx* c, float* xt_out, void* stream) {
  SGA_CHECK(c && c->bound && xt_out, "bad argument");
  const Dims& d = c->d;
  // crop [n, Hs4, Ws4, 3] -> [n, H, W, 3]  (sga.py:123)
  for (int n = 0; n < d.N; ++n)
    SGA_CUDA(cudaMemcpy2DAsync(xt_out + static_cast<size_t>(n) * d.H * d.W * 3, static_cast<size_t>(d.W) * 3 * sizeof(float),
                               c->f_xt + static_cast<size_t>(n) * d.Hs[4] * d.Ws[4] * 3,
                               static_cast<size_t>(d.Ws[4]) * 3 * sizeof(float), static_cast<size_t>(d.W) * 3 * sizeof(float),
                               d.H, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return SGA_OK;
}

long long sga_launch_count(void) { return sga::g_launches; }

int sga_profile_enable(sga_ctx* c, int32_t enable) {
  SGA_CHECK(c != nullptr, "null context");
  c->prof = enable != 0;
  return SGA_OK;
}

int sga_profile_read(sga_ctx* c, float* ms, int32_t* count, int32_t n_labels) {
  SGA_CHECK(c != nullptr && ms != nullptr, "null argument");
  for (auto& r : c->prof_recs) {
    SGA_CUDA(cudaEventSynchronize(r.b));
    float t = 0.0f;
    SGA_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    if (r.label >= 0 && r.label < n_labels) {
      ms[r.label] += t;
      if (count) count[r.label] += 1;
    }
    c->prof_pool.push_back(r.a);
    c->prof_pool.push_back(r.b);
  }
  c->prof_recs.clear();
  return SGA_OK;
}

// ---------------------------------------------------------------------------------------------------
// operator-level entry points (parity tests)
// ---------------------------------------------------------------------------------------------------
int sga_op_soft_round(const float* v, const float* u, int64_t numel, float T, float eps, uint64_t seed, int32_t step,
                      int32_t tensor_id, float* vt, float* dvt, void* stream) {
  SGA_CHECK(T > 0.0f, "temperature must be positive");
  return launch_soft_round(v, u, numel, T, eps, seed, step, tensor_id, vt, dvt, static_cast<cudaStream_t>(stream));
}

int sga_op_eb_likelihood(const float* zt, int64_t numel, int32_t C, const float* const matrix[4],
                         const float* const bias[4], const float* const factor[3], float* p, float* dp, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SGA_CHECK(C > 0 && numel % C == 0, "numel must be a multiple of channels");
  float* pk = nullptr;
  SGA_CUDA(cudaMalloc(&pk, 44 * C * sizeof(float)));
  int s = launch_eb_pack(matrix, bias, factor, C, pk, st);
  if (s == SGA_OK) s = launch_eb_likelihood(zt, numel, C, pk, p, dp, st);
  cudaStreamSynchronize(st);
  cudaFree(pk);
  return s;
}

int sga_op_gauss_likelihood(const float* yt, const float* mu, const float* ls, int64_t numel, float* p, float* dpdy,
                            float* dpdsb, void* stream) {
  return launch_gauss_likelihood(yt, mu, ls, numel, p, dpdy, dpdsb, static_cast<cudaStream_t>(stream));
}

int sga_op_adam(float* p, float* m, float* v, const float* g, int64_t numel, double lr, double beta1, double beta2,
                double eps, int32_t iterations, void* stream) {
  return launch_adam(p, m, v, g, nullptr, nullptr, numel, lr, beta1, beta2, eps, iterations, nullptr,
                     static_cast<cudaStream_t>(stream));
}

int sga_op_conv(int32_t kind, int32_t operand, const float* x, const float* kernel, const float* bias, int32_t n,
                int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t cin, int32_t cout, int32_t relu, float* out,
                void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SGA_CHECK(kind >= 0 && kind <= 5, "unknown conv kind %d", kind);
  ConvPlan plan;
  SGA_TRY(make_conv_plan(kind, kh, kw, n, h, w, cin, cout, &plan));
  const long long nw = static_cast<long long>(kh) * kw * cin * cout;
  const long long nin = static_cast<long long>(n) * h * w * plan.g.Cin;
  int s = SGA_OK;
  if (operand == SGA_OPERAND_F32) {
    float* wt = nullptr;
    const float* wp = kernel;
    if (plan.w_transposed) {
      SGA_CUDA(cudaMalloc(&wt, nw * sizeof(float)));
      s = launch_transpose_taps(kernel, kh * kw, cin, cout, wt, st);
      wp = wt;
    }
    if (s == SGA_OK) s = launch_conv_simt(plan, x, wp, bias, nullptr, relu, out, st);
    cudaStreamSynchronize(st);
    if (wt) cudaFree(wt);
    return s;
  }
  bf16 *xb = nullptr, *wb = nullptr;
  SGA_CUDA(cudaMalloc(&xb, (nin + 8) * sizeof(bf16)));
  SGA_CUDA(cudaMalloc(&wb, nw * sizeof(bf16)));
  TcConv* tcv = nullptr;
  s = launch_f32_to_bf16(x, nin, xb, st);
  if (s == SGA_OK) s = launch_pack_weights_bf16(kernel, kh * kw, cin, cout, plan.w_transposed, wb, st);
  if (s == SGA_OK) s = tc_conv_create(plan, xb, wb, 0, nullptr, &tcv);
  if (s == SGA_OK) s = tc_conv_launch(tcv, bias, relu, nullptr, nullptr, out, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (s == SGA_OK && e != cudaSuccess) { set_error("tcgen05 conv failed: %s", cudaGetErrorString(e)); s = SGA_ERR_CUDA; }
  tc_conv_destroy(tcv);
  cudaFree(xb); cudaFree(wb);
  return s;
}

int sga_op_gdn(int32_t operand, int32_t inverse, const float* x, const float* beta, const float* gamma, int64_t pixels,
               int32_t C, float* out, float* s_out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = pixels * C;
  int s = SGA_OK;
  if (operand == SGA_OPERAND_F32) {
    float *tmp = nullptr, *nrm = nullptr;
    SGA_CUDA(cudaMalloc(&tmp, n * sizeof(float)));
    SGA_CUDA(cudaMalloc(&nrm, n * sizeof(float)));
    s = gdn_f32(nullptr, inverse, x, beta, gamma, pixels, C, tmp, nrm, out, s_out, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp); cudaFree(nrm);
    return s;
  }
  SGA_CHECK(inverse, "the tcgen05 path only implements the inverse GDN used inside the loop");
  float* gt = nullptr;
  bf16 *gtb = nullptr, *xb = nullptr, *ob = nullptr, *sb = nullptr;
  SGA_CUDA(cudaMalloc(&gt, C * C * sizeof(float)));
  SGA_CUDA(cudaMalloc(&gtb, C * C * sizeof(bf16)));
  SGA_CUDA(cudaMalloc(&xb, n * sizeof(bf16)));
  SGA_CUDA(cudaMalloc(&ob, n * sizeof(bf16)));
  SGA_CUDA(cudaMalloc(&sb, n * sizeof(bf16)));
  IgdnTc* ig = nullptr;
  s = launch_transpose_taps(gamma, 1, C, C, gt, st);
  if (s == SGA_OK) s = launch_f32_to_bf16(gt, C * C, gtb, st);
  if (s == SGA_OK) s = launch_f32_to_bf16(x, n, xb, st);
  if (s == SGA_OK) s = igdn_tc_create(gtb, C, pixels, 0, &ig);
  if (s == SGA_OK) s = igdn_tc_forward(ig, xb, beta, ob, sb, st);
  if (s == SGA_OK) s = launch_bf16_to_f32(ob, n, out, st);
  if (s == SGA_OK && s_out) s = launch_bf16_to_f32(sb, n, s_out, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (s == SGA_OK && e != cudaSuccess) { set_error("tcgen05 igdn failed: %s", cudaGetErrorString(e)); s = SGA_ERR_CUDA; }
  igdn_tc_destroy(ig);
  cudaFree(gt); cudaFree(gtb); cudaFree(xb); cudaFree(ob); cudaFree(sb);
  return s;
}

int sga_op_igdn_bwd(int32_t operand, const float* g, const float* x, const float* s_in, const float* gamma,
                    int64_t pixels, int32_t C, float* gx, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = pixels * C;
  int s = SGA_OK;
  if (operand == SGA_OPERAND_F32) {
    float *tmp = nullptr, *u = nullptr, *gt = nullptr;
    SGA_CUDA(cudaMalloc(&tmp, n * sizeof(float)));
    SGA_CUDA(cudaMalloc(&u, n * sizeof(float)));
    SGA_CUDA(cudaMalloc(&gt, C * C * sizeof(float)));
    s = launch_transpose_taps(gamma, 1, C, C, gt, st);
    if (s == SGA_OK) s = igdn_bwd_f32(g, x, s_in, gt, pixels, C, tmp, u, gx, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp); cudaFree(u); cudaFree(gt);
    return s;
  }
  bf16 *gb = nullptr, *xb = nullptr, *sb = nullptr, *ob = nullptr, *gmb = nullptr;
  SGA_CUDA(cudaMalloc(&gb, n * sizeof(bf16)));
  SGA_CUDA(cudaMalloc(&xb, n * sizeof(bf16)));
  SGA_CUDA(cudaMalloc(&sb, n * sizeof(bf16)));
  SGA_CUDA(cudaMalloc(&ob, n * sizeof(bf16)));
  SGA_CUDA(cudaMalloc(&gmb, C * C * sizeof(bf16)));
  IgdnTc* ig = nullptr;
  s = launch_f32_to_bf16(g, n, gb, st);
  if (s == SGA_OK) s = launch_f32_to_bf16(x, n, xb, st);
  if (s == SGA_OK) s = launch_f32_to_bf16(s_in, n, sb, st);
  if (s == SGA_OK) s = launch_f32_to_bf16(gamma, C * C, gmb, st);
  if (s == SGA_OK) s = igdn_tc_create(gmb, C, pixels, 1, &ig);
  if (s == SGA_OK) s = igdn_tc_backward(ig, gb, xb, sb, ob, st);
  if (s == SGA_OK) s = launch_bf16_to_f32(ob, n, gx, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (s == SGA_OK && e != cudaSuccess) { set_error("tcgen05 igdn bwd failed: %s", cudaGetErrorString(e)); s = SGA_ERR_CUDA; }
  igdn_tc_destroy(ig);
  cudaFree(gb); cudaFree(xb); cudaFree(sb); cudaFree(ob); cudaFree(gmb);
  return s;
}

}  // extern "C"
