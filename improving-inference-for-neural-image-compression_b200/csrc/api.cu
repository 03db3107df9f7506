// api.cu -- the C ABI of libsga_b200.so (include/sga_b200.h): context, weight binding, and the per-iteration
// orchestration of the SGA graph (reference sga.py:88-174, 207-215, 240-247) on one CUDA stream.
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>
#include <vector>
#include <type_traits>

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
  int fwd_dup = 1;          // wb_fwd holds the weights fwd_dup times along K (hi|lo split input operand)
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
  TcConv* tc_ig2 = nullptr;     // C = 256: IGDN_2 backward as the fused epilogue of an identity 1x1 "conv" over d loss / d act_2
  bf16* ident_b = nullptr;
  L3Tc* l3 = nullptr;
  bool fused = false;   // IGDN runs inside the (de)conv epilogues (num_filters in {64,128,192})
  int split0 = 0;       // y_tilde / z_tilde enter g_s / h_s layer 0 as hi|lo bf16 pairs (K = 2C), see store_operand_pair

  // two-stream schedule + graph replay (sga_loop)
  cudaStream_t side = nullptr, main = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_soft = nullptr, ev_join = nullptr, ev_in = nullptr, ev_out = nullptr;
  static constexpr int kSpCap = 1024;      // iterations per upload of the step-parameter table
  StepParams* sp_table = nullptr;          // device
  int* sp_idx = nullptr;                   // device counter, advanced by the last kernel of a replayed step
  StepParams* sp_host = nullptr;           // pinned staging, 2 x kSpCap
  cudaEvent_t ev_sp[2] = {nullptr, nullptr};
  int sp_half = 0;
  float* stats_dev = nullptr;
  struct GraphKey { const void *x, *y, *z, *my, *vy, *mz, *vz; };
  struct GraphRec { GraphKey key; cudaGraphExec_t exec; long long kernels; };
  std::vector<GraphRec> graphs;

  // one arena for the activations: the fp32 set (init, final eval, fp32 steps) and the bf16 set (the loop) are never live
  // at the same time -- every step / evaluation rewrites all of its activations -- so they overlap
  char* arena = nullptr;
  size_t arena_bytes = 0;

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

// the layer-3 dgrad carries the IGDN_2 backward in its epilogue: with gamma, a saved x / s tile pair and the gather operands
// in shared memory for C <= 192, and in "direct" mode (x / s from global memory, one shared weights / t buffer) for C = 256
static bool l3_fusable(int C) {
  if (C == 256) return getenv("SGA_L3_NO_FUSE256") == nullptr;   // the direct-mode kernel; the switch keeps the fallback testable
  return C == 64 || C == 128 || C == 192;
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

// dup_fwd = 2: the forward operand of this layer is a hi|lo split tensor with 2*cin channels (weights stacked twice)
static int bind_conv(sga_ctx* c, ConvW& w, const float* kernel, const float* bias, int kh, int kw, int cin, int cout,
                     bool need_bf16, cudaStream_t st, int dup_fwd = 1) {
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
    SGA_TRY(c->alloc(&w.wb_fwd, n * dup_fwd));
    SGA_TRY(c->alloc(&w.wb_dgrad, n));
    w.fwd_dup = dup_fwd;
    SGA_TRY(launch_pack_weights_bf16(w.w, kh * kw, cin, cout, false, w.wb_fwd, st, dup_fwd));
    SGA_TRY(launch_pack_weights_bf16(w.w, kh * kw, cin, cout, true, w.wb_dgrad, st));
  }
  return SGA_OK;
}

static int make_tc(int kind, const ConvW& w, int N, int Hin, int Win, const bf16* x, TcConv** out, int fuse = 0,
                   const bf16* mix_op = nullptr) {
  ConvPlan plan;
  const bool fwd_kind = (kind == kUp2 || kind == kS1True);
  SGA_TRY(make_conv_plan(kind, w.kh, w.kw, N, Hin, Win, fwd_kind ? w.cin * w.fwd_dup : w.cin, w.cout, &plan));
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
  if (x == nullptr) return SGA_OK;          // decoder side: reconstruction only (sga_reconstruction reads f_xt)
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
      if (l3_fusable(d.C)) {
        PROF(c, SGA_PROF_L3_DGRAD, st, l3_tc_dgrad_igdn(c->l3, c->b_gx4, c->b_gb[2], st));
      } else {
        PROF(c, SGA_PROF_L3_DGRAD, st, l3_tc_dgrad(c->l3, c->b_gx4, c->b_ga[2], st));
        PROF(c, SGA_PROF_IGDN_BWD, st, tc_conv_launch_igdn_bwd(c->tc_ig2, c->b_act[2], c->b_s[2], c->b_gb[2], st));
      }
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

// ---------------------------------------------------------------------------------------------------
// The bf16 iteration of sga.py:211-215 as a two-stream schedule.  The synthesis chain (94 % of the FLOPs) only needs
// y_tilde, so it starts right after the soft round of y; the hyper path (z soft round + prior rate -> h_s -> Gaussian rate
// of y -> h_s dgrads -> Adam on z) runs beside it on the context's side stream and joins before Adam on y:
//
//   main : memset acc . fork . y_soft .(ev_soft). gs0 gs1 gs2 fwd . l3 fwd . l3 dgrad . gs2 gs1 gs0 dgrad . join . adam(y) . stats
//   side :               z_path hs0 hs1 hs2 fwd . wait(ev_soft) . y_rate . hs2 hs1 hs0 dgrad . adam(z)
//
// The same function is what sga_loop captures into a CUDA graph (spt / spi: temperature, noise step and lr_t then come
// from the device table).  With per-launch profiling on, everything is issued on the one stream so that the event pairs
// bracket single kernels.
// ---------------------------------------------------------------------------------------------------
struct StepOut {
  float *y, *z, *m_y, *v_y, *m_z, *v_z;   // in-place Adam update (sga_step / sga_loop) ...
  float *grad_y, *grad_z;                 // ... or the gradients only (sga_grad)
};

static int step_bf16(sga_ctx* c, const float* x, const float* y, const float* z, float T, int step, const float* ny,
                     const float* nz, float* stats, cudaStream_t st, const StepOut& o, const StepParams* spt = nullptr,
                     const int* spi = nullptr, int* spi_adv = nullptr) {
  const Dims& d = c->d;
  const sga_config& g = c->cfg;
  const float rate_scale = static_cast<float>(1.0 / (log(2.0) * d.H * static_cast<double>(d.W) * g.n_ref));
  const bool two = !c->prof && c->side != nullptr && getenv("SGA_ONE_STREAM") == nullptr;
  cudaStream_t s2 = two ? c->side : st;
  const long long off_y = g.first_image * (d.ny / d.N), off_z = g.first_image * (d.nz / d.N);
  SGA_CUDA(cudaMemsetAsync(c->acc, 0, 5 * d.N * sizeof(double), st));
  if (two) {
    SGA_CUDA(cudaEventRecord(c->ev_fork, st));
    SGA_CUDA(cudaStreamWaitEvent(s2, c->ev_fork, 0));
  }
  PROF(c, SGA_PROF_ELEMENTWISE, st, launch_y_soft(y, ny, d.ny, d.C, T, g.clip_eps, g.seed, step, c->f_yt, c->b_yt, c->dy, off_y,
                                                  st, spt, spi, c->split0));
  if (two) SGA_CUDA(cudaEventRecord(c->ev_soft, st));
  // ---- hyper path, forward
  PROF(c, SGA_PROF_ELEMENTWISE, s2, launch_z_path<bf16>(z, nz, d.nz, d.nz / d.N, d.C, c->eb_pack, T, g.clip_eps, g.seed, step, 0,
                                                        rate_scale, c->b_zt, c->gz_rate, c->dz, c->acc + d.N, off_z, s2, spt, spi,
                                                        c->split0));
  PROF(c, SGA_PROF_CONV_HS + 0, s2, tc_conv_launch(c->tc_hs[0], c->hs[0].bias, 1, nullptr, c->b_h0, nullptr, s2));
  PROF(c, SGA_PROF_CONV_HS + 1, s2, tc_conv_launch(c->tc_hs[1], c->hs[1].bias, 1, nullptr, c->b_h1, nullptr, s2));
  PROF(c, SGA_PROF_CONV_HS + 2, s2, tc_conv_launch(c->tc_hs[2], c->hs[2].bias, 0, nullptr, nullptr, c->h2, s2));
  // ---- synthesis path, forward (main stream)
  for (int i = 0; i < 3; ++i) {
    if (c->fused) {
      PROF(c, SGA_PROF_CONV_GS + i, st, tc_conv_launch_igdn_fwd(c->tc_gs[i], c->gs[i].bias, c->gs_beta[i], c->b_s[i],
                                                                c->b_act[i], st));
    } else {
      PROF(c, SGA_PROF_CONV_GS + i, st, tc_conv_launch(c->tc_gs[i], c->gs[i].bias, 0, nullptr, c->b_xpre[i], nullptr, st));
      PROF(c, SGA_PROF_IGDN_FWD, st, igdn_tc_forward(c->ig_f[i], c->b_xpre[i], c->gs_beta[i], c->b_act[i], c->b_s[i], st));
    }
  }
  // ---- rate of y and the hyper path's backward (side stream)
  if (two) SGA_CUDA(cudaStreamWaitEvent(s2, c->ev_soft, 0));
  PROF(c, SGA_PROF_ELEMENTWISE, s2, launch_y_path<bf16>(nullptr, nullptr, c->h2, d.N, d.Hy, d.Wy, d.Hh, d.Wh, d.C, T, g.clip_eps,
                                                        g.seed, step, 0, rate_scale, nullptr, c->gy_rate, nullptr, c->b_gh2,
                                                        c->acc, off_y, s2, c->f_yt));
  PROF(c, SGA_PROF_CONV_HS + 5, s2, tc_conv_launch(c->tc_hs_d[2], nullptr, 0, c->b_h1, c->b_gh1, nullptr, s2));
  PROF(c, SGA_PROF_CONV_HS + 4, s2, tc_conv_launch(c->tc_hs_d[1], nullptr, 0, c->b_h0, c->b_gh0, nullptr, s2));
  PROF(c, SGA_PROF_CONV_HS + 3, s2, tc_conv_launch(c->tc_hs_d[0], nullptr, 0, nullptr, nullptr, c->gz_hyper, s2));
  if (o.z != nullptr) {
    PROF(c, SGA_PROF_ELEMENTWISE, s2, launch_adam(o.z, o.m_z, o.v_z, c->gz_rate, c->gz_hyper, c->dz, d.nz, g.lr, g.beta1, g.beta2,
                                                  g.adam_eps, step, nullptr, s2, spt, spi));
  } else if (o.grad_z != nullptr) {
    SGA_TRY(launch_adam(nullptr, nullptr, nullptr, c->gz_rate, c->gz_hyper, c->dz, d.nz, g.lr, g.beta1, g.beta2, g.adam_eps, 0,
                        o.grad_z, s2));
  }
  if (two) SGA_CUDA(cudaEventRecord(c->ev_join, s2));
  // ---- distortion and the synthesis path's backward (main stream)
  const float grad_scale =
      (g.lmbda > 0.0f) ? static_cast<float>(2.0 * 65025.0 * g.lmbda / (static_cast<double>(g.n_ref) * d.H * d.W * 3.0))
                       : 0.0f;
  PROF(c, SGA_PROF_L3_FWD, st, l3_tc_forward(c->l3, c->gs3_bias_host, x, grad_scale, nullptr, c->b_gx4, c->acc + 2 * d.N,
                                             c->acc + 3 * d.N, st));
  if (g.lmbda > 0.0f) {
    if (c->fused) {
      // every IGDN backward runs in the epilogue of the dgrad that produces its upstream gradient: IGDN_2 inside the
      // layer-3 dgrad, IGDN_1 / IGDN_0 inside the dgrad convs of layers 2 / 1
      if (l3_fusable(d.C)) {
        PROF(c, SGA_PROF_L3_DGRAD, st, l3_tc_dgrad_igdn(c->l3, c->b_gx4, c->b_gb[2], st));
      } else {   // C = 256: IGDN_2 backward as the fused epilogue of an identity 1x1 conv over d loss / d act_2
        PROF(c, SGA_PROF_L3_DGRAD, st, l3_tc_dgrad(c->l3, c->b_gx4, c->b_ga[2], st));
        PROF(c, SGA_PROF_IGDN_BWD, st, tc_conv_launch_igdn_bwd(c->tc_ig2, c->b_act[2], c->b_s[2], c->b_gb[2], st));
      }
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
  if (two) SGA_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
  if (o.y != nullptr) {
    PROF(c, SGA_PROF_ELEMENTWISE, st, launch_adam(o.y, o.m_y, o.v_y, c->gy_rate, c->gy_dist, c->dy, d.ny, g.lr, g.beta1, g.beta2,
                                                  g.adam_eps, step, nullptr, st, spt, spi));
  } else if (o.grad_y != nullptr) {
    SGA_TRY(launch_adam(nullptr, nullptr, nullptr, c->gy_rate, c->gy_dist, c->dy, d.ny, g.lr, g.beta1, g.beta2, g.adam_eps, 0,
                        o.grad_y, st));
  }
  if (stats != nullptr || spi_adv != nullptr)
    PROF(c, SGA_PROF_ELEMENTWISE, st, launch_finalize_stats(c->acc, d.N, g.n_ref, d.H, d.W, g.lmbda,
                                                            stats ? stats : c->stats_dev, st, spi_adv));
  return SGA_OK;
}

static int grad_common(sga_ctx* c, const float* x, const float* y, const float* z, float T, int step, const float* ny,
                       const float* nz, float* stats, cudaStream_t st, const StepOut& o) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(T > 0.0f, "temperature must be positive");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_SGA, "bits-back contexts are driven through sga_bb_grad / sga_bb_step");
  if (c->cfg.operand == SGA_OPERAND_BF16) return step_bf16(c, x, y, z, T, step, ny, nz, stats, st, o);
  SGA_TRY(forward_f32(c, x, y, z, T, step, ny, nz, 0, true, stats, st));
  SGA_TRY(backward_f32(c, st));
  const sga_config& g = c->cfg;
  if (o.y != nullptr) {
    SGA_TRY(launch_adam(o.y, o.m_y, o.v_y, c->gy_rate, c->gy_dist, c->dy, c->d.ny, g.lr, g.beta1, g.beta2, g.adam_eps, step,
                        nullptr, st));
    SGA_TRY(launch_adam(o.z, o.m_z, o.v_z, c->gz_rate, c->gz_hyper, c->dz, c->d.nz, g.lr, g.beta1, g.beta2, g.adam_eps, step,
                        nullptr, st));
  } else {
    SGA_TRY(launch_adam(nullptr, nullptr, nullptr, c->gy_rate, c->gy_dist, c->dy, c->d.ny, g.lr, g.beta1, g.beta2, g.adam_eps,
                        0, o.grad_y, st));
    SGA_TRY(launch_adam(nullptr, nullptr, nullptr, c->gz_rate, c->gz_hyper, c->dz, c->d.nz, g.lr, g.beta1, g.beta2, g.adam_eps,
                        0, o.grad_z, st));
  }
  return SGA_OK;
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
  const bool bf = (cfg->operand == SGA_OPERAND_BF16);
  c->fused = bf && tc_conv_can_fuse(d.C) && getenv("SGA_NO_FUSE") == nullptr;
  {
    // opt-in (SGA_SPLIT0=1): measured no accuracy gain even at |y| ~ 120 (test_grad_bf16_large_latents_and_the_split_operand)
    // for +55 % time of g_s layer 0
    const char* es = getenv("SGA_SPLIT0");
    c->split0 = (bf && cfg->variant == SGA_VARIANT_SGA && es != nullptr && es[0] == '1') ? 1 : 0;
  }
  // ---- live across phases
  A(&c->acc, 5 * d.N);
  if (cfg->variant == SGA_VARIANT_BITS_BACK) A(&c->bb_ha, 2 * d.nz);
  A(&c->dz, d.nz); A(&c->gz_rate, d.nz); A(&c->gz_hyper, d.nz);
  A(&c->dy, d.ny); A(&c->gy_rate, d.ny); A(&c->gy_dist, d.ny);
  A(&c->h2, nh2);
  A(&c->f_yt, d.ny);                       // fp32 y_tilde: input of the fp32 forward and of the bf16 loop's rate kernel
  A(&c->stats_dev, SGA_STATS_HEADER + 6 * d.N);
  A(&c->sp_table, sga_ctx::kSpCap); A(&c->sp_idx, 1);
  // ---- the two overlapping activation sets: two passes over the same list, first sizing, then carving
  const long long ninit = static_cast<long long>(d.N) * d.Hg[1] * d.Wg[1] * d.C;
  const long long nscratch = e3 > ninit ? e3 : ninit;
  for (int i = 0; i < 3; ++i) { c->f_ga[i] = nullptr; c->f_gb[i] = nullptr; c->b_xpre[i] = nullptr; c->b_ga[i] = nullptr; }
  c->f_gh1 = c->f_gh0 = nullptr;
  for (int pass = 0; pass < 2 && st == SGA_OK; ++pass) {
    size_t off = 0;
    auto R = [&](auto** p, long long n) {
      typedef typename std::remove_pointer<typename std::remove_pointer<decltype(p)>::type>::type T;
      const size_t nb = ((static_cast<size_t>(n > 0 ? n : 1) * sizeof(T)) + 1023u) & ~static_cast<size_t>(1023u);
      if (pass == 1) *p = reinterpret_cast<T*>(c->arena + off);
      off += nb;
    };
    // fp32 set
    R(&c->f_zt, d.nz); R(&c->f_h0, nh0); R(&c->f_h1, nh1); R(&c->f_gh2, nh2);
    for (int i = 0; i < 3; ++i) { R(&c->f_xpre[i], gs_elems(d, i + 1)); R(&c->f_s[i], gs_elems(d, i + 1)); R(&c->f_act[i], gs_elems(d, i + 1)); }
    R(&c->f_xt, nxt); R(&c->f_gx, nxt);
    R(&c->f_tmp, nscratch); R(&c->f_nrm, nscratch);
    R(&c->f_init[0], ninit); R(&c->f_init[1], ninit);
    if (!bf) {
      for (int i = 0; i < 3; ++i) { R(&c->f_ga[i], gs_elems(d, i + 1)); R(&c->f_gb[i], gs_elems(d, i + 1)); }
      R(&c->f_gh1, nh1); R(&c->f_gh0, nh0);
    }
    const size_t f32_bytes = off;
    off = 0;
    if (bf) {
      const int dup = c->split0 ? 2 : 1;
      R(&c->b_zt, dup * d.nz); R(&c->b_h0, nh0); R(&c->b_h1, nh1); R(&c->b_yt, dup * d.ny); R(&c->b_gh2, nh2);
      for (int i = 0; i < 3; ++i) {
        R(&c->b_s[i], gs_elems(d, i + 1)); R(&c->b_act[i], gs_elems(d, i + 1)); R(&c->b_gb[i], gs_elems(d, i + 1));
        if (!c->fused) { R(&c->b_xpre[i], gs_elems(d, i + 1)); R(&c->b_ga[i], gs_elems(d, i + 1)); }
        else if (i == 2 && !l3_fusable(d.C)) R(&c->b_ga[2], gs_elems(d, 3));   // C = 256: layer-3 dgrad + IGDN_2 backward stay apart
      }
      R(&c->b_gx4, static_cast<long long>(d.N) * d.Hs[4] * d.Ws[4] * 4);
      R(&c->b_gh1, nh1); R(&c->b_gh0, nh0);
    }
    if (pass == 0) {
      c->arena_bytes = f32_bytes > off ? f32_bytes : off;
      A(&c->arena, static_cast<long long>(c->arena_bytes));
    }
  }
  if (st == SGA_OK && bf) {
    // The library's own main stream gets the highest priority and the side stream the lowest: with SGA_PRIO=1 sga_loop runs
    // the synthesis chain on it, so that the hyper path's blocks only fill what the synthesis kernels leave free.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaError_t e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_lo);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->main, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_soft, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    for (int h = 0; h < 2 && e == cudaSuccess; ++h) e = cudaEventCreateWithFlags(&c->ev_sp[h], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost(&c->sp_host, 2 * sga_ctx::kSpCap * sizeof(StepParams));
    if (e != cudaSuccess) { set_error("stream / event setup failed: %s", cudaGetErrorString(e)); st = SGA_ERR_CUDA; }
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
  tc_conv_destroy(c->tc_ig2);
  for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->main) cudaStreamDestroy(c->main);
  for (cudaEvent_t e : {c->ev_fork, c->ev_soft, c->ev_join, c->ev_in, c->ev_out, c->ev_sp[0], c->ev_sp[1]}) if (e) cudaEventDestroy(e);
  if (c->sp_host) cudaFreeHost(c->sp_host);
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
  for (int i = 0; i < 4; ++i)
    SGA_TRY(bind_conv(c, c->gs[i], w->gs_kernel[i], w->gs_bias[i], 5, 5, C, gs_co[i], bf && i < 3, st, (i == 0 && c->split0) ? 2 : 1));
  SGA_TRY(bind_conv(c, c->ha[0], w->ha_kernel[0], w->ha_bias[0], 3, 3, C, C, false, st));
  SGA_TRY(bind_conv(c, c->ha[1], w->ha_kernel[1], w->ha_bias[1], 5, 5, C, C, false, st));
  // bits-back variant: h_a emits 2C channels, z_mean | z_logvar (bb_sga.py:69,99)
  SGA_TRY(bind_conv(c, c->ha[2], w->ha_kernel[2], nullptr, 5, 5, C, c->cfg.variant == SGA_VARIANT_BITS_BACK ? 2 * C : C, false, st));
  SGA_TRY(bind_conv(c, c->hs[0], w->hs_kernel[0], w->hs_bias[0], 5, 5, C, C, bf, st, c->split0 ? 2 : 1));
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
    if (c->fused && !l3_fusable(C)) {
      // The layer-3 dgrad cannot carry IGDN_2's backward at C = 256 (shared memory).  Rather than the stand-alone IGDN kernel
      // (ALU-bound: 3.6 ms for the 604 M elements), run it as the fused backward epilogue of k_conv_tc over an identity
      // 1x1 convolution of d loss / d act_2: the accumulator then holds g exactly, and the epilogue + mixing GEMM are the
      // same tensor-rate code as in the other dgrads.
      std::vector<float> eye(static_cast<size_t>(C) * C, 0.0f);
      for (int i = 0; i < C; ++i) eye[static_cast<size_t>(i) * C + i] = 1.0f;
      float* eye_d = nullptr;
      SGA_TRY(c->alloc(&eye_d, static_cast<long long>(C) * C));
      SGA_TRY(c->alloc(&c->ident_b, static_cast<long long>(C) * C));
      SGA_CUDA(cudaMemcpyAsync(eye_d, eye.data(), eye.size() * sizeof(float), cudaMemcpyHostToDevice, st));
      SGA_TRY(launch_f32_to_bf16(eye_d, static_cast<long long>(C) * C, c->ident_b, st));
      SGA_CUDA(cudaStreamSynchronize(st));
      ConvPlan plan;
      SGA_TRY(make_conv_plan(kS1Corr, 1, 1, d.N, d.Hs[3], d.Ws[3], C, C, &plan));
      SGA_TRY(tc_conv_create(plan, c->b_ga[2], c->ident_b, 2, c->gs_gamma_b[2], &c->tc_ig2));
    }
    SGA_TRY(c->alloc(&c->l3_wscatter, 80 * C)); SGA_TRY(c->alloc(&c->l3_wgather, 80 * C));
    SGA_TRY(l3_pack_weights(c->gs[3].w, C, c->l3_wscatter, c->l3_wgather, st));
    // fused: the IGDN backward recovers x_2 from the saved act_2 and s_2 (the fused forward does not store x)
    SGA_TRY(l3_tc_create(c->b_act[2], c->l3_wscatter, c->l3_wgather, (c->fused && l3_fusable(C)) ? c->gs_gamma_b[2] : nullptr,
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
  SGA_CHECK(grad_y != nullptr && grad_z != nullptr, "null gradient output");
  const StepOut o{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, grad_y, grad_z};
  return grad_common(c, x, y, z, T, step, ny, nz, stats, static_cast<cudaStream_t>(stream), o);
}

int sga_step(sga_ctx* c, const float* x, float* y, float* z, float* m_y, float* v_y, float* m_z, float* v_z, float T,
             int32_t step, const float* ny, const float* nz, float* stats, void* stream) {
  SGA_CHECK(y && z && m_y && v_y && m_z && v_z, "null state tensor");
  const StepOut o{y, z, m_y, v_y, m_z, v_z, nullptr, nullptr};
  return grad_common(c, x, y, z, T, step, ny, nz, stats, static_cast<cudaStream_t>(stream), o);
}

// n_steps iterations of sga.py:210-215 in one call: the two-stream step is enqueued n_steps times from this C++ loop (no
// Python, no host synchronisation per iteration).  With SGA_GRAPH=1 the step is instead captured once into a CUDA graph
// (both streams, programmatic dependent launches included) and replayed; temperature / noise step / Adam lr_t of
// iteration k then come from a device table indexed by a device counter that the step's last kernel advances.
int sga_loop(sga_ctx* c, const float* x, float* y, float* z, float* m_y, float* v_y, float* m_z, float* v_z,
             const float* temperatures, int32_t first_step, int32_t n_steps, float* stats, void* stream) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_SGA, "bits-back contexts are driven through sga_bb_step");
  SGA_CHECK(x && y && z && m_y && v_y && m_z && v_z && temperatures, "null argument");
  SGA_CHECK(n_steps >= 0 && first_step >= 0, "bad step range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const sga_config& g = c->cfg;
  for (int k = 0; k < n_steps; ++k) SGA_CHECK(temperatures[k] > 0.0f, "temperature must be positive");
  const StepOut o{y, z, m_y, v_y, m_z, v_z, nullptr, nullptr};
  // Graph replay is opt-in (SGA_GRAPH=1).  Measured on the B200 (profiles/r2_schedule_ab.md): the eager two-stream issue is
  // never launch-bound (the host enqueues a step in ~0.1 ms), programmatic dependent launch overlaps consecutive steps,
  // while a replayed graph puts a full dependency between steps: 24 images 4.08 ms eager vs 4.18 ms replayed, 3 images
  // 0.566 vs 0.562 ms.
  const char* eg = getenv("SGA_GRAPH");
  const bool graphed = c->cfg.operand == SGA_OPERAND_BF16 && !c->prof && eg != nullptr && eg[0] == '1';
  const char* ep = getenv("SGA_PRIO");
  const bool own_main = (ep != nullptr && ep[0] == '1' && st != c->main && !c->prof && c->cfg.operand == SGA_OPERAND_BF16);
  if (own_main || (graphed && (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread))) {
    // the default streams cannot be captured: run on the context's own stream, ordered after / before the caller's
    SGA_CUDA(cudaEventRecord(c->ev_in, st));
    SGA_CUDA(cudaStreamWaitEvent(c->main, c->ev_in, 0));
    const int rc = sga_loop(c, x, y, z, m_y, v_y, m_z, v_z, temperatures, first_step, n_steps, stats, c->main);
    SGA_CUDA(cudaEventRecord(c->ev_out, c->main));
    SGA_CUDA(cudaStreamWaitEvent(st, c->ev_out, 0));
    return rc;
  }
  if (!graphed) {
    // fp32 contexts (cross-check arithmetic) and per-launch profiling: plain launches
    for (int k = 0; k < n_steps; ++k)
      SGA_TRY(grad_common(c, x, y, z, temperatures[k], first_step + k, nullptr, nullptr, (k + 1 == n_steps) ? stats : nullptr, st, o));
    return SGA_OK;
  }
  // the captured step (keyed by the tensors it is bound to)
  const sga_ctx::GraphKey key{x, y, z, m_y, v_y, m_z, v_z};
  sga_ctx::GraphRec* rec = nullptr;
  for (auto& r : c->graphs)
    if (memcmp(&r.key, &key, sizeof(key)) == 0) rec = &r;
  if (rec == nullptr && n_steps > 0) {
    if (c->graphs.size() >= 8) { cudaGraphExecDestroy(c->graphs.front().exec); c->graphs.erase(c->graphs.begin()); }
    cudaGraph_t graph = nullptr;
    const long long l0 = g_launches;
    SGA_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = step_bf16(c, x, y, z, 1.0f, 0, nullptr, nullptr, c->stats_dev, st, o, c->sp_table, c->sp_idx, c->sp_idx);
    const cudaError_t ec = cudaStreamEndCapture(st, &graph);
    const long long kernels = g_launches - l0;
    g_launches = l0;
    if (rc != SGA_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ec != cudaSuccess) { set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(ec)); return SGA_ERR_CUDA; }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ei)); return SGA_ERR_CUDA; }
    c->graphs.push_back(sga_ctx::GraphRec{key, exec, kernels});
    rec = &c->graphs.back();
  }
  for (int done = 0; done < n_steps;) {
    const int n = (n_steps - done < sga_ctx::kSpCap) ? n_steps - done : sga_ctx::kSpCap;
    const int h = c->sp_half;
    c->sp_half ^= 1;
    SGA_CUDA(cudaEventSynchronize(c->ev_sp[h]));          // the previous upload from this half of the staging buffer is done
    StepParams* hp = c->sp_host + h * sga_ctx::kSpCap;
    for (int k = 0; k < n; ++k) {
      const int it = first_step + done + k;
      hp[k] = StepParams{temperatures[done + k], static_cast<float>(adam_lr_t(g.lr, g.beta1, g.beta2, it)), it, 0};
    }
    SGA_CUDA(cudaMemcpyAsync(c->sp_table, hp, n * sizeof(StepParams), cudaMemcpyHostToDevice, st));
    SGA_CUDA(cudaEventRecord(c->ev_sp[h], st));
    SGA_CUDA(cudaMemsetAsync(c->sp_idx, 0, sizeof(int), st));
    for (int k = 0; k < n; ++k) SGA_CUDA(cudaGraphLaunch(rec->exec, st));
    g_launches += rec->kernels * n;
    done += n;
  }
  if (stats != nullptr && n_steps > 0)
    SGA_CUDA(cudaMemcpyAsync(stats, c->stats_dev, (SGA_STATS_HEADER + SGA_STATS_PER_IMAGE * c->d.N) * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
  return SGA_OK;
}

int sga_eval_hard(sga_ctx* c, const float* x, const float* y, const float* z, float* stats, void* stream) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_SGA, "bits-back contexts are evaluated through sga_bb_eval");
  return forward_f32(c, x, y, z, 1.0f, 0, nullptr, nullptr, 1, false, stats, static_cast<cudaStream_t>(stream));
}

/* Everything downstream of given (y_tilde, z_tilde), fp32: the eval tensors of mbt2018.py:171-185 for the baseline's
 * mean-centred quantisation, and -- with x = NULL -- the decoder's synthesis (mbt2018.py:275-283). */
int sga_eval_tilde(sga_ctx* c, const float* x, const float* y_tilde, const float* z_tilde, float* stats, void* stream) {
  SGA_CHECK(c && c->bound, "weights are not bound (call sga_bind_weights first)");
  SGA_CHECK(c->cfg.variant == SGA_VARIANT_SGA, "bits-back contexts are evaluated through sga_bb_eval");
  SGA_CHECK(y_tilde && z_tilde, "null latent");
  return forward_f32(c, x, y_tilde, z_tilde, 1.0f, 0, nullptr, nullptr, 2, false, x ? stats : nullptr,
                     static_cast<cudaStream_t>(stream));
}

/* ---- entropy coding (codec.cu) ------------------------------------------------------------------------------------ */
int sga_codec_create(sga_ctx* c, sga_codec** out, void* stream) {
  SGA_CHECK(c && c->bound && out, "bad argument (weights must be bound: the z tables come from the factorised prior)");
  sga_codec* cd = codec_new();
  const int rc = codec_build_tables(c->eb_pack, c->d.C, cd, static_cast<cudaStream_t>(stream));
  if (rc != SGA_OK) { sga_codec_destroy(cd); return rc; }
  *out = cd;
  return SGA_OK;
}

/* Symbol binning.  z holds the hyper-latents (rounded here: z_hat = rint(z)); mu / sigma = h_s(z_hat) are evaluated with the
 * fp32 kernels exactly as the decoder will evaluate them.  y (may be NULL on the decoder side) holds the latents;
 * mode 0: y_hat = rint(y) (SGA result, sga.py:240-241), mode 1: y_hat = rint(y - mu) + mu (mbt2018.py baseline).
 * Device outputs, one entry per latent element: tab_y / sym_y / base_y (y_hat = base_y + sym_y) / yhat (optional) and
 * tab_z / sym_z (optional; z_hat = centre[channel] + sym_z). */
int sga_codec_bin(sga_ctx* c, const sga_codec* cd, const float* y, const float* z, int32_t mode, int32_t* tab_y,
                  int32_t* sym_y, float* base_y, float* yhat, int32_t* tab_z, int32_t* sym_z, void* stream) {
  SGA_CHECK(c && c->bound && cd && z && tab_y && base_y, "bad argument");
  SGA_CHECK(codec_num_channels(cd) == 0 || codec_num_channels(cd) == c->d.C, "codec was built for %d channels", codec_num_channels(cd));
  SGA_CHECK(mode == 0 || mode == 1, "unknown quantisation mode %d", mode);
  SGA_CHECK(y == nullptr || sym_y != nullptr, "sym_y is required when y is given");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Dims& d = c->d;
  const sga_config& g = c->cfg;
  SGA_CUDA(cudaMemsetAsync(c->acc, 0, 5 * d.N * sizeof(double), st));
  SGA_TRY(launch_z_path<float>(z, nullptr, d.nz, d.nz / d.N, d.C, c->eb_pack, 1.0f, g.clip_eps, g.seed, 0, 1, 0.0f, c->f_zt,
                               nullptr, nullptr, c->acc + d.N, 0, st));
  SGA_TRY(conv_f32(kUp2, c->hs[0], d.N, d.Hz, d.Wz, c->f_zt, true, 1, nullptr, c->f_h0, st));
  SGA_TRY(conv_f32(kUp2, c->hs[1], d.N, 2 * d.Hz, 2 * d.Wz, c->f_h0, true, 1, nullptr, c->f_h1, st));
  SGA_TRY(conv_f32(kS1True, c->hs[2], d.N, d.Hh, d.Wh, c->f_h1, true, 0, nullptr, c->h2, st));
  SGA_TRY(launch_bin_y(cd, y, c->h2, d.N, d.Hy, d.Wy, d.Hh, d.Wh, d.C, mode, tab_y, sym_y, base_y, yhat, st));
  if (tab_z && sym_z) SGA_TRY(launch_bin_z(z, d.nz, d.C, cd, tab_z, sym_z, st));
  return SGA_OK;
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
