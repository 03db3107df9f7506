// kernels.h -- host-side launchers shared between the translation units of libsga_b200.so
#pragma once
#include "common.cuh"

namespace sga {

constexpr int TENSOR_ID_Y = 1;  // Philox counter word 2 (oracle/sga_oracle.py TENSOR_ID_*)
constexpr int TENSOR_ID_Z = 2;
constexpr int TENSOR_ID_EPS = 3;  // bits-back variant: the Gaussian noise of z_tilde (oracle/bb_oracle.py)

// Per-iteration scalars of the graph-replayed loop (sga_loop): a device table indexed by a device counter that the last
// kernel of every step advances.
struct StepParams { float T; float lr_t; int step; int pad; };

// ---- elementwise.cu
int launch_soft_round(const float* v, const float* inj, long long numel, float T, float eps, uint64_t seed, int step,
                      int tensor_id, float* vt, float* dvt, cudaStream_t st);
int launch_eb_pack(const float* const m[4], const float* const b[4], const float* const f[3], int C, float* pk,
                   cudaStream_t st);
int launch_eb_likelihood(const float* zt, long long numel, int C, const float* pk, float* p, float* dp, cudaStream_t st);
int launch_gauss_likelihood(const float* yt, const float* mu, const float* ls, long long numel, float* p, float* dpdy,
                            float* dpdsb, cudaStream_t st);
template <typename AT>
int launch_z_path(const float* z, const float* inj, long long numel, long long per_image, int C, const float* pk,
                  float T, float eps, uint64_t seed, int step, int hard, float rate_scale, AT* zt_act, float* g_rate,
                  float* dz, double* bits, long long ctr_off, cudaStream_t st, const StepParams* spt = nullptr,
                  const int* spi = nullptr, int split = 0);
// yt_in != nullptr: y_tilde (fp32) was produced by launch_y_soft; only the rate terms are evaluated
template <typename AT>
int launch_y_path(const float* y, const float* inj, const float* h2, int N, int Hy, int Wy, int Hh, int Wh, int C,
                  float T, float eps, uint64_t seed, int step, int hard, float rate_scale, AT* yt_act, float* g_rate,
                  float* dy, AT* g_h2, double* bits, long long ctr_off, cudaStream_t st, const float* yt_in = nullptr);
// soft round of y alone: fp32 y_tilde, d y_tilde / d y and the bf16 operand of g_s layer 0 (split: [pix][hi C | lo C])
int launch_y_soft(const float* y, const float* inj, long long numel, int C, float T, float eps, uint64_t seed, int step,
                  float* yt_f32, __nv_bfloat16* yt_act, float* dy, long long ctr_off, cudaStream_t st,
                  const StepParams* spt = nullptr, const int* spi = nullptr, int split = 0);
// bits-back variant (bb_sga.py): z_tilde = eps*exp(logvar/2) + mean, prior density + slope, log q; see elementwise.cu
template <typename AT>
int launch_zbb_path(const float* z_mean, const float* z_logvar, const float* eps_inj, long long numel, long long per_image,
                    int C, const float* pk, uint64_t seed, int step, float rate_scale, AT* zt_act, float* g_zt, float* dzl,
                    double* bits, double* logq, long long ctr_off, cudaStream_t st);
int launch_bb_grads(const float* g_rate, const float* g_hyper, const float* dzl, long long numel, float half_scale,
                    float* g_mean, float* g_lv, cudaStream_t st);
int launch_split2(const float* in, long long pixels, int C, float* a, float* b, cudaStream_t st);
int launch_finalize_stats_bb(const double* acc, int N, int n_ref, int H, int W, float lmbda, int rate_only, float* out,
                             cudaStream_t st);
int launch_distortion(const float* x, const float* xt, int N, int H, int W, int Ho, int Wo, float grad_scale, float* gx,
                      double* sse, double* sse_r, cudaStream_t st);
int launch_finalize_stats(const double* acc, int N, int n_ref, int H, int W, float lmbda, float* out, cudaStream_t st,
                          int* spi_advance = nullptr);
double adam_lr_t(double lr, double beta1, double beta2, int iterations);
int launch_adam(float* p, float* m, float* v, const float* ga, const float* gb, const float* fac, long long numel,
                double lr, double beta1, double beta2, double eps, int iterations, float* g_out, cudaStream_t st,
                const StepParams* spt = nullptr, const int* spi = nullptr);
int launch_square(const float* x, long long n, float* o, cudaStream_t st);
int launch_gdn_finish(const float* x, const float* nrm, long long n, int inverse, float* out, float* s_out,
                      cudaStream_t st);
int launch_igdn_bwd_pre(const float* g, const float* x, const float* s, long long n, float* t, cudaStream_t st);
int launch_igdn_bwd_finish(const float* g, const float* x, const float* s, const float* u, long long n, float* gx,
                           cudaStream_t st);
int launch_f32_to_bf16(const float* in, long long n, __nv_bfloat16* out, cudaStream_t st);
int launch_bf16_to_f32(const __nv_bfloat16* in, long long n, float* out, cudaStream_t st);

// ---- codec.cu : quantised CDF tables, symbol binning, host range coder (reference mbt2018.py:84-85, 211-221, 248-295)
}  // namespace sga
struct sga_codec;
namespace sga {
int codec_build_tables(const float* eb_pack, int C, sga_codec* cd, cudaStream_t st);
sga_codec* codec_new();
// y may be nullptr (decoder side: only the table index and the base of every element are produced)
int launch_bin_y(const sga_codec* cd, const float* y, const float* h2, int N, int Hy, int Wy, int Hh, int Wh, int C, int mode,
                 int* tab, int* sym, float* base, float* yhat, cudaStream_t st);
int launch_bin_z(const float* z, long long numel, int C, const sga_codec* cd, int* tab, int* sym, cudaStream_t st);
int codec_num_channels(const sga_codec* cd);

// ---- conv_plan.cu : tap tables shared by the SIMT and tcgen05 paths
struct ConvGeom {
  int kind, kh, kw;
  int N, Hin, Win, Cin;    // input tensor [N,Hin,Win,Cin]
  int Hout, Wout, Cout;    // output tensor
};
// A job = one dense multi-tap accumulation into one output view.
struct ConvJobPlan {
  int nviews;
  int view_py[kMaxViews], view_px[kMaxViews];  // phase of each input view (stride-2 sub-grids) or 0,0
  int in_phase_stride;                          // 1 (plain) or 2 (phase views of the input)
  int out_py, out_px, out_phase_stride;         // output view: plain (stride 1) or one phase of the 2x grid
  int ntaps;
  Tap taps[kMaxTaps];
  int Ho, Wo;                                   // pixel grid of this job's output view
};
struct ConvPlan {
  ConvGeom g;
  int njobs;
  ConvJobPlan job[4];
  bool w_transposed;  // B operand is kernel[tap]^T (dgrad kinds)
};
int make_conv_plan(int kind, int kh, int kw, int N, int Hin, int Win, int Cin, int Cout_kernel, ConvPlan* plan);

// ---- conv_simt.cu : fp32 implicit GEMM on CUDA cores (init, final eval, cross-checks)
// w: [taps][K][Nn] with K = contraction channels, Nn = output channels of the op (already transposed for dgrads)
int launch_conv_simt(const ConvPlan& plan, const float* x, const float* w, const float* bias, const float* relu_mask,
                     int relu, float* out, cudaStream_t st);
// w_ref (kh,kw,Cin,Cout) -> [tap][Cout][Cin]
int launch_transpose_taps(const float* w_ref, int taps, int Cin, int Cout, float* w_t, cudaStream_t st);

// ---- conv_tc.cu : bf16 tcgen05 implicit GEMM
struct TcConv;  // opaque: tensor maps + tap tables for one layer at one shape
// fuse: 0 = plain epilogue, 1 = + IGDN forward, 2 = + IGDN backward (mix_op = K-major bf16 [C][C] mixing operand:
// gamma^T for forward, gamma for backward); fusion needs Cout in {64,128,192}
bool tc_conv_can_fuse(int cout);
int tc_conv_create(const ConvPlan& plan, const __nv_bfloat16* x, const __nv_bfloat16* w_kmajor, int fuse,
                   const __nv_bfloat16* mix_op, TcConv** out);
void tc_conv_destroy(TcConv* c);
// epilogue: v = acc + bias[n]; relu; *= (mask > 0); store as bf16 (out_bf16) and/or fp32 (out_f32)
int tc_conv_launch(const TcConv* c, const float* bias, int relu, const __nv_bfloat16* relu_mask, __nv_bfloat16* out_bf16,
                   float* out_f32, cudaStream_t st);
// x = acc + bias; s = sqrt(beta + x^2 @ gamma); act = x*s  -- s and act stored as bf16 (x = act / s in the backward)
int tc_conv_launch_igdn_fwd(TcConv* c, const float* bias, const float* beta, __nv_bfloat16* s_out, __nv_bfloat16* act_out,
                            cudaStream_t st);
// g = acc; x = act/s; g_x = g*s + x*((g*x/s) @ gamma^T) stored as bf16
int tc_conv_launch_igdn_bwd(const TcConv* c, const __nv_bfloat16* act_saved, const __nv_bfloat16* s_saved,
                            __nv_bfloat16* gx_out, cudaStream_t st);
// w_ref (kh,kw,Cin,Cout) fp32 -> K-major bf16 B operand: [tap][Nn][K] (K = Cin for forward kinds, Cout for dgrads)
// dup = 2 (forward kinds only): K = 2*Cin, the weights stacked twice along K for a hi|lo split input operand
int launch_pack_weights_bf16(const float* w_ref, int taps, int Cin, int Cout, bool transposed, __nv_bfloat16* out,
                             cudaStream_t st, int dup = 1);
int tc_selftest_available();

}  // namespace sga
