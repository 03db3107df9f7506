// elementwise.cu -- the HBM-bound kernels of the SGA iteration: Gumbel soft rounding fused with the two
// entropy-model rate terms, the distortion/PSNR reduction, the IGDN point-wise pieces of the fp32 path and
// the Adam update.  Coalesced (pairs of adjacent channels per thread, one Philox call per pair), warp-shuffle
// + one atomicAdd(double) per warp for the per-image reductions.
//
// Reference arithmetic restated here (file:line into /root/reference):
//   soft round        sga.py:88-98, 111-121 (+ tfp RelaxedOneHotCategorical.sample)
//   factorised prior  tfc EntropyBottleneck._likelihood, sga.py:100-104; learned_prior.py:78-121, 296-347
//   Gaussian box      tfc GaussianConditional._likelihood, sga.py:129-136; utils.py:80-102
//   lower_bound grad  math_ops.py:63-76
//   bpp / mse / psnr  sga.py:143-154, 167-174
//   Adam              adam.py:36-57
#include "common.cuh"
#include "philox.cuh"
#include "kernels.h"
#include <type_traits>
#include <stdlib.h>

namespace sga {

// Programmatic dependent launch (see tc_common.cuh): the loop's elementwise kernels let the tensor kernel behind them
// start its prologue early, and wait for the kernel in front of them before touching memory.
__device__ __forceinline__ void ew_pdl_begin() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();   // conv_tc.cu

template <typename... KArgs, typename... Args>
static int launch_ew(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool ew_pdl = pdl_enabled() && getenv("SGA_NO_EW_PDL") == nullptr;
  cfg.numAttrs = ew_pdl ? 1 : 0;
  SGA_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  ++g_launches;
  return SGA_OK;
}

// ------------------------------------------------------------------------------------------------
// soft round of one element. Returns v_tilde, writes d v_tilde / d v.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float soft_round_one(float v, float u0, float u1, float T, float eps, float* dvt) {
  const float f = floorf(v), c = ceilf(v);
  const float lo = -1.0f + eps, hi = 1.0f - eps;
  const float d = v - f, e = c - v;
  const float dc = fminf(fmaxf(d, lo), hi), ec = fminf(fmaxf(e, lo), hi);
  const float l0 = -atanhf(dc) / T, l1 = -atanhf(ec) / T;
  const float g0 = -logf(-logf(u0)), g1 = -logf(-logf(u1));
  const float a = ((l1 + g1) - (l0 + g0)) / T;
  const float s1 = 1.0f / (1.0f + expf(-a));
  const float s0 = 1.0f - s1;
  const float p0 = (d >= lo && d <= hi) ? 1.0f : 0.0f;
  const float p1 = (e >= lo && e <= hi) ? 1.0f : 0.0f;
  const float dl0 = -p0 / ((1.0f - dc * dc) * T);
  const float dl1 = p1 / ((1.0f - ec * ec) * T);
  *dvt = (c - f) * s1 * s0 * (dl1 - dl0) / T;
  return f * s0 + c * s1;
}

// Same function for the bf16 loop (operands are rounded to 8 mantissa bits right after): the two atanh and the two
// Gumbel terms collapse into two logarithms of ratios,
//   l1 - l0 = (atanh(d) - atanh(e)) / T = log((1+d)(1-e) / ((1-d)(1+e))) / (2T),   g1 - g0 = log(log(u0) / log(u1)),
// evaluated with the SFU intrinsics (MUFU.LG2 / EX2 / RCP) where their absolute error is harmless.
__device__ __forceinline__ float soft_round_fast(float v, float u0, float u1, float T, float eps, float* dvt) {
  const float f = floorf(v), c = ceilf(v);
  const float hi = 1.0f - eps;
  const float d = v - f, e = c - v;                       // both in [0, 1]: only the upper clip can bind
  const float dc = fminf(d, hi), ec = fminf(e, hi);
  const float invT = __frcp_rn(T);
  const float ld = __logf(__fdividef((1.0f + dc) * (1.0f - ec), (1.0f - dc) * (1.0f + ec)));
  // -log(u) keeps the accurate logf: near u = 1 the SFU logarithm's absolute error exceeds the value.  The uniform
  // ((bits >> 8) + 0.5) * 2^-24 rounds to exactly 1.0f once in 2^24 draws (about once per iteration at the headline
  // shape): the reference form then has G = -log(-log 1) = +inf and a hard 0/1 sample; the floor keeps the ratio finite
  // and saturates the sigmoid the same way.
  const float n0 = fmaxf(-logf(u0), 1e-37f), n1 = fmaxf(-logf(u1), 1e-37f);
  const float lg = __logf(__fdividef(n0, n1));
  const float a = (0.5f * invT * ld + lg) * invT;
  const float s1 = __frcp_rn(1.0f + __expf(-a));
  const float s0 = 1.0f - s1;
  const float p0 = (d <= hi) ? 1.0f : 0.0f, p1 = (e <= hi) ? 1.0f : 0.0f;
  const float dl = p1 * __frcp_rn(1.0f - ec * ec) + p0 * __frcp_rn(1.0f - dc * dc);     // T * (dl1 - dl0)
  *dvt = (c - f) * s1 * s0 * dl * invT * invT;
  return f * s0 + c * s1;
}
// Graph-replayed loop (sga_loop): temperature, noise step and Adam's lr_t come from a device table indexed by a device
// counter, so that one captured step can be replayed for every iteration without the host touching it.
__device__ __forceinline__ void step_override(const StepParams* __restrict__ spt, const int* __restrict__ spi, float& T,
                                              int& step) {
  if (spt != nullptr) {
    const StepParams sp = spt[*spi];
    T = sp.T;
    step = sp.step;
  }
}

// operand store of a soft-rounded pair (c, c+1) of pixel `pix`: plain [pix][C], or -- split -- [pix][2C] with the bf16
// value in [0,C) and the bf16 of its rounding residual in [C,2C); the first layer then runs with K = 2C against the
// weights stacked twice, so that it sees v_tilde to ~16 mantissa bits (|y| >= 64 would otherwise lose the fraction
// that SGA optimises: bf16 spacing 0.5)
__device__ __forceinline__ void store_operand_pair(float* base, long long pix, int C, int c, float a, float b, int split) {
  *reinterpret_cast<float2*>(base + pix * C + c) = make_float2(a, b);
}
__device__ __forceinline__ void store_operand_pair(__nv_bfloat16* base, long long pix, int C, int c, float a, float b,
                                                   int split) {
  const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
  if (!split) {
    *reinterpret_cast<__nv_bfloat162*>(base + pix * C + c) = hi;
  } else {
    *reinterpret_cast<__nv_bfloat162*>(base + pix * 2 * C + c) = hi;
    *reinterpret_cast<__nv_bfloat162*>(base + pix * 2 * C + C + c) =
        __floats2bfloat162_rn(a - __low2float(hi), b - __high2float(hi));
  }
}

template <bool FAST>
__device__ __forceinline__ float soft_round_sel(float v, float u0, float u1, float T, float eps, float* dvt) {
  return FAST ? soft_round_fast(v, u0, u1, T, eps, dvt) : soft_round_one(v, u0, u1, T, eps, dvt);
}

__device__ __forceinline__ void pair_uniforms(const float* __restrict__ inj, long long numel, long long e0,
                                              uint64_t seed, int step, int tensor_id, float u[4],
                                              long long ctr_off = 0) {
  if (inj != nullptr) {  // injected: [2][numel], DOWN logits first
    u[0] = inj[e0];
    u[1] = inj[numel + e0];
    const bool has2 = (e0 + 1 < numel);
    u[2] = has2 ? inj[e0 + 1] : 0.5f;
    u[3] = has2 ? inj[numel + e0 + 1] : 0.5f;
  } else {
    const Philox4 r = philox4x32_10(static_cast<uint32_t>((e0 + ctr_off) >> 1), static_cast<uint32_t>(step),
                                    static_cast<uint32_t>(tensor_id), 0u, static_cast<uint32_t>(seed),
                                    static_cast<uint32_t>(seed >> 32));
    u[0] = bits_to_uniform(r.x);
    u[1] = bits_to_uniform(r.y);
    u[2] = bits_to_uniform(r.z);
    u[3] = bits_to_uniform(r.w);
  }
}

__global__ void k_soft_round(const float* __restrict__ v, const float* __restrict__ inj, long long numel, float T,
                             float eps, uint64_t seed, int step, int tensor_id, float* __restrict__ vt,
                             float* __restrict__ dvt) {
  const long long pair = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long e0 = pair * 2;
  if (e0 >= numel) return;
  float u[4];
  pair_uniforms(inj, numel, e0, seed, step, tensor_id, u);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (e0 + k < numel) {
      float d;
      const float r = soft_round_one(v[e0 + k], u[2 * k], u[2 * k + 1], T, eps, &d);
      vt[e0 + k] = r;
      if (dvt) dvt[e0 + k] = d;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// factorised prior: per-channel 1->3->3->3->1 network evaluated in forward mode (value and slope).
// Packed parameters pk[j*C + c], j in [0,44): softplus(matrix) and tanh(factor) are precomputed at bind time.
//   0..2 M0 | 3..5 B0 | 6..8 F0 | 9..17 M1 (r*3+k) | 18..20 B1 | 21..23 F1 | 24..32 M2 | 33..35 B2 | 36..38 F2
//   39..41 M3 | 42 B3 | 43 pad
// ------------------------------------------------------------------------------------------------
struct EbParams { float p[44]; };

__device__ __forceinline__ void eb_load(const float* __restrict__ pk, int C, int c, EbParams& P) {
#pragma unroll
  for (int j = 0; j < 43; ++j) P.p[j] = __ldg(pk + j * C + c);
}

__device__ __forceinline__ void eb_logit(const EbParams& P, float u, float* L, float* dL) {
  float h[3], d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    h[r] = P.p[r] * u + P.p[3 + r];
    d[r] = P.p[r];
    const float t = tanhf(h[r]);
    h[r] += P.p[6 + r] * t;
    d[r] *= 1.0f + P.p[6 + r] * (1.0f - t * t);
  }
#pragma unroll
  for (int layer = 0; layer < 2; ++layer) {
    const int mo = 9 + 15 * layer, bo = 18 + 15 * layer, fo = 21 + 15 * layer;
    float h2[3], d2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float a = P.p[bo + r], g = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a += P.p[mo + r * 3 + k] * h[k];
        g += P.p[mo + r * 3 + k] * d[k];
      }
      const float t = tanhf(a);
      h2[r] = a + P.p[fo + r] * t;
      d2[r] = g * (1.0f + P.p[fo + r] * (1.0f - t * t));
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { h[r] = h2[r]; d[r] = d2[r]; }
  }
  float a = P.p[42], g = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) { a += P.p[39 + k] * h[k]; g += P.p[39 + k] * d[k]; }
  *L = a;
  *dL = g;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float signf_(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

// p = |sigmoid(s*upper) - sigmoid(s*lower)|, s = -sign(lower+upper) (stop-gradient)
__device__ __forceinline__ float eb_likelihood_one(const EbParams& P, float zt, float* dp) {
  float lower, dlower, upper, dupper;
  eb_logit(P, zt - 0.5f, &lower, &dlower);
  eb_logit(P, zt + 0.5f, &upper, &dupper);
  const float s = -signf_(lower + upper);
  const float su = sigmoidf_(s * upper), sl = sigmoidf_(s * lower);
  const float diff = su - sl;
  *dp = signf_(diff) * s * (su * (1.0f - su) * dupper - sl * (1.0f - sl) * dlower);
  return fabsf(diff);
}

__global__ void k_eb_likelihood(const float* __restrict__ zt, long long numel, int C, const float* __restrict__ pk,
                                float* __restrict__ p_out, float* __restrict__ dp_out) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= numel) return;
  EbParams P;
  eb_load(pk, C, static_cast<int>(e % C), P);
  float dp;
  const float p = eb_likelihood_one(P, zt[e], &dp);
  if (p_out) p_out[e] = p;
  if (dp_out) dp_out[e] = dp;
}

__global__ void k_eb_pack(const float* m0, const float* m1, const float* m2, const float* m3, const float* b0,
                          const float* b1, const float* b2, const float* b3, const float* f0, const float* f1,
                          const float* f2, int C, float* __restrict__ pk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  auto sp = [](float x) { return (x > 20.0f) ? x : log1pf(expf(x)); };
  for (int r = 0; r < 3; ++r) {
    pk[(0 + r) * C + c] = sp(m0[c * 3 + r]);
    pk[(3 + r) * C + c] = b0[c * 3 + r];
    pk[(6 + r) * C + c] = tanhf(f0[c * 3 + r]);
    pk[(18 + r) * C + c] = b1[c * 3 + r];
    pk[(21 + r) * C + c] = tanhf(f1[c * 3 + r]);
    pk[(33 + r) * C + c] = b2[c * 3 + r];
    pk[(36 + r) * C + c] = tanhf(f2[c * 3 + r]);
    pk[(39 + r) * C + c] = sp(m3[c * 3 + r]);
    for (int k = 0; k < 3; ++k) {
      pk[(9 + r * 3 + k) * C + c] = sp(m1[c * 9 + r * 3 + k]);
      pk[(24 + r * 3 + k) * C + c] = sp(m2[c * 9 + r * 3 + k]);
    }
  }
  pk[42 * C + c] = b3[c];
  pk[43 * C + c] = 0.0f;
}

// ------------------------------------------------------------------------------------------------
// K1(z)+K2: soft round of z fused with the factorised-prior rate.
//   zt_act : z_tilde in the activation dtype (input of h_s)       g_rate: d(rate)/d z_tilde (fp32)
//   dz     : d z_tilde / d z                                      bits[n] += -ln p  (double)
// hard != 0: z_tilde = rint(z) (sga.py:240-241), no noise.
// ------------------------------------------------------------------------------------------------
template <typename AT, bool FAST = std::is_same<AT, __nv_bfloat16>::value>
__global__ void k_z_path(const float* __restrict__ z, const float* __restrict__ inj, long long numel,
                         long long per_image, int C, const float* __restrict__ pk, float T, float eps, uint64_t seed,
                         int step, int hard, float rate_scale, AT* __restrict__ zt_act, float* __restrict__ g_rate,
                         float* __restrict__ dz, double* __restrict__ bits, long long ctr_off,
                         const StepParams* __restrict__ spt, const int* __restrict__ spi, int split) {
  ew_pdl_begin();
  step_override(spt, spi, T, step);
  const long long pair = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long e0 = pair * 2;
  float nll = 0.0f;
  long long img = -1;
  if (e0 < numel) {
    img = e0 / per_image;
    float u[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    if (!hard) pair_uniforms(inj, numel, e0, seed, step, TENSOR_ID_Z, u, ctr_off);
    float ztk[2] = {0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long e = e0 + k;
      if (e < numel) {
        float d = 0.0f, zt;
        if (hard) zt = (hard == 2) ? z[e] : rintf(z[e]);      // 2: z already holds z_tilde
        else zt = soft_round_sel<FAST>(z[e], u[2 * k], u[2 * k + 1], T, eps, &d);
        EbParams P;
        eb_load(pk, C, static_cast<int>(e % C), P);
        float dp;
        const float p = eb_likelihood_one(P, zt, &dp);
        const float pc = fmaxf(p, 1e-9f);
        nll += -logf(pc);
        ztk[k] = zt;
        if (g_rate) g_rate[e] = (-rate_scale / pc) * dp;
        if (dz) dz[e] = d;
      }
    }
    // C is even and e0 is even: the pair (e0, e0 + 1) lies in one pixel
    if (e0 + 1 < numel) store_operand_pair(zt_act, e0 / C, C, static_cast<int>(e0 % C), ztk[0], ztk[1], split);
    else zt_act[e0] = from_f32<AT>(ztk[0]);
  }
  // per-image reduction: a warp may straddle an image boundary only if per_image is not a multiple of 64
  const unsigned full = 0xffffffffu;
  const long long img0 = __shfl_sync(full, img, 0);
  const bool uniform = __all_sync(full, img == img0);
  if (uniform) {
    const float s = warp_sum(nll);
    if ((threadIdx.x & 31) == 0 && img0 >= 0) atomicAdd(bits + img0, static_cast<double>(s));
  } else if (img >= 0) {
    atomicAdd(bits + img, static_cast<double>(nll));
  }
}

// ------------------------------------------------------------------------------------------------
// bits-back variant (bb_sga.py:99-111, 132-141): z_tilde = eps*exp(logvar/2) + mean, rate under the DENSITY of the
// flexible prior (learned_prior.py:296-347: product of the layer Jacobians times sigmoid'), bits-back term log q.
// The density's slope needs the second derivative of the CDF network: forward mode, second order.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void eb_gate2(float f, float& a, float& a1, float& a2) {
  // y = a + f tanh(a):  y' = a'(1 + f(1-t^2)),  y'' = a''(1 + f(1-t^2)) - 2 f t (1-t^2) a'^2
  const float t = tanhf(a);
  const float q = 1.0f - t * t;
  const float gate = 1.0f + f * q;
  a2 = a2 * gate - 2.0f * f * t * q * a1 * a1;
  a1 = a1 * gate;
  a = a + f * t;
}
__device__ __forceinline__ void eb_logit2(const EbParams& P, float u, float* L, float* L1, float* L2) {
  float h[3], d[3], dd[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    h[r] = P.p[r] * u + P.p[3 + r];
    d[r] = P.p[r];
    dd[r] = 0.0f;
    eb_gate2(P.p[6 + r], h[r], d[r], dd[r]);
  }
#pragma unroll
  for (int layer = 0; layer < 2; ++layer) {
    const int mo = 9 + 15 * layer, bo = 18 + 15 * layer, fo = 21 + 15 * layer;
    float h2[3], d2[3], dd2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float a = P.p[bo + r], g = 0.0f, gg = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a += P.p[mo + r * 3 + k] * h[k];
        g += P.p[mo + r * 3 + k] * d[k];
        gg += P.p[mo + r * 3 + k] * dd[k];
      }
      eb_gate2(P.p[fo + r], a, g, gg);
      h2[r] = a; d2[r] = g; dd2[r] = gg;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { h[r] = h2[r]; d[r] = d2[r]; dd[r] = dd2[r]; }
  }
  float a = P.p[42], g = 0.0f, gg = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) { a += P.p[39 + k] * h[k]; g += P.p[39 + k] * d[k]; gg += P.p[39 + k] * dd[k]; }
  *L = a; *L1 = g; *L2 = gg;
}

template <typename AT>
__global__ void k_zbb_path(const float* __restrict__ z_mean, const float* __restrict__ z_logvar,
                           const float* __restrict__ eps_inj, long long numel, long long per_image, int C,
                           const float* __restrict__ pk, uint64_t seed, int step, float rate_scale, AT* __restrict__ zt_act,
                           float* __restrict__ g_zt, float* __restrict__ dzl, double* __restrict__ bits,
                           double* __restrict__ logq, long long ctr_off) {
  ew_pdl_begin();
  const long long pair = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long e0 = pair * 2;
  float nll = 0.0f, lq = 0.0f;
  long long img = -1;
  if (e0 < numel) {
    img = e0 / per_image;
    float u[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    if (eps_inj == nullptr) pair_uniforms(nullptr, numel, e0, seed, step, TENSOR_ID_EPS, u, ctr_off);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long e = e0 + k;
      if (e < numel) {
        // Box-Muller: eps = sqrt(-2 ln u0) cos(2 pi u1)   (oracle/bb_oracle.py normal_from_uniforms)
        const float eps = eps_inj ? eps_inj[e] : sqrtf(-2.0f * logf(u[2 * k])) * cosf(6.283185307179586f * u[2 * k + 1]);
        const float lv = z_logvar[e];
        const float sd = expf(0.5f * lv);
        const float zt = eps * sd + z_mean[e];
        EbParams P;
        eb_load(pk, C, static_cast<int>(e % C), P);
        float L, L1, L2;
        eb_logit2(P, zt, &L, &L1, &L2);
        const float sg = sigmoidf_(L);
        const float w = sg * (1.0f - sg);
        const float pdf = w * L1;
        const float dpdf = w * ((1.0f - 2.0f * sg) * L1 * L1 + L2);
        const float pc = fmaxf(pdf, 1e-9f);                      // lower_bound: the (negative) upstream gradient always passes
        nll += -logf(pc);
        const float r = zt - z_mean[e];
        lq += -0.5f * (r * r * expf(-lv) + lv + 1.8378770664093453f);   // utils.py:75-77
        zt_act[e] = from_f32<AT>(zt);
        if (g_zt) g_zt[e] = (-rate_scale / pc) * dpdf;
        if (dzl) dzl[e] = 0.5f * eps * sd;                       // d z_tilde / d logvar
      }
    }
  }
  const unsigned full = 0xffffffffu;
  const long long img0 = __shfl_sync(full, img, 0);
  const bool uniform = __all_sync(full, img == img0);
  if (uniform) {
    const float s = warp_sum(nll), q = warp_sum(lq);
    if ((threadIdx.x & 31) == 0 && img0 >= 0) {
      atomicAdd(bits + img0, static_cast<double>(s));
      atomicAdd(logq + img0, static_cast<double>(q));
    }
  } else if (img >= 0) {
    atomicAdd(bits + img, static_cast<double>(nll));
    atomicAdd(logq + img, static_cast<double>(lq));
  }
}

// g_mean = g_zt_rate + g_hyper;  g_logvar = g_mean * dz_tilde/dlogvar - rate_scale/2   (d log q / d logvar = -1/2 enters
// the loss with +rate_scale: eval_bpp = ... - bpp_back, bpp_back = -sum log q / (ln2 HW))
// (no __restrict__: the host passes g_mean = g_rate and g_lv = dzl to update in place)
__global__ void k_bb_grads(const float* g_rate, const float* __restrict__ g_hyper, const float* dzl, long long numel,
                           float half_scale, float* g_mean, float* g_lv) {
  ew_pdl_begin();
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= numel) return;
  const float g = g_rate[e] + g_hyper[e];
  const float d = dzl[e];
  g_mean[e] = g;
  g_lv[e] = g * d - half_scale;
}

// [pixels, 2C] -> two [pixels, C] tensors (z_mean | z_logvar, bb_sga.py:99)
__global__ void k_split2(const float* __restrict__ in, long long pixels, int C, float* __restrict__ a, float* __restrict__ b) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= pixels * C) return;
  const long long p = e / C;
  const int c = static_cast<int>(e % C);
  a[e] = in[p * 2 * C + c];
  b[e] = in[p * 2 * C + C + c];
}

// bits-back statistics: rd_loss, train_mse, train_bpp, then per image psnr, y_bpp, z_bpp, sse, mse_rounded, bpp_back
__global__ void k_finalize_stats_bb(const double* __restrict__ acc, int N, int n_ref, int H, int W, float lmbda,
                                    int rate_only, float* __restrict__ out) {
  ew_pdl_begin();
  // acc: [0,N) y nats, [N,2N) z nats, [2N,3N) sse, [3N,4N) sse rounded, [4N,5N) sum log q
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double npix = static_cast<double>(H) * W;
  const double ln2 = 0.69314718055994530942;
  double bpp_sum = 0.0, sse_sum = 0.0;
  for (int n = 0; n < N; ++n) {
    const double yb = acc[n] / (ln2 * npix), zb = acc[N + n] / (ln2 * npix), back = -acc[4 * N + n] / (ln2 * npix);
    const double mse_r = acc[3 * N + n] / (npix * 3.0);
    out[3 + 6 * n + 0] = rate_only ? 0.0f : static_cast<float>(20.0 * log10(255.0) - 10.0 * log10(mse_r));
    out[3 + 6 * n + 1] = static_cast<float>(yb);
    out[3 + 6 * n + 2] = static_cast<float>(zb);
    out[3 + 6 * n + 3] = static_cast<float>(acc[2 * N + n]);
    out[3 + 6 * n + 4] = static_cast<float>(mse_r);
    out[3 + 6 * n + 5] = static_cast<float>(back);
    bpp_sum += yb + zb - back;
    sse_sum += acc[2 * N + n];
  }
  const double train_mse = sse_sum / (static_cast<double>(n_ref) * npix * 3.0) * 65025.0;
  const double train_bpp = bpp_sum / n_ref;
  out[0] = static_cast<float>((lmbda > 0.0f && !rate_only) ? lmbda * train_mse + train_bpp : train_bpp);
  out[1] = static_cast<float>(train_mse);
  out[2] = static_cast<float>(train_bpp);
}

// ------------------------------------------------------------------------------------------------
// K1(y)+K4: soft round of y fused with the Gaussian-conditional rate.  Iterates over the pixel grid of
// h2 = h_s(z_tilde) [N,Hh,Wh,2C]; pixels outside y's grid (the crop of sga.py:126-128) get zero gradient.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float std_cdf(float x) { return 0.5f * erfcf(-0.70710678118654752440f * x); }

__device__ __forceinline__ float gauss_likelihood_one(float yt, float mu, float logsigma, float* dp_dy, float* dp_dsb,
                                                      float* sigma_out) {
  const float sigma = expf(logsigma);
  const float sb = fmaxf(sigma, 0.11f);
  const float d = yt - mu;
  const float v = fabsf(d);
  const float a = (0.5f - v) / sb, b = (-0.5f - v) / sb;
  const float p = std_cdf(a) - std_cdf(b);
  const float k = 0.39894228040143267794f;
  const float pa = expf(-0.5f * a * a) * k, pb = expf(-0.5f * b * b) * k;
  *dp_dy = -signf_(d) * (pa - pb) / sb;
  *dp_dsb = -(a * pa - b * pb) / sb;
  *sigma_out = sigma;
  return p;
}

// bf16-loop variant: one reciprocal instead of four divisions, SFU exponentials (erfc stays the accurate one: the
// likelihood reaches 1e-9 in the tails and enters the loss through its logarithm)
__device__ __forceinline__ float gauss_likelihood_fast(float yt, float mu, float logsigma, float* dp_dy, float* dp_dsb,
                                                       float* sigma_out) {
  const float sigma = __expf(logsigma);
  const float sb = fmaxf(sigma, 0.11f);
  const float inv = __frcp_rn(sb);
  const float d = yt - mu;
  const float v = fabsf(d);
  const float a = (0.5f - v) * inv, b = (-0.5f - v) * inv;
  const float p = std_cdf(a) - std_cdf(b);
  const float k = 0.39894228040143267794f;
  const float pa = __expf(-0.5f * a * a) * k, pb = __expf(-0.5f * b * b) * k;
  *dp_dy = -signf_(d) * (pa - pb) * inv;
  *dp_dsb = -(a * pa - b * pb) * inv;
  *sigma_out = sigma;
  return p;
}

__global__ void k_gauss_likelihood(const float* __restrict__ yt, const float* __restrict__ mu,
                                   const float* __restrict__ ls, long long numel, float* __restrict__ p_out,
                                   float* __restrict__ dpdy, float* __restrict__ dpdsb) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= numel) return;
  float a, b, s;
  const float p = gauss_likelihood_one(yt[e], mu[e], ls[e], &a, &b, &s);
  if (p_out) p_out[e] = p;
  if (dpdy) dpdy[e] = a;
  if (dpdsb) dpdsb[e] = b;
}

__device__ __forceinline__ void store_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store_pair(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// FAST (the bf16 loop): SFU-intrinsic soft round / Gaussian terms; the fp32 instantiation keeps the reference formulas
template <typename AT, bool FAST = std::is_same<AT, __nv_bfloat16>::value>
__global__ void k_y_path(const float* __restrict__ y, const float* __restrict__ inj, const float* __restrict__ h2,
                         int N, int Hy, int Wy, int Hh, int Wh, int C, float T, float eps, uint64_t seed, int step,
                         int hard, float rate_scale, AT* __restrict__ yt_act, float* __restrict__ g_rate,
                         float* __restrict__ dy, AT* __restrict__ g_h2, double* __restrict__ bits, long long ctr_off,
                         const float* __restrict__ yt_in) {
  ew_pdl_begin();
  const int halfC = C >> 1;
  const long long pair = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long npairs = static_cast<long long>(N) * Hh * Wh * halfC;
  float nll = 0.0f;
  long long img = -1;
  if (pair < npairs) {
    const int cp = static_cast<int>(pair % halfC);
    long long pix = pair / halfC;
    const int j = static_cast<int>(pix % Wh);
    pix /= Wh;
    const int i = static_cast<int>(pix % Hh);
    const int n = static_cast<int>(pix / Hh);
    const long long hpix = (static_cast<long long>(n) * Hh + i) * Wh + j;
    const int c = cp * 2;
    if (i < Hy && j < Wy) {
      img = n;
      const long long numel = static_cast<long long>(N) * Hy * Wy * C;
      const long long e0 = ((static_cast<long long>(n) * Hy + i) * Wy + j) * C + c;
      float u[4] = {0.5f, 0.5f, 0.5f, 0.5f};
      if (!hard && yt_in == nullptr) pair_uniforms(inj, numel, e0, seed, step, TENSOR_ID_Y, u, ctr_off);
      // the pair (c, c+1) is contiguous in every tensor and c is even: 8-byte / 4-byte vector accesses
      // yt_in: y_tilde was produced by k_y_soft (two-stream schedule): only the rate terms are evaluated here
      const float2 yv = *reinterpret_cast<const float2*>((yt_in ? yt_in : y) + e0);
      const float2 muv = *reinterpret_cast<const float2*>(h2 + hpix * 2 * C + c);
      const float2 lsv = *reinterpret_cast<const float2*>(h2 + hpix * 2 * C + C + c);
      const float yk[2] = {yv.x, yv.y}, muk[2] = {muv.x, muv.y}, lsk[2] = {lsv.x, lsv.y};
      float ytk[2], grk[2], dk[2], gmu[2], gls[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float d = 0.0f, yt;
        if (yt_in) yt = yk[k];
        else if (hard) yt = (hard == 2) ? yk[k] : rintf(yk[k]);   // 2: y already holds y_tilde
        else yt = soft_round_sel<FAST>(yk[k], u[2 * k], u[2 * k + 1], T, eps, &d);
        float dpdy, dpdsb, sigma;
        const float p = FAST ? gauss_likelihood_fast(yt, muk[k], lsk[k], &dpdy, &dpdsb, &sigma)
                             : gauss_likelihood_one(yt, muk[k], lsk[k], &dpdy, &dpdsb, &sigma);
        const float pc = fmaxf(p, 1e-9f);
        nll += FAST ? -__logf(pc) : -logf(pc);
        const float gp = FAST ? -rate_scale * __frcp_rn(pc) : -rate_scale / pc;  // upstream of p: always negative -> the bound passes it
        const float gsb = gp * dpdsb;
        // lower_bound(sigma, 0.11): pass iff sigma >= bound or gradient < 0 (math_ops.py:74-76)
        const float gsig = (sigma >= 0.11f || gsb < 0.0f) ? gsb : 0.0f;
        ytk[k] = yt; grk[k] = gp * dpdy; dk[k] = d; gmu[k] = -gp * dpdy; gls[k] = gsig * sigma;
      }
      if (yt_in == nullptr) store_pair(yt_act + e0, ytk[0], ytk[1]);
      if (g_rate) {
        *reinterpret_cast<float2*>(g_rate + e0) = make_float2(grk[0], grk[1]);
        if (yt_in == nullptr) *reinterpret_cast<float2*>(dy + e0) = make_float2(dk[0], dk[1]);
        store_pair(g_h2 + hpix * 2 * C + c, gmu[0], gmu[1]);
        store_pair(g_h2 + hpix * 2 * C + C + c, gls[0], gls[1]);
      }
    } else if (g_h2) {
      store_pair(g_h2 + hpix * 2 * C + c, 0.0f, 0.0f);
      store_pair(g_h2 + hpix * 2 * C + C + c, 0.0f, 0.0f);
    }
  }
  const unsigned full = 0xffffffffu;
  const long long img0 = __shfl_sync(full, img, 0);
  const bool uniform = __all_sync(full, img == img0);
  if (uniform) {
    const float s = warp_sum(nll);
    if ((threadIdx.x & 31) == 0 && img0 >= 0) atomicAdd(bits + img0, static_cast<double>(s));
  } else if (img >= 0) {
    atomicAdd(bits + img, static_cast<double>(nll));
  }
}

// K1(y) alone (two-stream schedule of the bf16 loop): the synthesis chain only needs y_tilde, so the soft round runs
// first on the main stream while z_path -> h_s -> rate (k_y_path with yt_in) run on the side stream.
//   yt_f32 : y_tilde in fp32 for the rate kernel      yt_act : GEMM operand of g_s layer 0 (plain or hi|lo split)
template <bool FAST>
__global__ void k_y_soft(const float* __restrict__ y, const float* __restrict__ inj, long long numel, int C, float T,
                         float eps, uint64_t seed, int step, float* __restrict__ yt_f32, __nv_bfloat16* __restrict__ yt_act,
                         float* __restrict__ dy, long long ctr_off, const StepParams* __restrict__ spt,
                         const int* __restrict__ spi, int split) {
  ew_pdl_begin();
  step_override(spt, spi, T, step);
  const long long pair = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long e0 = pair * 2;
  if (e0 >= numel) return;             // numel = N*Hy*Wy*C is even (C is even)
  float u[4];
  pair_uniforms(inj, numel, e0, seed, step, TENSOR_ID_Y, u, ctr_off);
  const float2 yv = *reinterpret_cast<const float2*>(y + e0);
  float d0, d1;
  const float t0 = soft_round_sel<FAST>(yv.x, u[0], u[1], T, eps, &d0);
  const float t1 = soft_round_sel<FAST>(yv.y, u[2], u[3], T, eps, &d1);
  *reinterpret_cast<float2*>(yt_f32 + e0) = make_float2(t0, t1);
  *reinterpret_cast<float2*>(dy + e0) = make_float2(d0, d1);
  store_operand_pair(yt_act, e0 / C, C, static_cast<int>(e0 % C), t0, t1, split);
}

// ------------------------------------------------------------------------------------------------
// K6: distortion.  xt [N,Ho,Wo,3] (fp32) is the un-cropped output of g_s; x [N,H,W,3].
//   sse[n]   += (x - xt)^2                         (float reconstruction, sga.py:150)
//   sse_r[n] += (255 x - round(clip(xt,0,1) 255))^2 (sga.py:167-174)
//   gx = grad_scale * (xt - x) inside the crop, 0 outside (sga.py:123)
// ------------------------------------------------------------------------------------------------
__global__ void k_distortion(const float* __restrict__ x, const float* __restrict__ xt, int N, int H, int W, int Ho,
                             int Wo, float grad_scale, float* __restrict__ gx, double* __restrict__ sse,
                             double* __restrict__ sse_r) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // over [N,Ho,Wo] pixels
  const long long npix = static_cast<long long>(N) * Ho * Wo;
  float s = 0.0f, sr = 0.0f;
  long long img = -1;
  if (idx < npix) {
    const int j = static_cast<int>(idx % Wo);
    const long long t = idx / Wo;
    const int i = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    img = n;
    const bool inside = (i < H && j < W);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g = 0.0f;
      if (inside) {
        const float xv = x[((static_cast<long long>(n) * H + i) * W + j) * 3 + c];
        const float tv = xt[idx * 3 + c];
        const float d = tv - xv;
        s += d * d;
        const float r = rintf(fminf(fmaxf(tv, 0.0f), 1.0f) * 255.0f) - xv * 255.0f;
        sr += r * r;
        g = grad_scale * d;
      }
      if (gx) gx[idx * 3 + c] = g;
    }
  }
  const unsigned full = 0xffffffffu;
  const long long img0 = __shfl_sync(full, img, 0);
  const bool uniform = __all_sync(full, img == img0);
  if (uniform) {
    const float a = warp_sum(s), b = warp_sum(sr);
    if ((threadIdx.x & 31) == 0 && img0 >= 0) {
      atomicAdd(sse + img0, static_cast<double>(a));
      atomicAdd(sse_r + img0, static_cast<double>(b));
    }
  } else if (img >= 0) {
    atomicAdd(sse + img, static_cast<double>(s));
    atomicAdd(sse_r + img, static_cast<double>(sr));
  }
}

// stats_out (sga_b200.h): rd_loss, train_mse, train_bpp, then per image psnr, y_bpp, z_bpp, sse, mse_rounded
__global__ void k_finalize_stats(const double* __restrict__ acc, int N, int n_ref, int H, int W, float lmbda,
                                 float* __restrict__ out, int* __restrict__ spi_advance) {
  ew_pdl_begin();
  if (spi_advance != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *spi_advance += 1;   // last kernel of a replayed step
  // acc: [0,N) y nats, [N,2N) z nats, [2N,3N) sse, [3N,4N) sse rounded
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double npix = static_cast<double>(H) * W;
  const double ln2 = 0.69314718055994530942;
  double bpp_sum = 0.0, sse_sum = 0.0;
  for (int n = 0; n < N; ++n) {
    const double yb = acc[n] / (ln2 * npix), zb = acc[N + n] / (ln2 * npix);
    const double mse_r = acc[3 * N + n] / (npix * 3.0);
    out[3 + 5 * n + 0] = static_cast<float>(20.0 * log10(255.0) - 10.0 * log10(mse_r));
    out[3 + 5 * n + 1] = static_cast<float>(yb);
    out[3 + 5 * n + 2] = static_cast<float>(zb);
    out[3 + 5 * n + 3] = static_cast<float>(acc[2 * N + n]);
    out[3 + 5 * n + 4] = static_cast<float>(mse_r);
    bpp_sum += yb + zb;
    sse_sum += acc[2 * N + n];
  }
  const double train_mse = sse_sum / (static_cast<double>(n_ref) * npix * 3.0) * 65025.0;
  const double train_bpp = bpp_sum / n_ref;
  out[0] = static_cast<float>(lmbda > 0.0f ? lmbda * train_mse + train_bpp : train_bpp);
  out[1] = static_cast<float>(train_mse);
  out[2] = static_cast<float>(train_bpp);
}

// ------------------------------------------------------------------------------------------------
// K8: Adam (adam.py:36-57).  g = (ga + gb) * fac  (the two gradient contributions of a latent times the soft-round
// derivative); separate roundings (no FMA contraction) to follow numpy's float32 arithmetic.
// ------------------------------------------------------------------------------------------------
__global__ void k_adam(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ ga,
                       const float* __restrict__ gb, const float* __restrict__ fac, long long numel, float lr_t,
                       float b1, float omb1, float b2, float omb2, float eps, float* __restrict__ g_out,
                       const StepParams* __restrict__ spt, const int* __restrict__ spi) {
  ew_pdl_begin();
  if (spt != nullptr) lr_t = spt[*spi].lr_t;
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= numel) return;
  float g = ga[e];
  if (gb) g = __fadd_rn(g, gb[e]);
  if (fac) g = __fmul_rn(g, fac[e]);
  if (g_out) g_out[e] = g;
  if (p == nullptr) return;
  const float mt = __fadd_rn(__fmul_rn(b1, m[e]), __fmul_rn(omb1, g));
  const float vt = __fadd_rn(__fmul_rn(b2, v[e]), __fmul_rn(omb2, __fmul_rn(g, g)));
  const float den = __fadd_rn(__fsqrt_rn(vt), eps);
  p[e] = __fsub_rn(p[e], __fdiv_rn(__fmul_rn(lr_t, mt), den));
  m[e] = mt;
  v[e] = vt;
}

// four elements per thread with 16-byte accesses (same per-element arithmetic and rounding order)
__device__ __forceinline__ float adam_one(float g, float& p, float& m, float& v, float lr_t, float b1, float omb1, float b2,
                                          float omb2, float eps) {
  const float mt = __fadd_rn(__fmul_rn(b1, m), __fmul_rn(omb1, g));
  const float vt = __fadd_rn(__fmul_rn(b2, v), __fmul_rn(omb2, __fmul_rn(g, g)));
  const float den = __fadd_rn(__fsqrt_rn(vt), eps);
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(lr_t, mt), den));
  m = mt;
  v = vt;
  return g;
}
__global__ void k_adam_v4(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ ga,
                          const float* __restrict__ gb, const float* __restrict__ fac, long long numel4, float lr_t, float b1,
                          float omb1, float b2, float omb2, float eps, const StepParams* __restrict__ spt,
                          const int* __restrict__ spi) {
  ew_pdl_begin();
  if (spt != nullptr) lr_t = spt[*spi].lr_t;
  const long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (q >= numel4) return;
  float4 g4 = reinterpret_cast<const float4*>(ga)[q];
  if (gb) {
    const float4 t = reinterpret_cast<const float4*>(gb)[q];
    g4.x = __fadd_rn(g4.x, t.x); g4.y = __fadd_rn(g4.y, t.y); g4.z = __fadd_rn(g4.z, t.z); g4.w = __fadd_rn(g4.w, t.w);
  }
  if (fac) {
    const float4 t = reinterpret_cast<const float4*>(fac)[q];
    g4.x = __fmul_rn(g4.x, t.x); g4.y = __fmul_rn(g4.y, t.y); g4.z = __fmul_rn(g4.z, t.z); g4.w = __fmul_rn(g4.w, t.w);
  }
  float4 p4 = reinterpret_cast<float4*>(p)[q], m4 = reinterpret_cast<float4*>(m)[q], v4 = reinterpret_cast<float4*>(v)[q];
  adam_one(g4.x, p4.x, m4.x, v4.x, lr_t, b1, omb1, b2, omb2, eps);
  adam_one(g4.y, p4.y, m4.y, v4.y, lr_t, b1, omb1, b2, omb2, eps);
  adam_one(g4.z, p4.z, m4.z, v4.z, lr_t, b1, omb1, b2, omb2, eps);
  adam_one(g4.w, p4.w, m4.w, v4.w, lr_t, b1, omb1, b2, omb2, eps);
  reinterpret_cast<float4*>(p)[q] = p4;
  reinterpret_cast<float4*>(m)[q] = m4;
  reinterpret_cast<float4*>(v)[q] = v4;
}

// ------------------------------------------------------------------------------------------------
// point-wise pieces of (I)GDN for the fp32 path
// ------------------------------------------------------------------------------------------------
__global__ void k_square(const float* __restrict__ x, long long n, float* __restrict__ o) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e < n) o[e] = x[e] * x[e];
}
// nrm = beta + x^2 @ gamma (bias already added by the GEMM). inverse: out = x*sqrt(nrm) else x/sqrt(nrm)
__global__ void k_gdn_finish(const float* __restrict__ x, const float* __restrict__ nrm, long long n, int inverse,
                             float* __restrict__ out, float* __restrict__ s_out) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= n) return;
  const float s = sqrtf(nrm[e]);
  out[e] = inverse ? x[e] * s : x[e] / s;
  if (s_out) s_out[e] = s;
}
__global__ void k_igdn_bwd_pre(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ s,
                               long long n, float* __restrict__ t) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e < n) t[e] = g[e] * x[e] / s[e];
}
__global__ void k_igdn_bwd_finish(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ s,
                                  const float* __restrict__ u, long long n, float* __restrict__ gx) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e < n) gx[e] = g[e] * s[e] + x[e] * u[e];
}

template <typename TI, typename TO>
__global__ void k_convert(const TI* __restrict__ in, long long n, TO* __restrict__ out) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e < n) out[e] = from_f32<TO>(to_f32<TI>(in[e]));
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(long long n, int bs) { return static_cast<unsigned>((n + bs - 1) / bs); }
// SGA_EW_EXACT=1: the bf16 loop's elementwise kernels use the reference formulas (expf / logf / atanhf / divisions)
// instead of the SFU-intrinsic variants
static bool ew_exact() { const char* e = getenv("SGA_EW_EXACT"); return e != nullptr && e[0] == '1'; }

int launch_soft_round(const float* v, const float* inj, long long numel, float T, float eps, uint64_t seed, int step,
                      int tensor_id, float* vt, float* dvt, cudaStream_t st) {
  if (numel <= 0) return SGA_OK;
  const long long pairs = (numel + 1) / 2;
  k_soft_round<<<blocks_for(pairs, 256), 256, 0, st>>>(v, inj, numel, T, eps, seed, step, tensor_id, vt, dvt);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

int launch_eb_pack(const float* const m[4], const float* const b[4], const float* const f[3], int C, float* pk,
                   cudaStream_t st) {
  k_eb_pack<<<blocks_for(C, 128), 128, 0, st>>>(m[0], m[1], m[2], m[3], b[0], b[1], b[2], b[3], f[0], f[1], f[2], C, pk);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

int launch_eb_likelihood(const float* zt, long long numel, int C, const float* pk, float* p, float* dp,
                         cudaStream_t st) {
  if (numel <= 0) return SGA_OK;
  k_eb_likelihood<<<blocks_for(numel, 256), 256, 0, st>>>(zt, numel, C, pk, p, dp);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

int launch_gauss_likelihood(const float* yt, const float* mu, const float* ls, long long numel, float* p, float* dpdy,
                            float* dpdsb, cudaStream_t st) {
  if (numel <= 0) return SGA_OK;
  k_gauss_likelihood<<<blocks_for(numel, 256), 256, 0, st>>>(yt, mu, ls, numel, p, dpdy, dpdsb);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

template <typename AT>
int launch_z_path(const float* z, const float* inj, long long numel, long long per_image, int C, const float* pk,
                  float T, float eps, uint64_t seed, int step, int hard, float rate_scale, AT* zt_act, float* g_rate,
                  float* dz, double* bits, long long ctr_off, cudaStream_t st, const StepParams* spt, const int* spi,
                  int split) {
  const long long pairs = (numel + 1) / 2;
  if (ew_exact())
    return launch_ew(k_z_path<AT, false>, blocks_for(pairs, 256), 256, st, z, inj, numel, per_image, C, pk, T, eps, seed, step,
                     hard, rate_scale, zt_act, g_rate, dz, bits, ctr_off, spt, spi, split);
  return launch_ew(k_z_path<AT>, blocks_for(pairs, 256), 256, st, z, inj, numel, per_image, C, pk, T, eps, seed, step, hard,
                   rate_scale, zt_act, g_rate, dz, bits, ctr_off, spt, spi, split);
}
template int launch_z_path<float>(const float*, const float*, long long, long long, int, const float*, float, float,
                                  uint64_t, int, int, float, float*, float*, float*, double*, long long, cudaStream_t,
                                  const StepParams*, const int*, int);
template int launch_z_path<__nv_bfloat16>(const float*, const float*, long long, long long, int, const float*, float,
                                          float, uint64_t, int, int, float, __nv_bfloat16*, float*, float*, double*,
                                          long long, cudaStream_t, const StepParams*, const int*, int);

template <typename AT>
int launch_y_path(const float* y, const float* inj, const float* h2, int N, int Hy, int Wy, int Hh, int Wh, int C,
                  float T, float eps, uint64_t seed, int step, int hard, float rate_scale, AT* yt_act, float* g_rate,
                  float* dy, AT* g_h2, double* bits, long long ctr_off, cudaStream_t st, const float* yt_in) {
  const long long pairs = static_cast<long long>(N) * Hh * Wh * (C / 2);
  if (ew_exact())
    return launch_ew(k_y_path<AT, false>, blocks_for(pairs, 256), 256, st, y, inj, h2, N, Hy, Wy, Hh, Wh, C, T, eps, seed, step,
                     hard, rate_scale, yt_act, g_rate, dy, g_h2, bits, ctr_off, yt_in);
  return launch_ew(k_y_path<AT>, blocks_for(pairs, 256), 256, st, y, inj, h2, N, Hy, Wy, Hh, Wh, C, T, eps, seed, step, hard,
                   rate_scale, yt_act, g_rate, dy, g_h2, bits, ctr_off, yt_in);
}
template int launch_y_path<float>(const float*, const float*, const float*, int, int, int, int, int, int, float, float,
                                  uint64_t, int, int, float, float*, float*, float*, float*, double*, long long, cudaStream_t,
                                  const float*);
template int launch_y_path<__nv_bfloat16>(const float*, const float*, const float*, int, int, int, int, int, int, float,
                                          float, uint64_t, int, int, float, __nv_bfloat16*, float*, float*,
                                          __nv_bfloat16*, double*, long long, cudaStream_t, const float*);

int launch_y_soft(const float* y, const float* inj, long long numel, int C, float T, float eps, uint64_t seed, int step,
                  float* yt_f32, __nv_bfloat16* yt_act, float* dy, long long ctr_off, cudaStream_t st, const StepParams* spt,
                  const int* spi, int split) {
  const long long pairs = numel / 2;
  if (ew_exact())
    return launch_ew(k_y_soft<false>, blocks_for(pairs, 256), 256, st, y, inj, numel, C, T, eps, seed, step, yt_f32, yt_act, dy,
                     ctr_off, spt, spi, split);
  return launch_ew(k_y_soft<true>, blocks_for(pairs, 256), 256, st, y, inj, numel, C, T, eps, seed, step, yt_f32, yt_act, dy,
                   ctr_off, spt, spi, split);
}

int launch_distortion(const float* x, const float* xt, int N, int H, int W, int Ho, int Wo, float grad_scale, float* gx,
                      double* sse, double* sse_r, cudaStream_t st) {
  const long long npix = static_cast<long long>(N) * Ho * Wo;
  k_distortion<<<blocks_for(npix, 256), 256, 0, st>>>(x, xt, N, H, W, Ho, Wo, grad_scale, gx, sse, sse_r);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

template <typename AT>
int launch_zbb_path(const float* z_mean, const float* z_logvar, const float* eps_inj, long long numel, long long per_image,
                    int C, const float* pk, uint64_t seed, int step, float rate_scale, AT* zt_act, float* g_zt, float* dzl,
                    double* bits, double* logq, long long ctr_off, cudaStream_t st) {
  const long long pairs = (numel + 1) / 2;
  return launch_ew(k_zbb_path<AT>, blocks_for(pairs, 256), 256, st, z_mean, z_logvar, eps_inj, numel, per_image, C, pk, seed, step,
                   rate_scale, zt_act, g_zt, dzl, bits, logq, ctr_off);
}
template int launch_zbb_path<float>(const float*, const float*, const float*, long long, long long, int, const float*,
                                    uint64_t, int, float, float*, float*, float*, double*, double*, long long, cudaStream_t);
template int launch_zbb_path<__nv_bfloat16>(const float*, const float*, const float*, long long, long long, int,
                                            const float*, uint64_t, int, float, __nv_bfloat16*, float*, float*, double*,
                                            double*, long long, cudaStream_t);

int launch_bb_grads(const float* g_rate, const float* g_hyper, const float* dzl, long long numel, float half_scale,
                    float* g_mean, float* g_lv, cudaStream_t st) {
  if (numel <= 0) return SGA_OK;
  return launch_ew(k_bb_grads, blocks_for(numel, 256), 256, st, g_rate, g_hyper, dzl, numel, half_scale, g_mean, g_lv);
}

int launch_split2(const float* in, long long pixels, int C, float* a, float* b, cudaStream_t st) {
  if (pixels <= 0) return SGA_OK;
  k_split2<<<blocks_for(pixels * C, 256), 256, 0, st>>>(in, pixels, C, a, b);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

int launch_finalize_stats_bb(const double* acc, int N, int n_ref, int H, int W, float lmbda, int rate_only, float* out,
                             cudaStream_t st) {
  return launch_ew(k_finalize_stats_bb, 1, 32, st, acc, N, n_ref, H, W, lmbda, rate_only, out);
}

int launch_finalize_stats(const double* acc, int N, int n_ref, int H, int W, float lmbda, float* out, cudaStream_t st,
                          int* spi_advance) {
  return launch_ew(k_finalize_stats, 1, 32, st, acc, N, n_ref, H, W, lmbda, out, spi_advance);
}

double adam_lr_t(double lr, double beta1, double beta2, int iterations) {
  const double t = iterations + 1.0;
  return lr * (sqrt(1.0 - pow(beta2, t)) / (1.0 - pow(beta1, t)));  // adam.py:40-42 (float64 scalar)
}

int launch_adam(float* p, float* m, float* v, const float* ga, const float* gb, const float* fac, long long numel,
                double lr, double beta1, double beta2, double eps, int iterations, float* g_out, cudaStream_t st,
                const StepParams* spt, const int* spi) {
  if (numel <= 0) return SGA_OK;
  const double lr_t = adam_lr_t(lr, beta1, beta2, iterations);
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const float f_lr = static_cast<float>(lr_t), f_b1 = static_cast<float>(beta1), f_omb1 = static_cast<float>(1.0 - beta1),
              f_b2 = static_cast<float>(beta2), f_omb2 = static_cast<float>(1.0 - beta2), f_eps = static_cast<float>(eps);
  long long done = 0;
  if (p != nullptr && g_out == nullptr && numel >= 4 && al16(p) && al16(m) && al16(v) && al16(ga) && al16(gb) && al16(fac)) {
    const long long n4 = numel / 4;
    SGA_TRY(launch_ew(k_adam_v4, blocks_for(n4, 256), 256, st, p, m, v, ga, gb, fac, n4, f_lr, f_b1, f_omb1, f_b2, f_omb2, f_eps,
                      spt, spi));
    done = n4 * 4;
  }
  if (done < numel) {   // tail (or everything: unaligned buffers, gradient-only calls)
    SGA_TRY(launch_ew(k_adam, blocks_for(numel - done, 256), 256, st, p ? p + done : nullptr, m ? m + done : nullptr,
                      v ? v + done : nullptr, ga + done, gb ? gb + done : nullptr, fac ? fac + done : nullptr, numel - done,
                      f_lr, f_b1, f_omb1, f_b2, f_omb2, f_eps, g_out ? g_out + done : nullptr, spt, spi));
  }
  return SGA_OK;
}

int launch_square(const float* x, long long n, float* o, cudaStream_t st) {
  k_square<<<blocks_for(n, 256), 256, 0, st>>>(x, n, o);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
int launch_gdn_finish(const float* x, const float* nrm, long long n, int inverse, float* out, float* s_out,
                      cudaStream_t st) {
  k_gdn_finish<<<blocks_for(n, 256), 256, 0, st>>>(x, nrm, n, inverse, out, s_out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
int launch_igdn_bwd_pre(const float* g, const float* x, const float* s, long long n, float* t, cudaStream_t st) {
  k_igdn_bwd_pre<<<blocks_for(n, 256), 256, 0, st>>>(g, x, s, n, t);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
int launch_igdn_bwd_finish(const float* g, const float* x, const float* s, const float* u, long long n, float* gx,
                           cudaStream_t st) {
  k_igdn_bwd_finish<<<blocks_for(n, 256), 256, 0, st>>>(g, x, s, u, n, gx);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
int launch_f32_to_bf16(const float* in, long long n, __nv_bfloat16* out, cudaStream_t st) {
  if (n <= 0) return SGA_OK;
  k_convert<float, __nv_bfloat16><<<blocks_for(n, 256), 256, 0, st>>>(in, n, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
int launch_bf16_to_f32(const __nv_bfloat16* in, long long n, float* out, cudaStream_t st) {
  if (n <= 0) return SGA_OK;
  k_convert<__nv_bfloat16, float><<<blocks_for(n, 256), 256, 0, st>>>(in, n, out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga
