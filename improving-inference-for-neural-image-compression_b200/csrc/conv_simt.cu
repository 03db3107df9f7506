// conv_simt.cu -- fp32 multi-tap implicit GEMM on CUDA cores.
//
// Used where full float32 arithmetic is wanted: y_init/z_init (sga.py:207), the final hard-rounded evaluation
// (sga.py:244-247) and as the on-device cross-check of the bf16 tcgen05 path.  One CTA computes a 64 (pixels) x 64
// (output channels) tile; 256 threads, 4x4 outputs each, K staged through shared memory 16 channels at a time.
#include "kernels.h"

namespace sga {

struct SimtJob {
  View in[kMaxViews];
  int ntaps;
  Tap taps[kMaxTaps];
  float* out;
  long long oN, oH, oW;
  int Ho, Wo;
  const float* mask;  // same addressing as out, or null
};

struct SimtParams {
  SimtJob job;
  const float* w;     // [tap][K][Nn]
  const float* bias;  // [Nn] or null
  int N, K, Nn, relu;
};

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) k_conv_simt(const __grid_constant__ SimtParams P) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const SimtJob& J = P.job;
  const int tid = threadIdx.x;
  const long long M = static_cast<long long>(P.N) * J.Ho * J.Wo;
  const long long m0 = static_cast<long long>(blockIdx.x) * BM;
  const int n0 = blockIdx.y * BN;

  // A loader: thread -> (row = tid/4, 4 consecutive k = (tid%4)*4)
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const long long a_m = m0 + a_row;
  int a_n = 0, a_i = 0, a_j = 0;
  const bool a_valid = a_m < M;
  if (a_valid) {
    a_j = static_cast<int>(a_m % J.Wo);
    const long long t = a_m / J.Wo;
    a_i = static_cast<int>(t % J.Ho);
    a_n = static_cast<int>(t / J.Ho);
  }
  // B loader: thread -> (k = tid/16, 4 consecutive n = (tid%16)*4)
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  const int ty = tid >> 4, tx = tid & 15;  // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const bool k_vec = (P.K % 4 == 0);
  const bool n_vec = (P.Nn % 4 == 0);

  for (int t = 0; t < J.ntaps; ++t) {
    const Tap tap = J.taps[t];
    const View& V = J.in[tap.view];
    const int ii = a_i + tap.dy, jj = a_j + tap.dx;
    const bool pix_ok = a_valid && ii >= 0 && ii < V.H && jj >= 0 && jj < V.W;
    const float* a_ptr = static_cast<const float*>(V.base) + a_n * V.sN + ii * V.sH + jj * V.sW;
    const float* w_tap = P.w + static_cast<long long>(tap.widx) * P.K * P.Nn;
    for (int k0 = 0; k0 < P.K; k0 += BK) {
      // ---- stage A (64 x 16) and B (16 x 64)
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (pix_ok) {
        const int k = k0 + a_k;
        if (k_vec && k + 3 < P.K) {
          const float4 v = *reinterpret_cast<const float4*>(a_ptr + k);
          av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k + q < P.K) av[q] = a_ptr[k + q];
        }
      }
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const int k = k0 + b_k, n = n0 + b_n;
        if (k < P.K) {
          const float* wp = w_tap + static_cast<long long>(k) * P.Nn + n;
          if (n_vec && n + 3 < P.Nn) {
            const float4 v = *reinterpret_cast<const float4*>(wp);
            bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (n + q < P.Nn) bv[q] = wp[q];
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) As[a_k + q][a_row] = av[q];
      *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int oj = static_cast<int>(m % J.Wo);
    const long long t = m / J.Wo;
    const int oi = static_cast<int>(t % J.Ho);
    const int on = static_cast<int>(t / J.Ho);
    const long long off = on * J.oN + oi * J.oH + oj * J.oW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= P.Nn) continue;
      float v = acc[i][j];
      if (P.bias) v += P.bias[n];
      if (P.relu) v = fmaxf(v, 0.0f);
      if (J.mask) v = (J.mask[off + n] > 0.0f) ? v : 0.0f;
      J.out[off + n] = v;
    }
  }
}

int launch_conv_simt(const ConvPlan& plan, const float* x, const float* w, const float* bias, const float* relu_mask,
                     int relu, float* out, cudaStream_t st) {
  const ConvGeom& g = plan.g;
  for (int jn = 0; jn < plan.njobs; ++jn) {
    const ConvJobPlan& jp = plan.job[jn];
    SimtParams P;
    memset(&P, 0, sizeof(P));
    P.w = w; P.bias = bias; P.N = g.N; P.K = g.Cin; P.Nn = g.Cout; P.relu = relu;
    SimtJob& J = P.job;
    for (int v = 0; v < jp.nviews; ++v) {
      View& V = J.in[v];
      if (jp.in_phase_stride == 1) {
        V.base = x; V.sN = static_cast<long long>(g.Hin) * g.Win * g.Cin;
        V.sH = static_cast<long long>(g.Win) * g.Cin; V.sW = g.Cin; V.H = g.Hin; V.W = g.Win;
      } else {
        const int py = jp.view_py[v], px = jp.view_px[v];
        V.base = x + (static_cast<long long>(py) * g.Win + px) * g.Cin;
        V.sN = static_cast<long long>(g.Hin) * g.Win * g.Cin;
        V.sH = 2LL * g.Win * g.Cin; V.sW = 2LL * g.Cin;
        V.H = (g.Hin - py + 1) / 2; V.W = (g.Win - px + 1) / 2;
      }
    }
    J.ntaps = jp.ntaps;
    for (int t = 0; t < jp.ntaps; ++t) J.taps[t] = jp.taps[t];
    const long long out_off = (jp.out_phase_stride == 2)
                                  ? (static_cast<long long>(jp.out_py) * g.Wout + jp.out_px) * g.Cout : 0;
    J.out = out + out_off;
    J.mask = relu_mask ? relu_mask + out_off : nullptr;
    J.oN = static_cast<long long>(g.Hout) * g.Wout * g.Cout;
    J.oH = static_cast<long long>(jp.out_phase_stride) * g.Wout * g.Cout;
    J.oW = static_cast<long long>(jp.out_phase_stride) * g.Cout;
    J.Ho = jp.Ho; J.Wo = jp.Wo;
    const long long M = static_cast<long long>(g.N) * J.Ho * J.Wo;
    if (M == 0) continue;
    dim3 grid(static_cast<unsigned>((M + BM - 1) / BM), static_cast<unsigned>((g.Cout + BN - 1) / BN));
    k_conv_simt<<<grid, 256, 0, st>>>(P);
    SGA_LAUNCH_CHECK();
  }
  return SGA_OK;
}

__global__ void k_transpose_taps(const float* __restrict__ w, int taps, int Cin, int Cout, float* __restrict__ o) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(taps) * Cin * Cout;
  if (e >= total) return;
  const int ci = static_cast<int>(e % Cin);
  const long long r = e / Cin;
  const int co = static_cast<int>(r % Cout);
  const int t = static_cast<int>(r / Cout);
  o[e] = w[(static_cast<long long>(t) * Cin + ci) * Cout + co];
}

int launch_transpose_taps(const float* w_ref, int taps, int Cin, int Cout, float* w_t, cudaStream_t st) {
  const long long total = static_cast<long long>(taps) * Cin * Cout;
  k_transpose_taps<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w_ref, taps, Cin, Cout, w_t);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga
