// conv_plan.cu -- every (de)conv of nn_models.py expressed as dense multi-tap accumulations.
//
// A transposed 5x5 stride-2 convolution is split into its 4 output phases (3x3 / 3x2 / 2x3 / 2x2 taps, 25 in
// total) so that no zero-insertion MACs are executed; its dgrad -- and the encoder's stride-2 correlations --
// read the 4 stride-2 phase sub-grids of their input.  Index conventions:
//   up2        (corr=False, strides_up=2)   out[t]  = sum_i x[i] k[t + p - 2i],  p = kh/2
//   up2 dgrad                                g_in[i] = sum_a g[2i + a - p] k[a]      (kernel transposed in/out)
//   s1 true    (corr=False, stride 1)        out[t]  = sum_a x[t + p - a] k[a]
//   s1 true dgrad                            g_in[i] = sum_a g[i + a - p] k[a]       (transposed)
//   down2 corr (corr=True, strides_down=2)   out[i]  = sum_a x[2i + a - p] k[a]
//   s1 corr    (corr=True, stride 1)         out[i]  = sum_a x[i + a - p] k[a]
#include "kernels.h"

namespace sga {

static int floor_div2(int v) { return (v >= 0) ? v / 2 : -((-v + 1) / 2); }

int make_conv_plan(int kind, int kh, int kw, int N, int Hin, int Win, int Cin, int Cout_kernel, ConvPlan* plan) {
  SGA_CHECK(kh % 2 == 1 && kw % 2 == 1 && kh * kw <= kMaxTaps, "unsupported kernel support %dx%d", kh, kw);
  memset(plan, 0, sizeof(*plan));
  ConvGeom& g = plan->g;
  g.kind = kind; g.kh = kh; g.kw = kw; g.N = N; g.Hin = Hin; g.Win = Win;
  const int ph = kh / 2, pw = kw / 2;
  const bool dgrad = (kind == kUp2Dgrad || kind == kS1TrueDgrad);
  plan->w_transposed = dgrad;
  // channel roles: `Cin`/`Cout_kernel` are the kernel's (in,out) dims; a dgrad consumes Cout and produces Cin
  g.Cin = dgrad ? Cout_kernel : Cin;
  g.Cout = dgrad ? Cin : Cout_kernel;
  switch (kind) {
    case kUp2: g.Hout = 2 * Hin; g.Wout = 2 * Win; break;
    case kUp2Dgrad:
      SGA_CHECK(Hin % 2 == 0 && Win % 2 == 0, "up2 dgrad needs an even-sized gradient");
      g.Hout = Hin / 2; g.Wout = Win / 2; break;
    case kDown2Corr: g.Hout = (Hin + 1) / 2; g.Wout = (Win + 1) / 2; break;
    default: g.Hout = Hin; g.Wout = Win; break;
  }
  if (kind == kUp2) {
    plan->njobs = 4;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        ConvJobPlan& j = plan->job[py * 2 + px];
        j.nviews = 1; j.in_phase_stride = 1;
        j.out_py = py; j.out_px = px; j.out_phase_stride = 2;
        j.Ho = Hin; j.Wo = Win;
        // out[2i'+py] takes k[a] with a = py + ph - 2*oy for input row i'+oy
        for (int a = 0; a < kh; ++a) {
          if (((a - py - ph) % 2) != 0) continue;
          const int oy = (py + ph - a) / 2;
          for (int b = 0; b < kw; ++b) {
            if (((b - px - pw) % 2) != 0) continue;
            const int ox = (px + pw - b) / 2;
            j.taps[j.ntaps++] = Tap{0, oy, ox, a * kw + b};
          }
        }
      }
  } else if (kind == kUp2Dgrad || kind == kDown2Corr) {
    plan->njobs = 1;
    ConvJobPlan& j = plan->job[0];
    j.nviews = 4; j.in_phase_stride = 2;
    for (int v = 0; v < 4; ++v) { j.view_py[v] = v >> 1; j.view_px[v] = v & 1; }
    j.out_phase_stride = 1;
    j.Ho = g.Hout; j.Wo = g.Wout;
    for (int a = 0; a < kh; ++a) {
      const int ra = a - ph;                       // input row 2i + ra
      const int pa = ((ra % 2) + 2) % 2, oa = floor_div2(ra - pa);
      for (int b = 0; b < kw; ++b) {
        const int rb = b - pw;
        const int pb = ((rb % 2) + 2) % 2, ob = floor_div2(rb - pb);
        j.taps[j.ntaps++] = Tap{pa * 2 + pb, oa, ob, a * kw + b};
      }
    }
  } else {
    plan->njobs = 1;
    ConvJobPlan& j = plan->job[0];
    j.nviews = 1; j.in_phase_stride = 1; j.out_phase_stride = 1;
    j.Ho = g.Hout; j.Wo = g.Wout;
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b) {
        const int dy = (kind == kS1True) ? (ph - a) : (a - ph);
        const int dx = (kind == kS1True) ? (pw - b) : (b - pw);
        j.taps[j.ntaps++] = Tap{0, dy, dx, a * kw + b};
      }
  }
  return SGA_OK;
}

}  // namespace sga
