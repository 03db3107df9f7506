// conv_tc.cuh -- definitions shared by the tcgen05 (de)conv kernels (conv_tc.cu, conv_igdn_fwd_tc.cu)
#pragma once
#include "kernels.h"
#include "tc_common.cuh"

namespace sga {

using namespace tc;

constexpr int kPlainGroups = 2;                  // epilogue warp groups of the un-fused kernel
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;

// halo-tile A reuse (forward fusion, opt-in): one box {64 ch, 16, 18, 1} -- the 8 x 16-pixel tile plus its halo, padded to a
// 16-pixel row pitch -- per (input view, K chunk); every tap of the view addresses its window inside it: M row r =
// pixel (r/8, r%8) sits at byte ((r/8 + dy')*16 + dx' + r%8)*128, i.e. 8-row groups 2048 B apart (descriptor SBO)
constexpr int kHaloPitch = 16, kHaloRows = 18;
constexpr int kHaloBytes = kHaloPitch * kHaloRows * 64 * 2;   // 36 KB
constexpr int kMaxGroups = 16;
struct TcHaloTap { int widx; uint32_t a_off16; uint32_t boff; };        // weight slice, (window byte offset) >> 4, start row & 7
struct TcGroup { int view, ox, oy, tap_begin, ntaps; };                 // taps reading one input view; halo origin (ox, oy)

struct TcJob {
  int group_begin, ngroups;
  int tap_begin, ntaps;
  int Ho, Wo;             // pixel grid of the output view
  long long out_off;      // element offset of the output view
  long long oH, oW;       // element strides of the output view
};

struct alignas(64) TcParams {
  CUtensorMap amap[kMaxViews];
  CUtensorMap bmap;
  CUtensorMap gmap;       // fused IGDN: channel-mixing operand [n][k] bf16, box {64, N/2}
  TcJob job[4];
  Tap taps[kMaxTaps];
  TcGroup groups[kMaxGroups];
  TcHaloTap htaps[kMaxTaps];
  int halo;               // 0 = off; 1 = descriptor base_offset from the window start; 2 = base_offset 0
  int a_stages;           // halo tiles in flight
  int njobs, N, K, Nn;
  int block_n, n_tiles_n, kchunks, stages;
  int TH, TW, tiles_h, tiles_w;
  int n_spatial;          // tiles per job = N * tiles_h * tiles_w * n_tiles_n (identical for all jobs of a plan)
  int n_sp;               // spatial tiles per job and N tile = N * tiles_h * tiles_w
  int n_sp_pairs;         // CTA-pair mode: ceil(n_sp / 2)
  int spread;             // small problem: spread the output phases over the grid (decode_slot)
  int relu;
  int dbg;                // development aid (env SGA_DBG): bit 3 = per-role wait-cycle accounting printed by two CTAs
  int mix_at0, mix_at1;   // fused: k-step (in eighths of the main loop) where the mixing halves of the previous tile go
  long long oN;
  const float* bias;
  const __nv_bfloat16* mask;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  // fused IGDN
  const float* beta;            // fwd
  __nv_bfloat16* s_out;         // fwd: sqrt(norm)
  const __nv_bfloat16* x_in;    // bwd: saved act = x*s
  const __nv_bfloat16* s_in;    // bwd: saved s
};

struct TileCoord { int job, img, h0, w0, n0; };

// slot = blockIdx.x + k * gridDim.x: wave k handles phase (k % njobs) of spatial tiles [(k / njobs) * G, +G), so the
// phases of a transposed conv re-read their shared input from L2 and every CTA gets the same mix of 9/6/6/4-tap tiles.
// Small problems (fewer units than CTAs, P.spread): the phases are spread over the grid instead, unit L = k*G + cta ->
// (spatial L / njobs, phase L % njobs), so that a 12-tile layer occupies 12 CTA pairs once rather than 3 pairs 4 times.
// CG = 2: the unit is a pair of horizontally adjacent spatial tiles sharing the N tile; a pair whose second tile does
// not exist still runs it with img = N (TMA zero-fills the loads, nothing is stored).
template <int CG>
__device__ __forceinline__ bool decode_slot(const TcParams& P, int k, uint32_t rank, TileCoord& tc_) {
  const int g = static_cast<int>(gridDim.x) / CG, cta = static_cast<int>(blockIdx.x) / CG;
  const int units = (CG == 1) ? P.n_spatial : P.n_sp_pairs * P.n_tiles_n;
  int jb, u;
  if (P.spread) {
    const int L = k * g + cta;
    if (L >= units * P.njobs) return false;
    jb = L % P.njobs;
    u = L / P.njobs;
  } else {
    jb = k % P.njobs;
    u = (k / P.njobs) * g + cta;
    if (u >= units) return false;
  }
  int sp, nt;
  if (CG == 1) {
    nt = u % P.n_tiles_n;
    sp = u / P.n_tiles_n;
  } else {
    nt = u / P.n_sp_pairs;
    sp = (u % P.n_sp_pairs) * 2 + static_cast<int>(rank);
  }
  if (sp >= P.n_sp) {
    tc_ = TileCoord{jb, P.N, 0, 0, nt * P.block_n};
    return true;
  }
  const int tw = sp % P.tiles_w; sp /= P.tiles_w;
  const int th = sp % P.tiles_h;
  const int img = sp / P.tiles_h;
  tc_ = TileCoord{jb, img, th * P.TH, tw * P.TW, nt * P.block_n};
  return true;
}
template <int CG>
__device__ __forceinline__ int num_waves(const TcParams& P) {
  const int units = (CG == 1) ? P.n_spatial : P.n_sp_pairs * P.n_tiles_n;
  const int g = static_cast<int>(gridDim.x) / CG;
  if (P.spread) return (units * P.njobs + g - 1) / g;
  return ((units + g - 1) / g) * P.njobs;
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct TcConv {
  TcParams P;
  int smem_bytes;
  int grid;
  int fuse;   // 0, 1 (igdn fwd), 2 (igdn bwd)
  int cg;     // 1, or 2 = CTA pairs (cluster of two, tcgen05.mma.cta_group::2)
  ConvPlan plan;
};

bool pdl_enabled();   // programmatic dependent launch between the kernels of a step (env SGA_NO_PDL disables)

// conv_igdn_fwd_tc.cu: the forward (de)conv fused with the inverse GDN (fuse == 1)
int igdn_fwd_smem_fixed(int cout, int cg);          // shared memory outside the operand ring
int igdn_fwd_set_attr(int cout, int cg);
int launch_igdn_fwd_kernel(const TcConv* c, const TcParams& P, cudaStream_t st);

}  // namespace sga
