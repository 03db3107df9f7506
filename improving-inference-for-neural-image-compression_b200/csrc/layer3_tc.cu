// layer3_tc.cu -- the last synthesis layer (5x5 stride-2 transposed conv, C -> 3 channels; nn_models.py:60-63),
// fused with the distortion term (sga.py:150-154, 167-174), and its dgrad (3 -> C).
//
// With only 3 output channels the per-phase implicit GEMM would have N = 3.  Instead:
//   forward ("scatter" form): P[pixel, (a,b,c)] = sum_cin act[pixel, cin] * k[a,b,cin,c]   -- one GEMM with
//     M = 128 input pixels, K = C, N = 75 (padded to 80); the epilogue overlap-adds the 5x5x3 patches (col2im) in
//     shared memory: out[2i+a-2, 2j+b-2, c] += P[(i,j),(a,b,c)].  Tiles carry a 1-pixel halo (8x16 loaded, 6x14
//     interior) so every output pixel is completed inside one CTA; the finished x_tilde values are compared with x
//     on the spot: SSE, rounded-SSE (PSNR) and d loss / d x_tilde (bf16, [N,Ho,Wo,4]) leave the kernel.
//   dgrad ("gather" form):  g_act[pixel, cin] = sum_{a,b,c} gx[2i+a-2, 2j+b-2, c] * k[a,b,cin,c]  -- M = 128 input
//     pixels, K = 75 (padded to 80), N = C.  The im2col A operand is built by 4 warps from a shared-memory copy of
//     the gradient patch, directly in the no-swizzle K-major core-matrix layout the tensor core reads.
#include <stdlib.h>
#include "kernels.h"
#include "tc_common.cuh"
#include "conv_tc.cuh"

namespace sga {

using namespace tc;

// =====================================================================================================
// forward
// =====================================================================================================
// warp 0 TMA, warp 1 MMA (REALLOC: + two idle warps completing the warpgroup), then kL3Groups epilogue groups of
// kL3GroupWarps warps; group g drains the tiles with it % kL3Groups == g (TMEM accumulator g, col2im buffer g).  A tile's
// epilogue is one dependent chain (accumulator -> registers -> shared memory -> barrier -> overlap-add -> global) whatever the
// number of warps that share it (8 warps per group: same time as 4), so throughput comes from the number of tiles in flight;
// L2 prefetch and ring depth change nothing.  What the kernel is left with per tile is the shared-memory traffic of the
// scatter / col2im formulation.
constexpr int kL3Groups = 3;
constexpr int kL3GroupWarps = 4;
constexpr int kL3GroupThreads = 32 * kL3GroupWarps;
// REALLOC: the two issue warps become a whole warpgroup (two of its warps idle) that drops to 56 registers, and the three
// epilogue warpgroups take the freed ones (setmaxnreg: 152 instead of the 128 that 14-16 warps start with) -- at 128 the
// epilogue spilled the prefetched image values and loop invariants inside its per-tile chain.
constexpr int l3f_issue_threads(bool realloc) { return realloc ? 128 : 64; }
constexpr int l3f_threads(bool realloc) { return l3f_issue_threads(realloc) + kL3Groups * kL3GroupThreads; }
constexpr int kL3MaxStages = 8;       // the ring takes what shared memory is left (7 stages at C = 192, 6 at C = 256)
constexpr int kL3NPad = 80;
constexpr int kPsStride = 81;
constexpr int kL3IntH = 6, kL3IntW = 14;   // interior of the 8x16 tile

// (img, tile row, tile column) of tile = first + k * stride for k = 0, 1, ...: advanced with carries instead of the two
// integer divisions per use (the worker warps of the fused dgrad needed three such pairs per tile, ~10 % of their
// instruction stream)
struct L3TileIter {
  int img, th, tw;
  int dimg, dth, dtw, tiles_h, tiles_w;
  __device__ __forceinline__ void init(int first, int stride, int th_n, int tw_n) {
    tiles_h = th_n; tiles_w = tw_n;
    const int per_img = th_n * tw_n;
    img = first / per_img; th = (first % per_img) / tw_n; tw = first % tw_n;
    dimg = stride / per_img; dth = (stride % per_img) / tw_n; dtw = stride % tw_n;
  }
  __device__ __forceinline__ void advance() {
    tw += dtw;
    if (tw >= tiles_w) { tw -= tiles_w; ++th; }
    th += dth;
    if (th >= tiles_h) { th -= tiles_h; ++img; }
    img += dimg;
  }
};

struct alignas(64) L3FwdParams {
  CUtensorMap amap;   // activation [N,Hin,Win,C] bf16
  CUtensorMap bmap;   // scatter weights [80][C] bf16
  const float* x;     // image [N,H,W,3]
  float* xt_out;      // optional x_tilde [N,2Hin,2Win,3]
  __nv_bfloat16* gx;  // optional d loss / d x_tilde [N,2Hin,2Win,4]
  double* sse;        // [N]
  double* sse_r;      // [N]
  float bias[3];
  float grad_scale;
  int N, Hin, Win, C, H, W;
  int tiles_h, tiles_w, total_tiles, kchunks;
  int prefetch;       // L2 prefetch distance in tiles (env SGA_L3_PREFETCH, default 0 = off)
  int stages;         // A-tile ring depth
};

template <bool REALLOC>
__global__ void __launch_bounds__(l3f_threads(REALLOC), 1) k_l3_fwd(const __grid_constant__ L3FwdParams P) {
  pdl_launch_dependents();
  constexpr int kIssueThreads = l3f_issue_threads(REALLOC), kIssueWarps = kIssueThreads / 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int b_chunk = kL3NPad * 128;                 // 10240
  const uint32_t b_base = base;
  const uint32_t a_base = base + P.kchunks * b_chunk;
  uint8_t* p_after = smem + P.kchunks * b_chunk + P.stages * 16384;
  float* Ps = reinterpret_cast<float*>(p_after);     // [kL3Groups][128][81]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_after + kL3Groups * 128 * kPsStride * 4);
  const uint32_t bar_b = smem_u32(bars);
  const uint32_t bar_full = bar_b + 8;
  const uint32_t bar_empty = bar_full + 8 * kL3MaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kL3MaxStages;
  const uint32_t bar_tempty = bar_tfull + 8 * kL3Groups;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kL3MaxStages + 2 * kL3Groups);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.amap);
    tma_prefetch_desc(&P.bmap);
    mbar_init(bar_b, 1);
    for (int s = 0; s < P.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < kL3Groups; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, kL3GroupWarps); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(s_tmem), 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_grid_dependency_wait();
  const uint32_t tmem_base = *s_tmem;
  const int tiles_per_img = P.tiles_h * P.tiles_w;

  // Issue loops are warp-uniform with one elected lane (as in conv_tc.cu): a single-lane loop makes the compiler wrap every
  // UTMALDG / UTCHMMA in an election loop with R2UR moves -- hundreds of cycles of dependent-issue latency per K chunk,
  // which at 3 K chunks per tile is what this kernel's fixed ~2300 cycles per tile were made of.
  if (warp < kIssueWarps) {
  if constexpr (REALLOC) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_b, P.kchunks * b_chunk);
      for (int kc = 0; kc < P.kchunks; ++kc) tma_load_2d(b_base + kc * b_chunk, &P.bmap, bar_b, kc * 64, 0);
    }
    int stage = 0; uint32_t phase = 0;
    const int pf = P.prefetch;           // tiles of L2 prefetch distance (0 = off)
    auto prefetch_tile = [&](int tile) {
      if (tile >= P.total_tiles) return;
      const int img = tile / tiles_per_img, tt = tile % tiles_per_img;
      const int h0 = (tt / P.tiles_w) * kL3IntH - 1, w0 = (tt % P.tiles_w) * kL3IntW - 1;
      if (elect_one())
        for (int kc = 0; kc < P.kchunks; ++kc) tma_prefetch_4d(&P.amap, kc * 64, w0, h0, img);
    };
    for (int d = 2; d < 2 + pf; ++d) prefetch_tile(blockIdx.x + d * gridDim.x);   // the ring itself runs ~2 tiles ahead
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const int img = tile / tiles_per_img, tt = tile % tiles_per_img;
      const int h0 = (tt / P.tiles_w) * kL3IntH - 1, w0 = (tt % P.tiles_w) * kL3IntW - 1;
      if (pf) prefetch_tile(tile + (2 + pf) * gridDim.x);
      for (int kc = 0; kc < P.kchunks; ++kc) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_full + 8 * stage, 16384);
          tma_load_4d(a_base + stage * 16384, &P.amap, bar_full + 8 * stage, kc * 64, w0, h0, img);
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = instr_desc_bf16(128, kL3NPad);
    // descriptors: only the 14-bit start-address field changes (stage / K chunk, K = 16 slice), added to a prebuilt base
    const uint64_t adesc0 = smem_desc_sw128(a_base), bdesc0 = smem_desc_sw128(b_base);
    mbar_wait(bar_b, 0);
    int stage = 0; uint32_t phase = 0; int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it % kL3Groups;
      mbar_wait(bar_tempty + 8 * buf, ((it / kL3Groups) & 1) ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 128;
      for (int kc = 0; kc < P.kchunks; ++kc) {
        mbar_wait(bar_full + 8 * stage, phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + static_cast<uint64_t>((stage * 16384) >> 4);
          const uint64_t bd = bdesc0 + static_cast<uint64_t>((kc * b_chunk) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (kc | kk) != 0 ? 1u : 0u);
          umma_commit(bar_empty + 8 * stage);
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(bar_tfull + 8 * buf);
    }
  }
  } else {
    if constexpr (REALLOC) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    // Epilogue groups of 4 warps: group g drains the tiles with it % kL3Groups == g, i.e. TMEM / P buffer g.  The x
    // values of a thread's output pixels are fetched one tile ahead (the col2im loop is otherwise bound by that latency).
    const int q = warp & 3;                                 // TMEM lane quarter this warp may read
    const int grp = (warp - kIssueWarps) / kL3GroupWarps;
    const int row = q * 32 + lane;
    const int et = (threadIdx.x - kIssueThreads) % kL3GroupThreads;    // thread index inside the group
    const int Ho = 2 * P.Hin, Wo = 2 * P.Win;
    constexpr int kPix = 2 * kL3IntH * 2 * kL3IntW;         // 336 output pixels per tile
    constexpr int kPer = (kPix + kL3GroupThreads - 1) / kL3GroupThreads;   // 3 per thread
    float acc_s = 0.0f, acc_r = 0.0f;
    int acc_img = -1;
    auto flush = [&]() {
      const float a = warp_sum(acc_s), b = warp_sum(acc_r);
      if (lane == 0 && acc_img >= 0) {
        atomicAdd(P.sse + acc_img, static_cast<double>(a));
        atomicAdd(P.sse_r + acc_img, static_cast<double>(b));
      }
      acc_s = 0.0f; acc_r = 0.0f;
    };
    // Everything about a thread's (up to) three output pixels that does not depend on the tile is computed once: position
    // inside the tile, offsets into the image / gradient tensors relative to the tile origin, and the address of tap
    // (o_y, o_x) = (0, 0), kernel index (py + 2, px + 2) inside the partial-sum buffer -- every other tap is a compile-time
    // offset from it.  (Recomputed per tile this address arithmetic was a third of the epilogue's instructions.)
    int ry[kPer], rx[kPer], xoff[kPer], ooff[kPer];
    bool pyz[kPer], pxz[kPer];               // the o = -1 taps exist (even output row / column)
    const float* pq[kPer];
    float* const Pb = Ps + grp * 128 * kPsStride;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int p = et + k * kL3GroupThreads;
      ry[k] = p / (2 * kL3IntW); rx[k] = p % (2 * kL3IntW);
      if (p >= kPix) { ry[k] = 1 << 20; rx[k] = 0; }          // never inside the image: fails every bound below
      const int py = ry[k] & 1, px = rx[k] & 1;
      pyz[k] = (py == 0); pxz[k] = (px == 0);
      const int tr = ((ry[k] >> 1) + 1) & 7, tcx = (rx[k] >> 1) + 1;   // tile coordinates of the co-located input pixel
      pq[k] = Pb + (tr * 16 + tcx) * kPsStride + ((py + 2) * 5 + (px + 2)) * 3;
      const int ryc = (p >= kPix) ? 0 : ry[k];
      xoff[k] = (ryc * P.W + rx[k]) * 3;
      ooff[k] = ryc * Wo + rx[k];
    }
    float xn[kPer][3];
    auto fetch_x = [&](const L3TileIter& t) {
      const int oy0 = 2 * t.th * kL3IntH, ox0 = 2 * t.tw * kL3IntW;
      const float* xb = P.x + ((static_cast<long long>(t.img) * P.H + oy0) * P.W + ox0) * 3;
      const int ly = P.H - oy0, lx = P.W - ox0;
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const bool in = (ry[k] < ly) && (rx[k] < lx);
#pragma unroll
        for (int c = 0; c < 3; ++c) xn[k][c] = in ? __ldg(xb + xoff[k] + c) : 0.0f;
      }
    };
    const int first = blockIdx.x + grp * gridDim.x;
    L3TileIter ti, tn;                   // this group's current tile / the tile whose x values are fetched (one ahead)
    ti.init(first, kL3Groups * static_cast<int>(gridDim.x), P.tiles_h, P.tiles_w);
    tn = ti;
    if (first < P.total_tiles) fetch_x(tn);
    int it = grp;
    for (int tile = first; tile < P.total_tiles; tile += kL3Groups * gridDim.x, it += kL3Groups, ti.advance()) {
      const int img = ti.img;
      const int oy0 = 2 * ti.th * kL3IntH, ox0 = 2 * ti.tw * kL3IntW;
      const int buf = grp;
      if (img != acc_img) { flush(); acc_img = img; }   // warp-uniform
      float xc[kPer][3];
#pragma unroll
      for (int k = 0; k < kPer; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) xc[k][c] = xn[k][c];
      tn.advance();
      if (tile + kL3Groups * static_cast<int>(gridDim.x) < P.total_tiles) fetch_x(tn);
      mbar_wait(bar_tfull + 8 * buf, (it / kL3Groups) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + buf * 128 + (static_cast<uint32_t>(q * 32) << 16);
      {
        uint32_t r[32], r1[32], r2[16];
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 32, r1);
        tmem_ld16(taddr + 64, r2);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) Pb[row * kPsStride + k] = __uint_as_float(r[k]);
#pragma unroll
        for (int k = 0; k < 32; ++k) Pb[row * kPsStride + 32 + k] = __uint_as_float(r1[k]);
#pragma unroll
        for (int k = 0; k < 16; ++k) Pb[row * kPsStride + 64 + k] = __uint_as_float(r2[k]);
      }
      tcgen05_fence_before();
      named_bar_sync(1 + grp, kL3GroupThreads);   // all 128 rows of P are in shared memory; TMEM buffer drained
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      // ---- col2im + distortion: 12 x 28 output pixels of this tile's interior
      const long long opix0 = (static_cast<long long>(img) * Ho + oy0) * Wo + ox0;
      const int ly = Ho - oy0, lx = Wo - ox0, lyx = P.H - oy0, lxx = P.W - ox0;
      // overlap-add of the thread's three pixels side by side (nine independent accumulation chains)
      float v[kPer][3];
#pragma unroll
      for (int k = 0; k < kPer; ++k) { v[k][0] = P.bias[0]; v[k][1] = P.bias[1]; v[k][2] = P.bias[2]; }
      // out[2i'+py] takes k[a], a = py + 2 - 2o, from input row i'+o (conv_plan.cu): o = -1 only for py = 0
#pragma unroll
      for (int oyy = 1; oyy >= -1; --oyy)
#pragma unroll
        for (int oxx = 1; oxx >= -1; --oxx)
#pragma unroll
          for (int k = 0; k < kPer; ++k) {
            const float* pp = pq[k] + (oyy * 16 + oxx) * kPsStride - (10 * oyy + 2 * oxx) * 3;
            // predicated, not masked: a tap that does not exist for the pixel's parity is not loaded at all (36 % of the
            // 81 loads; in the power-bound regime an executed shared-memory load costs even when its result is discarded)
            if ((oyy >= 0 || pyz[k]) && (oxx >= 0 || pxz[k])) { v[k][0] += pp[0]; v[k][1] += pp[1]; v[k][2] += pp[2]; }
          }
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        if (ry[k] >= ly || rx[k] >= lx) continue;
        const float v0 = v[k][0], v1 = v[k][1], v2 = v[k][2];
        const long long opix = opix0 + ooff[k];
        if (P.xt_out) { P.xt_out[opix * 3] = v0; P.xt_out[opix * 3 + 1] = v1; P.xt_out[opix * 3 + 2] = v2; }
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        if (ry[k] < lyx && rx[k] < lxx) {
          const float vv[3] = {v0, v1, v2};
          float gg[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float xv = xc[k][c];
            const float d = vv[c] - xv;
            acc_s += d * d;
            const float rr = rintf(fminf(fmaxf(vv[c], 0.0f), 1.0f) * 255.0f) - xv * 255.0f;
            acc_r += rr * rr;
            gg[c] = P.grad_scale * d;
          }
          g0 = gg[0]; g1 = gg[1]; g2 = gg[2];
        }
        if (P.gx) *reinterpret_cast<uint2*>(P.gx + opix * 4) = make_uint2(pack_bf16(g0, g1), pack_bf16(g2, 0.0f));
      }
      named_bar_sync(1 + grp, kL3GroupThreads);   // P buffer may be overwritten by this group's next tile
    }
    flush();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================================
// dgrad
// =====================================================================================================
constexpr int kL3dThreads = 288;      // warp 0 MMA, warps 1-4 builders, warps 5-8 epilogue
constexpr int kRegH = 19, kRegW = 35; // gradient patch needed by an 8x16 input tile

struct alignas(64) L3DgradParams {
  const __nv_bfloat16* gx;     // [N,Ho,Wo,4]
  const __nv_bfloat16* wpack;  // [10][C][8] bf16 : B operand in core-matrix order
  __nv_bfloat16* out;          // [N,Hin,Win,C]
  int N, Hin, Win, C;
  int tiles_h, tiles_w, total_tiles;
};

__global__ void __launch_bounds__(kL3dThreads, 1) k_l3_dgrad(const __grid_constant__ L3DgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int b_bytes = 10 * P.C * 16;
  const int a_bytes = 10 * 128 * 16;                 // 20480
  const int reg_bytes = ((kRegH * kRegW * 8 + 127) / 128) * 128;
  const uint32_t b_base = base;
  const uint32_t a_base = base + ((b_bytes + 1023) / 1024) * 1024;
  uint8_t* a_ptr = smem + ((b_bytes + 1023) / 1024) * 1024;
  uint8_t* reg_ptr = a_ptr + 2 * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reg_ptr + 2 * reg_bytes);
  const uint32_t bar_afull = smem_u32(bars);         // [2] count 128
  const uint32_t bar_aempty = bar_afull + 16;        // [2] commit
  const uint32_t bar_tfull = bar_aempty + 16;        // [2] commit
  const uint32_t bar_tempty = bar_tfull + 16;        // [2] count 4
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // resident B operand: straight copy (already in core-matrix order)
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(P.wpack) + i);
  fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_afull + 8 * b, 128); mbar_init(bar_aempty + 8 * b, 1);
        mbar_init(bar_tfull + 8 * b, 1);   mbar_init(bar_tempty + 8 * b, 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int tiles_per_img = P.tiles_h * P.tiles_w;
  const int Ho = 2 * P.Hin, Wo = 2 * P.Win;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc_bf16(128, P.C);
      int it = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * buf, ph ^ 1u);
        mbar_wait(bar_afull + 8 * buf, ph);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
          const uint64_t ad = smem_desc_interleave(a_base + buf * a_bytes + ks * 2 * 2048, 2048, 128);
          const uint64_t bd = smem_desc_interleave(b_base + ks * 2 * (P.C * 16), P.C * 16, 128);
          umma_bf16(d_tmem, ad, bd, idesc, ks != 0 ? 1u : 0u);
        }
        umma_commit(bar_aempty + 8 * buf);
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else if (warp <= 4) {
    // ---------------- builders: gradient patch -> im2col A operand ----------------
    const int bt = threadIdx.x - 32;    // 0..127 == tile row
    const int tr = bt >> 4, tcx = bt & 15;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int img = tile / tiles_per_img, tt = tile % tiles_per_img;
      const int h0 = (tt / P.tiles_w) * 8, w0 = (tt % P.tiles_w) * 16;
      uint2* reg = reinterpret_cast<uint2*>(reg_ptr + buf * reg_bytes);
      // region origin (2*h0-2, 2*w0-2); the previous user of reg[buf]/A[buf] was tile it-2
      mbar_wait(bar_aempty + 8 * buf, ph ^ 1u);
      for (int e = bt; e < kRegH * kRegW; e += 128) {
        const int ry = e / kRegW, rx = e % kRegW;
        const int gy = 2 * h0 - 2 + ry, gxx = 2 * w0 - 2 + rx;
        uint2 v = make_uint2(0u, 0u);
        if (gy >= 0 && gy < Ho && gxx >= 0 && gxx < Wo)
          v = __ldg(reinterpret_cast<const uint2*>(P.gx + ((static_cast<long long>(img) * Ho + gy) * Wo + gxx) * 4));
        reg[e] = v;
      }
      named_bar_sync(2, 128);
      // row bt = input pixel (tr, tcx): 25 patch pixels x 3 channels = 75 K values (+5 zeros)
      uint16_t kv[80];
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          const uint2 v = reg[(2 * tr + a) * kRegW + (2 * tcx + b)];
          kv[(a * 5 + b) * 3 + 0] = static_cast<uint16_t>(v.x & 0xFFFFu);
          kv[(a * 5 + b) * 3 + 1] = static_cast<uint16_t>(v.x >> 16);
          kv[(a * 5 + b) * 3 + 2] = static_cast<uint16_t>(v.y & 0xFFFFu);
        }
#pragma unroll
      for (int k = 75; k < 80; ++k) kv[k] = 0;
      uint8_t* ab = a_ptr + buf * a_bytes;
#pragma unroll
      for (int gq = 0; gq < 10; ++gq) {
        uint4 o;
        o.x = kv[gq * 8 + 0] | (static_cast<uint32_t>(kv[gq * 8 + 1]) << 16);
        o.y = kv[gq * 8 + 2] | (static_cast<uint32_t>(kv[gq * 8 + 3]) << 16);
        o.z = kv[gq * 8 + 4] | (static_cast<uint32_t>(kv[gq * 8 + 5]) << 16);
        o.w = kv[gq * 8 + 6] | (static_cast<uint32_t>(kv[gq * 8 + 7]) << 16);
        *reinterpret_cast<uint4*>(ab + gq * 2048 + bt * 16) = o;
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_afull + 8 * buf);
    }
  } else {
    // ---------------- epilogue ----------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int img = tile / tiles_per_img, tt = tile % tiles_per_img;
      const int i = (tt / P.tiles_w) * 8 + (row >> 4), j = (tt % P.tiles_w) * 16 + (row & 15);
      const bool valid = (i < P.Hin) && (j < P.Win);
      const long long off0 = ((static_cast<long long>(img) * P.Hin + i) * P.Win + j) * P.C;
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + buf * 256 + (static_cast<uint32_t>(q * 32) << 16);
      for (int col = 0; col < P.C; col += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + col, r);
        tmem_ld_wait();
        if (valid) {
          uint4* op = reinterpret_cast<uint4*>(P.out + off0 + col);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            op[u] = make_uint4(pack_bf16(__uint_as_float(r[u * 8]), __uint_as_float(r[u * 8 + 1])),
                               pack_bf16(__uint_as_float(r[u * 8 + 2]), __uint_as_float(r[u * 8 + 3])),
                               pack_bf16(__uint_as_float(r[u * 8 + 4]), __uint_as_float(r[u * 8 + 5])),
                               pack_bf16(__uint_as_float(r[u * 8 + 6]), __uint_as_float(r[u * 8 + 7])));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================================
// dgrad fused with the backward of the inverse GDN that precedes layer 3 (IGDN_2, nn_models.py:59)
//   g   = dgrad accumulator (d loss / d act_2),  t = g*x/s,  u = t @ gamma^T,  g_x = g*s + x*u      (bf16 out)
// One persistent CTA per SM: warp 0 issues the MMAs, warp 1 streams the saved x / s tiles in with TMA (the x tile lands
// in the buffer that pass 1 then overwrites in place with the mixing operand t), 4*G worker warps (G = C/64 column
// groups x 4 TMEM lane quarters) build the im2col A operand of the NEXT tile from a gradient patch they fetched one tile
// ahead and run the two point-wise passes of the current one.  gamma (C x C, K-major, SWIZZLE_128B) and the gather
// weights stay resident in shared memory.  TMEM: accumulator | mixing result.
// Per tile:  MMA(t) -> pass 1 (g, x, s -> t) -> mixing MMA || build(t+1), TMA x/s(t+1) -> pass 2 (stores g_x).
// =====================================================================================================
struct alignas(64) L3DgradFusedParams {
  CUtensorMap gmap;            // gamma [n = C rows][k = C cols] bf16, box {64, C}
  CUtensorMap xmap, smap;      // saved act_2 (= x_2*s_2) / s_2 [N,Hin,Win,C] bf16, box {64, 16, 8, 1}
  const __nv_bfloat16* x_ptr;  // the same tensors as plain pointers (NF = 256: read straight from global memory)
  const __nv_bfloat16* s_ptr;
  const __nv_bfloat16* gx;     // [N,Ho,Wo,4]
  const __nv_bfloat16* wpack;  // [10][C][8]
  __nv_bfloat16* out;          // d loss / d x_2
  int N, Hin, Win, C;
  int tiles_h, tiles_w, total_tiles;
  int prefetch;                // L2 prefetch distance of the x / s tiles, in tiles (env SGA_L3_PREFETCH, default 0 = off)
};

__device__ __forceinline__ float l3_bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float l3_bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ float l3_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-groups [G0, G0+NG) (8 bf16 = 16 bytes each) of im2col row `row` = input pixel (tr, tcx) of the tile
template <int G0, int NG>
__device__ __forceinline__ void l3_build_groups(const uint2* __restrict__ reg, int tr, int tcx, uint8_t* ab, int row) {
  constexpr int K0 = G0 * 8, K1 = (G0 + NG) * 8;
  constexpr int T0 = K0 / 3, T1 = ((K1 - 1) / 3 < 24) ? (K1 - 1) / 3 : 24;
  uint16_t kv[NG * 8];
#pragma unroll
  for (int k = 0; k < NG * 8; ++k) kv[k] = 0;
#pragma unroll
  for (int t = T0; t <= T1; ++t) {
    const int a = t / 5, b = t % 5;
    const uint2 v = reg[(2 * tr + a) * kRegW + (2 * tcx + b)];
    const uint16_t ch[3] = {static_cast<uint16_t>(v.x & 0xFFFFu), static_cast<uint16_t>(v.x >> 16),
                            static_cast<uint16_t>(v.y & 0xFFFFu)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int k = t * 3 + c;
      if (k >= K0 && k < K1) kv[k - K0] = ch[c];
    }
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    uint4 o;
    o.x = kv[g * 8 + 0] | (static_cast<uint32_t>(kv[g * 8 + 1]) << 16);
    o.y = kv[g * 8 + 2] | (static_cast<uint32_t>(kv[g * 8 + 3]) << 16);
    o.z = kv[g * 8 + 4] | (static_cast<uint32_t>(kv[g * 8 + 5]) << 16);
    o.w = kv[g * 8 + 6] | (static_cast<uint32_t>(kv[g * 8 + 7]) << 16);
    *reinterpret_cast<uint4*>(ab + (G0 + g) * 2048 + row * 16) = o;
  }
}

// NF = 256 ("direct" mode): gamma alone is 128 KB, so there is no room for x / s tiles and resident gather weights.  The
// workers read their act / s chunks straight from global memory (64-byte row segments, prefetched before the accumulator
// is awaited), and ONE 64 KB buffer holds the gather weights (40 KB) while the gather GEMM runs and the mixing operand t
// afterwards; warp 1 re-loads the weights (one 1-D bulk copy from L2) as soon as the mixing GEMM has read t.
// In direct mode the two point-wise passes are STATELESS: pass 1 only writes the mixing operand, pass 2 re-reads the
// accumulator g (which therefore stays alive until pass 2 is done: the next tile's gather GEMM waits for it) and re-loads
// act / s (L2 hits).  Keeping g*s and x as packed registers across the mixing GEMM, as the NF <= 192 variant does, needs
// 32 registers per 32-column chunk on top of the working set: with 256 columns that spilled at every thread count tried
// (576 threads / 96 registers: 2.0 ms, 68 % of the stall samples on local-memory loads).
constexpr int l3_groups(int nf) { return nf / 64; }
// PAIR (NF <= 192): the lanes of a pair (two horizontally adjacent pixels) exchange half of their output so that every
// store instruction writes 64 contiguous bytes per pixel.  A warp store of one 32-byte sector per lane touches 32 lines and
// drains at about one line per clock: with all 12 worker warps storing at the end of a tile that was ~1500 clocks per tile
// of write-after-read stalls on the store operands (17 % of the stall samples); 16 lines per instruction: 0.67 -> 0.62 ms.
template <int NF, bool PAIR = (NF != 256)>
__global__ void __launch_bounds__(64 + 128 * l3_groups(NF), 1) k_l3_dgrad_igdn(const __grid_constant__ L3DgradFusedParams P) {
  pdl_launch_dependents();
  constexpr bool kDirect = (NF == 256);
  static_assert(!(kDirect && PAIR), "the direct mode stores 16-column units");
  constexpr int G = l3_groups(NF);           // worker groups (column split)
  constexpr int NH = NF / (32 * G);          // 32-column chunks per worker warp: chunk c = grp + h * G, h < NH
  constexpr int kWorkers = 128 * G;
  constexpr int kKc = NF / 64;               // K chunks of the mixing GEMM
  constexpr int kGammaBytes = kKc * NF * 128;
  constexpr int kTileBytes = kKc * 16384;    // one [128 px x NF] bf16 tile in SWIZZLE_128B K-major chunks
  constexpr int kBBytes = 10 * NF * 16;
  constexpr int kBPad = ((kBBytes + 1023) / 1024) * 1024;
  constexpr int kAIm = 10 * 128 * 16;        // 20480: im2col A operand (no swizzle)
  constexpr int kRegBytes = ((kRegH * kRegW * 8 + 127) / 128) * 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  // gamma | x tile (becomes the mixing operand t in place) | s tile | gather weights | im2col A | gradient patch | barriers
  const uint32_t gamma_base = base;
  const uint32_t a2_base = base + kGammaBytes;
  uint8_t* a2_ptr = smem + kGammaBytes;
  const uint32_t s_base = a2_base + kTileBytes;
  uint8_t* s_ptr = a2_ptr + kTileBytes;
  // direct mode: no s tile; the gather weights share the x / t buffer
  const uint32_t b_base = kDirect ? a2_base : s_base + kTileBytes;
  uint8_t* b_ptr = kDirect ? a2_ptr : s_ptr + kTileBytes;
  const uint32_t a_base = kDirect ? a2_base + kTileBytes : b_base + kBPad;
  uint8_t* a_ptr = kDirect ? a2_ptr + kTileBytes : b_ptr + kBPad;
  uint8_t* reg_ptr = a_ptr + kAIm;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reg_ptr + kRegBytes);
  const uint32_t bar_g = smem_u32(bars);             // gamma loaded (tx)
  const uint32_t bar_afull = bar_g + 8;              // im2col operand built (count kWorkers)
  const uint32_t bar_aempty = bar_afull + 8;         // ... consumed (commit)
  const uint32_t bar_tfull = bar_aempty + 8;         // accumulator ready (commit)
  const uint32_t bar_a2 = bar_tfull + 8;             // pass 1 done: t in smem, accumulator + s tile drained (count kWorkers)
  const uint32_t bar_nfull = bar_a2 + 8;             // mixing result ready (commit); also: x/t tile free again
  const uint32_t bar_nempty = bar_nfull + 8;         // mixing result drained (count 4*G)
  const uint32_t bar_xfull = bar_nempty + 8;         // x tile landed (tx)
  const uint32_t bar_sfull = bar_xfull + 8;          // s tile landed (tx)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!kDirect) {
    for (int i = threadIdx.x; i < kBBytes / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(b_ptr)[i] = __ldg(reinterpret_cast<const uint4*>(P.wpack) + i);
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&P.gmap);
      if (!kDirect) {
        tma_prefetch_desc(&P.xmap);
        tma_prefetch_desc(&P.smap);
      }
      mbar_init(bar_g, 1);
      mbar_init(bar_afull, kWorkers);
      mbar_init(bar_aempty, 1);
      mbar_init(bar_tfull, 1);
      mbar_init(bar_a2, kWorkers);
      mbar_init(bar_nfull, 1);
      mbar_init(bar_nempty, 4 * G);
      mbar_init(bar_xfull, 1);
      mbar_init(bar_sfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_grid_dependency_wait();        // above: gather weights (constants) and on-chip state only
  const uint32_t tmem_base = *s_tmem;
  const int tiles_per_img = P.tiles_h * P.tiles_w;
  const int Ho = 2 * P.Hin, Wo = 2 * P.Win;
  constexpr uint32_t kNormCol = NF;

  if (warp == 0) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_g, kGammaBytes);
      for (int kc = 0; kc < kKc; ++kc) tma_load_2d(gamma_base + kc * NF * 128, &P.gmap, bar_g, kc * 64, 0);
    }
    mbar_wait(bar_g, 0);
    const uint32_t idesc = instr_desc_bf16(128, NF);
    const uint64_t a2desc0 = smem_desc_sw128(a2_base), gdesc0 = smem_desc_sw128(gamma_base);
    const uint64_t adesc0 = smem_desc_interleave(a_base, 2048, 128), bdesc0 = smem_desc_interleave(b_base, NF * 16, 128);
    int it = 0;
    auto mix = [&](int j) {   // mixing GEMM of tile j: [128 x NF] = t[128 x NF] * gamma^T
      mbar_wait(bar_a2, j & 1);                     // pass 1 of tile j done: operand in smem, accumulator drained
      mbar_wait(bar_nempty, (j & 1) ^ 1u);          // pass 2 of tile j-1 has read the previous mixing result
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < kKc; ++kc)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_base + kNormCol, a2desc0 + static_cast<uint64_t>((kc * 16384 + kk * 32) >> 4),
                      gdesc0 + static_cast<uint64_t>((kc * NF * 128 + kk * 32) >> 4), idesc, (kc | kk) != 0 ? 1u : 0u);
        umma_commit(bar_nfull);
      }
    };
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      if (it > 0) mix(it - 1);                      // also guarantees the accumulator of tile it-1 is drained
      mbar_wait(bar_afull, it & 1);
      if (kDirect) {
        mbar_wait(bar_xfull, it & 1);               // the gather weights are back in the shared x / t buffer
        if (it > 0) mbar_wait(bar_nempty, (it - 1) & 1);   // pass 2 of the previous tile has re-read its accumulator
      }
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
          umma_bf16(tmem_base, adesc0 + static_cast<uint64_t>((ks * 2 * 2048) >> 4),
                    bdesc0 + static_cast<uint64_t>((ks * 2 * (NF * 16)) >> 4), idesc, ks != 0 ? 1u : 0u);
        umma_commit(bar_aempty);
        umma_commit(bar_tfull);
      }
    }
    if (it > 0) mix(it - 1);
  } else if (warp == 1) {
    // ===================== TMA producer: saved x and s tiles (warp-uniform loop) =====================
    int it = 0;
    if (kDirect) {
      // the only thing to stream is the gather-weight block: before tile 0, then after every mixing GEMM (which reads t
      // out of the same buffer)
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
        if (it > 0) mbar_wait(bar_nfull, (it - 1) & 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_xfull, kBBytes);
          bulk_load_1d(b_base, P.wpack, kBBytes, bar_xfull);
        }
      }
    } else
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int img = tile / tiles_per_img, tt = tile % tiles_per_img;
      const int h0 = (tt / P.tiles_w) * 8, w0 = (tt % P.tiles_w) * 16;
      if (P.prefetch) {
        // the x / s buffers are single (shared memory is full), so the loads of tile it+1 can only be issued late in tile
        // it: pull them into L2 a whole tile earlier (opt-in: measured no gain)
        const int nt = tile + P.prefetch * static_cast<int>(gridDim.x);
        if (nt < P.total_tiles && elect_one()) {
          const int nimg = nt / tiles_per_img, ntt = nt % tiles_per_img;
          const int nh0 = (ntt / P.tiles_w) * 8, nw0 = (ntt % P.tiles_w) * 16;
          for (int kc = 0; kc < kKc; ++kc) {
            tma_prefetch_4d(&P.smap, kc * 64, nw0, nh0, nimg);
            tma_prefetch_4d(&P.xmap, kc * 64, nw0, nh0, nimg);
          }
        }
      }
      if (it > 0) mbar_wait(bar_a2, (it - 1) & 1);          // pass 1 of the previous tile has consumed its s tile
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_sfull, kTileBytes);
        for (int kc = 0; kc < kKc; ++kc) tma_load_4d(s_base + kc * 16384, &P.smap, bar_sfull, kc * 64, w0, h0, img);
      }
      if (it > 0) mbar_wait(bar_nfull, (it - 1) & 1);       // the mixing GEMM has finished reading t of the previous tile
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_xfull, kTileBytes);
        for (int kc = 0; kc < kKc; ++kc) tma_load_4d(a2_base + kc * 16384, &P.xmap, bar_xfull, kc * 64, w0, h0, img);
      }
    }
  } else {
    // ===================== workers =====================
    const int wt = threadIdx.x - 64;            // worker thread
    const int grp = (warp - 2) >> 2;            // column group
    const int q = warp & 3;                     // TMEM lane quarter
    const int row = q * 32 + lane;              // pixel row for the point-wise passes
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int brow = wt & 127, bpart = wt >> 7; // im2col: row and K part
    constexpr int kPatchPer = (kRegH * kRegW + kWorkers - 1) / kWorkers;
    uint2 pre[kPatchPer];                       // gradient patch entries of the tile built next, fetched one tile ahead

    auto fetch_patch = [&](const L3TileIter& t) {
      const int img = t.img;
      const int h0 = t.th * 8, w0 = t.tw * 16;
#pragma unroll
      for (int k = 0; k < kPatchPer; ++k) {
        const int e = wt + k * kWorkers;
        const int ry = e / kRegW, rx = e % kRegW;
        const int gy = 2 * h0 - 2 + ry, gxx = 2 * w0 - 2 + rx;
        pre[k] = make_uint2(0u, 0u);
        if (e < kRegH * kRegW && gy >= 0 && gy < Ho && gxx >= 0 && gxx < Wo)
          pre[k] = __ldg(reinterpret_cast<const uint2*>(P.gx + ((static_cast<long long>(img) * Ho + gy) * Wo + gxx) * 4));
      }
    };
    auto build = [&](int it) {                  // im2col operand of tile `it` from the prefetched patch
      uint2* reg = reinterpret_cast<uint2*>(reg_ptr);
      if (it > 0) mbar_wait(bar_aempty, (it - 1) & 1);            // MMA of the previous tile has read the operand
#pragma unroll
      for (int k = 0; k < kPatchPer; ++k) {
        const int e = wt + k * kWorkers;
        if (e < kRegH * kRegW) reg[e] = pre[k];
      }
      named_bar_sync(2, kWorkers);
      const int tr = brow >> 4, tcx = brow & 15;
      if (G == 3) {
        if (bpart == 0) l3_build_groups<0, 4>(reg, tr, tcx, a_ptr, brow);
        else if (bpart == 1) l3_build_groups<4, 3>(reg, tr, tcx, a_ptr, brow);
        else l3_build_groups<7, 3>(reg, tr, tcx, a_ptr, brow);
      } else if (G == 2) {
        if (bpart == 0) l3_build_groups<0, 5>(reg, tr, tcx, a_ptr, brow);
        else l3_build_groups<5, 5>(reg, tr, tcx, a_ptr, brow);
      } else if (G == 4) {
        if (bpart == 0) l3_build_groups<0, 3>(reg, tr, tcx, a_ptr, brow);
        else if (bpart == 1) l3_build_groups<3, 3>(reg, tr, tcx, a_ptr, brow);
        else if (bpart == 2) l3_build_groups<6, 2>(reg, tr, tcx, a_ptr, brow);
        else l3_build_groups<8, 2>(reg, tr, tcx, a_ptr, brow);
      } else {
        l3_build_groups<0, 10>(reg, tr, tcx, a_ptr, brow);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_afull);
      named_bar_sync(2, kWorkers);              // the patch may be overwritten by the next build
    };

    int it = 0;
    int tile = blockIdx.x;
    L3TileIter ti, tp;                         // the tile being processed / the tile whose gradient patch is fetched next
    ti.init(blockIdx.x, gridDim.x, P.tiles_h, P.tiles_w);
    tp = ti;
    if (tile < P.total_tiles) {
      fetch_patch(tp);
      build(0);
      tp.advance();
      if (tile + static_cast<int>(gridDim.x) < P.total_tiles) fetch_patch(tp);
    }
    for (; tile < P.total_tiles; tile += gridDim.x, ++it, ti.advance()) {
      const int img = ti.img;
      const int i = ti.th * 8 + (row >> 4), j = ti.tw * 16 + (row & 15);
      const bool valid = (i < P.Hin) && (j < P.Win);
      const long long off0 = ((static_cast<long long>(img) * P.Hin + i) * P.Win + j) * NF;
      // ---- pass 1: g (TMEM), x and s (smem)  ->  t in place of x, g*s and x kept as packed bf16
      uint32_t keep[kDirect ? 1 : NH][16], xk[kDirect ? 1 : NH][16];
      // direct mode works in units of 16 columns (half a chunk) with the operands of the next unit in flight: at 576 threads
      // there are 96 registers per thread, and a whole 32-column chunk of accumulator + act + s + prefetch does not fit
      uint32_t xg[8], sg[8];
      auto load_xs = [&](int unit) {            // unit = 2 * h + half: this thread's pixel row, 16 columns of act and s
        const int col = (grp + (unit >> 1) * G) * 32 + (unit & 1) * 16;
        if (valid) {
          ld_global_nc_v8(P.x_ptr + off0 + col, xg);
          ld_global_nc_v8(P.s_ptr + off0 + col, sg);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) { xg[e] = 0u; sg[e] = 0x3F803F80u; }
        }
      };
      if constexpr (kDirect) {
        load_xs(0);                             // in flight while the accumulator is awaited
        mbar_wait(bar_tfull, it & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int unit = 0; unit < 2 * NH; ++unit) {
          const int c = grp + (unit >> 1) * G, half = unit & 1;
          uint32_t r[16];
          tmem_ld16(tmem_base + lane_off + c * 32 + half * 16, r);
          uint32_t xa[8], sa[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { xa[e] = xg[e]; sa[e] = sg[e]; }
          if (unit + 1 < 2 * NH) load_xs(unit + 1);
          tmem_ld_wait();
          uint8_t* xc = a2_ptr + (c >> 1) * 16384;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t a2w[4];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const int e = u * 4 + e4;
              const float v0 = __uint_as_float(r[2 * e]), v1 = __uint_as_float(r[2 * e + 1]);
              const float s0 = l3_bf16_lo(sa[e]), s1 = l3_bf16_hi(sa[e]);
              const float r0 = l3_rcp(s0), r1 = l3_rcp(s1);
              const float x0 = l3_bf16_lo(xa[e]) * r0, x1 = l3_bf16_hi(xa[e]) * r1;
              a2w[e4] = pack_bf16(v0 * x0 * r0, v1 * x1 * r1);
            }
            *reinterpret_cast<uint4*>(xc + sw128_offset(row, (c & 1) * 4 + half * 2 + u)) = make_uint4(a2w[0], a2w[1], a2w[2], a2w[3]);
          }
        }
      } else {
        mbar_wait(bar_sfull, it & 1);
        mbar_wait(bar_xfull, it & 1);
        mbar_wait(bar_tfull, it & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int c = grp + h * G;
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_off + c * 32, r);
          uint8_t* xc = a2_ptr + (c >> 1) * 16384;
          const uint8_t* sc = s_ptr + (c >> 1) * 16384;
          uint32_t sv[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t o = sw128_offset(row, (c & 1) * 4 + u);
            const uint4 xq = *reinterpret_cast<const uint4*>(xc + o);
            const uint4 sq = *reinterpret_cast<const uint4*>(sc + o);
            xk[h][4 * u] = xq.x; xk[h][4 * u + 1] = xq.y; xk[h][4 * u + 2] = xq.z; xk[h][4 * u + 3] = xq.w;
            sv[4 * u] = sq.x; sv[4 * u + 1] = sq.y; sv[4 * u + 2] = sq.z; sv[4 * u + 3] = sq.w;
          }
          if (!valid) {                                          // out-of-image rows are zero-filled by TMA: s = 1
#pragma unroll
            for (int e = 0; e < 16; ++e) sv[e] = 0x3F803F80u;
          }
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t a2w[4];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const int e = u * 4 + e4;
              const float v0 = __uint_as_float(r[2 * e]), v1 = __uint_as_float(r[2 * e + 1]);
              const uint32_t aw = xk[h][e];                        // saved act = x*s: x = act / s
              const uint32_t sw = sv[e];
              const float s0 = l3_bf16_lo(sw), s1 = l3_bf16_hi(sw);
              const float r0 = l3_rcp(s0), r1 = l3_rcp(s1);
              const float x0 = l3_bf16_lo(aw) * r0, x1 = l3_bf16_hi(aw) * r1;
              xk[h][e] = pack_bf16(x0, x1);
              keep[h][e] = pack_bf16(v0 * s0, v1 * s1);
              a2w[e4] = pack_bf16(v0 * x0 * r0, v1 * x1 * r1);
            }
            *reinterpret_cast<uint4*>(xc + sw128_offset(row, (c & 1) * 4 + u)) = make_uint4(a2w[0], a2w[1], a2w[2], a2w[3]);
          }
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(bar_a2);
      // ---- im2col of the next tile while the mixing GEMM runs; then fetch the patch of the tile after it
      const int next = tile + gridDim.x;
      if (next < P.total_tiles) {
        build(it + 1);
        tp.advance();
        if (next + static_cast<int>(gridDim.x) < P.total_tiles) fetch_patch(tp);
      }
      // ---- pass 2
      if constexpr (kDirect) {
        // stateless: g from the accumulator (still alive: the next gather GEMM waits for bar_nempty), x = act / s again (L2)
        load_xs(0);                             // in flight while the mixing GEMM finishes
        mbar_wait(bar_nfull, it & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int unit = 0; unit < 2 * NH; ++unit) {
          const int c = grp + (unit >> 1) * G, half = unit & 1;
          uint32_t m[16], g[16];
          tmem_ld16(tmem_base + kNormCol + lane_off + c * 32 + half * 16, m);
          tmem_ld16(tmem_base + lane_off + c * 32 + half * 16, g);
          uint32_t xa[8], sa[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { xa[e] = xg[e]; sa[e] = sg[e]; }
          if (unit + 1 < 2 * NH) load_xs(unit + 1);
          tmem_ld_wait();
          if (unit == 2 * NH - 1) {             // accumulator and mixing result are in registers: release both
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_nempty);
          }
          if (valid) {
            uint32_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float s0 = l3_bf16_lo(sa[e]), s1 = l3_bf16_hi(sa[e]);
              const float x0 = l3_bf16_lo(xa[e]) * l3_rcp(s0), x1 = l3_bf16_hi(xa[e]) * l3_rcp(s1);
              o[e] = pack_bf16(__uint_as_float(g[2 * e]) * s0 + x0 * __uint_as_float(m[2 * e]),
                               __uint_as_float(g[2 * e + 1]) * s1 + x1 * __uint_as_float(m[2 * e + 1]));
            }
            st_global_v8(P.out + off0 + c * 32 + half * 16, o);
          }
        }
        continue;
      }
      mbar_wait(bar_nfull, it & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const int c = grp + h * G;
        uint32_t m[32];
        tmem_ld32(tmem_base + kNormCol + lane_off + c * 32, m);
        tmem_ld_wait();
        if (h == NH - 1) {                  // the last chunk of the mixing result is in registers: release it
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_nempty);
        }
        if constexpr (PAIR) {
          // lanes 2k, 2k+1 = pixels (i, j), (i, j+1) with j even: afterwards the even lane holds the first 32 bytes of both
          // pixels' 64-byte pieces and the odd lane the second 32 bytes; one store instruction per pixel of the pair
          uint32_t o[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint32_t kw = keep[h][e];
            const uint32_t xw = xk[h][e];
            o[e] = pack_bf16(l3_bf16_lo(kw) + l3_bf16_lo(xw) * __uint_as_float(m[2 * e]),
                             l3_bf16_hi(kw) + l3_bf16_hi(xw) * __uint_as_float(m[2 * e + 1]));
          }
          const bool odd = (lane & 1) != 0;
          uint32_t oa[8], ob[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t got = __shfl_xor_sync(0xffffffffu, odd ? o[e] : o[8 + e], 1);
            oa[e] = odd ? got : o[e];        // pixel j:   own first half   | the even lane's second half
            ob[e] = odd ? o[8 + e] : got;    // pixel j+1: the odd lane's first half | own second half
          }
          const long long offe = off0 - (odd ? NF : 0) + c * 32 + (odd ? 16 : 0);
          const int je = j & ~1;
          if (i < P.Hin && je < P.Win) st_global_v8(P.out + offe, oa);
          if (i < P.Hin && je + 1 < P.Win) st_global_v8(P.out + offe + NF, ob);
        } else if (valid) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t o[8];
#pragma unroll
            for (int e8 = 0; e8 < 8; ++e8) {
              const int e = u * 8 + e8;
              const uint32_t kw = keep[h][e];
              const uint32_t xw = xk[h][e];
              o[e8] = pack_bf16(l3_bf16_lo(kw) + l3_bf16_lo(xw) * __uint_as_float(m[2 * e]),
                                l3_bf16_hi(kw) + l3_bf16_hi(xw) * __uint_as_float(m[2 * e + 1]));
            }
            st_global_v8(P.out + off0 + c * 32 + u * 16, o);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================================
// host
// =====================================================================================================
template <typename Params>
static int launch_pdl(void (*kernel)(const Params), int grid, int threads, int smem, cudaStream_t st, const Params& P) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(static_cast<unsigned>(threads), 1, 1);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  SGA_CUDA(cudaLaunchKernelEx(&cfg, kernel, P));
  g_launches += 1;
  return SGA_OK;
}

struct L3Tc {
  L3FwdParams F;
  L3DgradParams D;
  L3DgradFusedParams DF;
  int f_smem, f_grid, d_smem, d_grid, df_smem;
  bool fused_ok;
};

// w_scatter: bf16 [80][C] (row n = (a*5+b)*3+c);  w_gather: bf16 [10][C][8]
__global__ void k_l3_pack(const float* __restrict__ k3, int C, __nv_bfloat16* __restrict__ w_scatter,
                          __nv_bfloat16* __restrict__ w_gather) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 80 * C) return;
  const int n = e / C, ci = e % C;
  float v = 0.0f;
  if (n < 75) {
    const int tap = n / 3, c = n % 3;
    v = k3[(static_cast<long long>(tap) * C + ci) * 3 + c];
  }
  w_scatter[e] = __float2bfloat16_rn(v);
  w_gather[(static_cast<long long>(n >> 3) * C + ci) * 8 + (n & 7)] = __float2bfloat16_rn(v);
}

int l3_pack_weights(const float* k3, int C, __nv_bfloat16* w_scatter, __nv_bfloat16* w_gather, cudaStream_t st) {
  k_l3_pack<<<(80 * C + 255) / 256, 256, 0, st>>>(k3, C, w_scatter, w_gather);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

// gamma_bwd_op (optional): bf16 [C][C] K-major operand of the IGDN backward mixing (= gamma in its (in,out) layout)
int l3_tc_create(const __nv_bfloat16* act, const __nv_bfloat16* w_scatter, const __nv_bfloat16* w_gather,
                 const __nv_bfloat16* gamma_bwd_op, const __nv_bfloat16* x_saved, const __nv_bfloat16* s_saved, int N,
                 int Hin, int Win, int C, int H, int W, L3Tc** out) {
  SGA_CHECK(C % 64 == 0 && C <= 256, "tcgen05 layer-3 kernels need C %% 64 == 0 and <= 256 (got %d)", C);
  L3Tc* c = new L3Tc();
  memset(c, 0, sizeof(*c));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  {
    L3FwdParams& F = c->F;
    F.N = N; F.Hin = Hin; F.Win = Win; F.C = C; F.H = H; F.W = W;
    F.kchunks = C / 64;
    F.tiles_h = ceil_div(Hin, kL3IntH); F.tiles_w = ceil_div(Win, kL3IntW);
    F.total_tiles = N * F.tiles_h * F.tiles_w;
    const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                              static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(Win) * C * 2,
                                 static_cast<uint64_t>(Hin) * Win * C * 2};
    const uint32_t box[4] = {64, 16, 8, 1};
    int st = tma_encode_tiled(&F.amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, act, dims, strides, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != SGA_OK) { delete c; return st; }
    const uint64_t bdims[2] = {static_cast<uint64_t>(C), 80};
    const uint64_t bstr[1] = {static_cast<uint64_t>(C) * 2};
    const uint32_t bbox[2] = {64, 80};
    st = tma_encode_tiled(&F.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_scatter, bdims, bstr, bbox,
                          CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != SGA_OK) { delete c; return st; }
    const int f_fixed = 1024 + F.kchunks * kL3NPad * 128 + kL3Groups * 128 * kPsStride * 4 + 256;
    F.stages = (227 * 1024 - f_fixed) / 16384;
    if (F.stages > kL3MaxStages) F.stages = kL3MaxStages;
    if (const char* e = getenv("SGA_L3_STAGES")) { const int v = atoi(e); if (v >= 2 && v < F.stages) F.stages = v; }
    SGA_CHECK(F.stages >= 2, "not enough shared memory for the layer-3 pipeline");
    // measured (profiles/r2_schedule_ab.md): neither the L2 prefetch nor a deeper ring changes the kernel's time
    { const char* e = getenv("SGA_L3_PREFETCH"); F.prefetch = e ? atoi(e) : 0; }
    c->f_smem = f_fixed + F.stages * 16384;
    c->f_grid = F.total_tiles < sms ? F.total_tiles : sms;
  }
  {
    L3DgradParams& D = c->D;
    D.N = N; D.Hin = Hin; D.Win = Win; D.C = C; D.wpack = w_gather;
    D.tiles_h = ceil_div(Hin, 8); D.tiles_w = ceil_div(Win, 16);
    D.total_tiles = N * D.tiles_h * D.tiles_w;
    const int b_bytes = 10 * C * 16;
    const int reg_bytes = ((kRegH * kRegW * 8 + 127) / 128) * 128;
    c->d_smem = 1024 + ((b_bytes + 1023) / 1024) * 1024 + 2 * 20480 + 2 * reg_bytes + 256;
    c->d_grid = D.total_tiles < sms ? D.total_tiles : sms;
    L3DgradFusedParams& DF = c->DF;
    DF.N = N; DF.Hin = Hin; DF.Win = Win; DF.C = C; DF.wpack = w_gather;
    DF.tiles_h = D.tiles_h; DF.tiles_w = D.tiles_w; DF.total_tiles = D.total_tiles;
    { const char* e = getenv("SGA_L3_PREFETCH"); DF.prefetch = e ? (atoi(e) > 0 ? 1 : 0) : 0; }
    c->fused_ok = (C == 64 || C == 128 || C == 192 || C == 256) && gamma_bwd_op != nullptr && x_saved != nullptr &&
                  s_saved != nullptr;
    DF.x_ptr = x_saved; DF.s_ptr = s_saved;
    if (c->fused_ok) {
      const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                                static_cast<uint64_t>(N)};
      const uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(Win) * C * 2,
                                   static_cast<uint64_t>(Hin) * Win * C * 2};
      const uint32_t box[4] = {64, 16, 8, 1};
      int stx = tma_encode_tiled(&DF.xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x_saved, dims, strides, box,
                                 CU_TENSOR_MAP_SWIZZLE_128B);
      if (stx == SGA_OK)
        stx = tma_encode_tiled(&DF.smap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, s_saved, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
      if (stx != SGA_OK) { delete c; return stx; }
      const uint64_t gdims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(C)};
      const uint64_t gstr[1] = {static_cast<uint64_t>(C) * 2};
      const uint32_t gbox[2] = {64, static_cast<uint32_t>(C)};
      int st = tma_encode_tiled(&DF.gmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gamma_bwd_op, gdims, gstr, gbox,
                                CU_TENSOR_MAP_SWIZZLE_128B);
      if (st != SGA_OK) { delete c; return st; }
    }
    const int kc = C / 64;
    c->df_smem = 1024 + kc * C * 128 + 2 * kc * 16384 + ((b_bytes + 1023) / 1024) * 1024 + 20480 + reg_bytes + 256;
    if (C == 256)       // direct mode: gamma | shared weights / t buffer | im2col | patch
      c->df_smem = 1024 + kc * C * 128 + kc * 16384 + 20480 + reg_bytes + 256;
  }
  static bool attr_set = false;
  if (!attr_set) {
    SGA_CUDA(cudaFuncSetAttribute(k_l3_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SGA_CUDA(cudaFuncSetAttribute(k_l3_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SGA_CUDA(cudaFuncSetAttribute(k_l3_dgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SGA_CUDA(cudaFuncSetAttribute(k_l3_dgrad_igdn<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SGA_CUDA(cudaFuncSetAttribute(k_l3_dgrad_igdn<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SGA_CUDA(cudaFuncSetAttribute(k_l3_dgrad_igdn<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SGA_CUDA((cudaFuncSetAttribute(k_l3_dgrad_igdn<192, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
    SGA_CUDA(cudaFuncSetAttribute(k_l3_dgrad_igdn<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  *out = c;
  return SGA_OK;
}

void l3_tc_destroy(L3Tc* c) { delete c; }

int l3_tc_forward(const L3Tc* c, const float* bias_dev_host3, const float* x, float grad_scale, float* xt_out,
                  __nv_bfloat16* gx, double* sse, double* sse_r, cudaStream_t st) {
  if (c->F.total_tiles == 0) return SGA_OK;
  L3FwdParams F = c->F;
  F.bias[0] = bias_dev_host3[0]; F.bias[1] = bias_dev_host3[1]; F.bias[2] = bias_dev_host3[2];
  F.x = x; F.grad_scale = grad_scale; F.xt_out = xt_out; F.gx = gx; F.sse = sse; F.sse_r = sse_r;
  static const bool realloc = [] { const char* e = getenv("SGA_L3F_REALLOC"); return !(e && atoi(e) == 0); }();   // A/B switch
  if (realloc) SGA_TRY(launch_pdl(k_l3_fwd<true>, c->f_grid, l3f_threads(true), c->f_smem, st, F));
  else SGA_TRY(launch_pdl(k_l3_fwd<false>, c->f_grid, l3f_threads(false), c->f_smem, st, F));
  return SGA_OK;
}

int l3_tc_dgrad(const L3Tc* c, const __nv_bfloat16* gx, __nv_bfloat16* out, cudaStream_t st) {
  if (c->D.total_tiles == 0) return SGA_OK;
  L3DgradParams D = c->D;
  D.gx = gx; D.out = out;
  k_l3_dgrad<<<c->d_grid, kL3dThreads, c->d_smem, st>>>(D);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

bool l3_tc_can_fuse_igdn(const L3Tc* c) { return c->fused_ok; }

int l3_tc_dgrad_igdn(const L3Tc* c, const __nv_bfloat16* gx, __nv_bfloat16* out, cudaStream_t st) {
  SGA_CHECK(c->fused_ok, "layer-3 dgrad + IGDN backward fusion needs C in {64,128,192,256} and the gamma operand at create");
  if (c->DF.total_tiles == 0) return SGA_OK;
  const int C = c->DF.C;
  L3DgradFusedParams D = c->DF;
  D.gx = gx; D.out = out;
  const int threads = 64 + 128 * l3_groups(C);
  if (C == 256) SGA_TRY(launch_pdl(k_l3_dgrad_igdn<256>, c->d_grid, threads, c->df_smem, st, D));
  else if (C == 192) {
    static const bool pair = [] { const char* e = getenv("SGA_L3D_PAIR"); return !(e && atoi(e) == 0); }();   // A/B switch
    if (pair) SGA_TRY(launch_pdl(k_l3_dgrad_igdn<192>, c->d_grid, threads, c->df_smem, st, D));
    else SGA_TRY((launch_pdl(k_l3_dgrad_igdn<192, false>, c->d_grid, threads, c->df_smem, st, D)));
  }
  else if (C == 128) SGA_TRY(launch_pdl(k_l3_dgrad_igdn<128>, c->d_grid, threads, c->df_smem, st, D));
  else SGA_TRY(launch_pdl(k_l3_dgrad_igdn<64>, c->d_grid, threads, c->df_smem, st, D));
  return SGA_OK;
}

}  // namespace sga
