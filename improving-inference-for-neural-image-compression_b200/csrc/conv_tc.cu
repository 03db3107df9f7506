// conv_tc.cu -- bf16 implicit-GEMM (de)convolution on the sm_100a tensor cores, optionally fused with the inverse
// GDN that follows it in the synthesis transform (forward) or precedes it in the backward pass.
//
// Every (de)conv of the synthesis / hyper-synthesis transforms (nn_models.py:39-70, 140-170) and its dgrad is a
// sum over filter taps of  [128 pixels x Cin] * [Cin x Cout]  products (conv_plan.cu).  Per tap and 64-channel slice
//   A tile : TMA 4-D box {64 ch, TW, TH, 1} of the NHWC activation at the tap's spatial offset -- out-of-bounds
//            pixels/channels are zero-filled by the TMA unit, which implements "same_zeros" padding and ragged edges;
//            the box lands in shared memory as a K-major SWIZZLE_128B operand tile (128 rows x 128 B)
//   B tile : TMA 2-D box {64 ch, BLOCK_N} of the pre-packed K-major weights [tap][Cout][Cin]
//   MMA    : 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) accumulating in TMEM (fp32)
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, then G groups of 4 epilogue warps (one warp
// per TMEM lane quarter; the groups split the accumulator columns) -- the epilogue must finish a tile faster than the
// tensor pipe does (~6 us for a 128x192 tile of the widest layer), which 4 warps cannot do for the fused variants.
// Persistent CTAs (one per SM); the TMEM accumulator is double-buffered so the epilogue of tile i overlaps the MMAs
// of tile i+1.
// CTA pairs (template CG = 2, the default): the kernel is L2->SM bandwidth bound when every CTA streams its own A and B
// tiles (40 KB per 64-wide K chunk at N = 192, i.e. ~100 B/clk/SM at the tensor rate), so two CTAs of one TPC run as a
// cluster and issue tcgen05.mma.cta_group::2 (M = 256): each CTA loads the A tile of its own 128 pixels and only HALF of
// the weight tile, the tensor core reads the other half from the peer's shared memory.  Only the leader (cluster rank 0)
// issues MMAs; both producers signal the leader's "full" barrier, the MMA commits are multicast to both CTAs' barriers,
// and the peer's epilogue warps arrive on the leader's barriers through shared::cluster addresses.  Tiles are issued in waves of gridDim.x spatial tiles per output phase, so the 4 phases of a transposed
// conv re-read their (shared) input from L2 instead of HBM and every CTA gets the same mix of 9/6/6/4-tap tiles.
//
// Fused IGDN (template FUSE): the CxC channel mixing of tfc.GDN(inverse=True) (nn_models.py:51,55,59) runs as a second
// GEMM on the finished accumulator tile.
//   forward  (FUSE=1): x = acc + bias;  s = sqrt(beta + x^2 @ gamma);  out = x*s           (stores x, s, out as bf16)
//   backward (FUSE=2): g = acc;  t = g*x/s;  u = t @ gamma^T;  g_x = g*s + x*u             (stores g_x as bf16)
// Epilogue pass 1 reads the accumulator, writes the A operand of the mixing GEMM (x^2 or t, bf16, swizzled K-major) to
// shared memory and releases the accumulator; the MMA warp slots the mixing MMAs of tile i into the main loop of tile
// i+1 (two halves of N/2 columns into a third TMEM region; gamma streams through the regular B stages); epilogue pass 2
// finishes the point-wise math from registers (x resp. g*s are kept packed as bf16) and the mixing result in TMEM.
#include "kernels.h"
#include "tc_kernels.h"
#include "tc_common.cuh"
#include "conv_tc.cuh"
#include <stdlib.h>

namespace sga {

using namespace tc;

template <int NV>
__device__ __forceinline__ void epilogue_store(const TcParams& P, const uint32_t (&r)[NV], const float* s_bias, int ncol,
                                               long long off, bool valid) {
  if (!valid) return;
  float v[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    v[q] = __uint_as_float(r[q]) + s_bias[ncol + q];
    if (P.relu) v[q] = fmaxf(v[q], 0.0f);
  }
  if (P.mask) {
    const uint4* mp = reinterpret_cast<const uint4*>(P.mask + off);
#pragma unroll
    for (int q = 0; q < NV / 8; ++q) {
      const uint4 m = __ldg(mp + q);
      const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&mw[e]);
        if (!(__low2float(b2) > 0.0f)) v[q * 8 + e * 2] = 0.0f;
        if (!(__high2float(b2) > 0.0f)) v[q * 8 + e * 2 + 1] = 0.0f;
      }
    }
  }
  if (P.out_bf16) {
#pragma unroll
    for (int q = 0; q < NV / 16; ++q) {
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = pack_bf16(v[q * 16 + 2 * e], v[q * 16 + 2 * e + 1]);
      st_global_v8(P.out_bf16 + off + q * 16, w);
    }
  }
  if (P.out_f32) {
#pragma unroll
    for (int q = 0; q < NV / 8; ++q) {
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = __float_as_uint(v[q * 8 + e]);
      st_global_v8(P.out_f32 + off + q * 8, w);
    }
  }
}

// FUSE: 0 = plain conv epilogue, 1 = + IGDN forward, 2 = + IGDN backward.  NF = Cout of the fused variants.
// G = number of epilogue warp groups (4 warps each: one per TMEM lane quarter); the groups split the accumulator
// columns in 32-column chunks.  Fused variants use G = NF/64, so every epilogue warp owns exactly one chunk in each
// half of the tile (chunk grp of half 0, chunk grp of half 1) and keeps its point-wise state in registers.
// CG = CTAs per MMA (1, or 2 = CTA pair launched as a cluster of two).
template <int FUSE, int NF, int G, int CG>
__global__ void __launch_bounds__(64 + 128 * G, 1) k_conv_tc(const __grid_constant__ TcParams P) {
  static_assert(FUSE == 0 || G * 64 == NF, "fused variants: one 32-column chunk per warp and half");
  static_assert(CG == 1 || CG == 2, "one CTA or a CTA pair");
  // FUSE == 1 (two-pass forward fusion) is only instantiated for NF = 256, where the split-warp kernel of
  // conv_igdn_fwd_tc.cu does not fit in TMEM (2 x 256 accumulator columns leave no room for the mixing result)
  static_assert(FUSE != 1 || NF == 256 || NF == 192, "the forward fusion of NF <= 128 lives in conv_igdn_fwd_tc.cu");
  // accumulator buffers: two (the epilogue of tile i overlaps the MMAs of tile i+1) unless 2*NF + NF/2 > 512 TMEM
  // columns -- NF = 256: one accumulator + the N/2-column mixing result = 384 columns; the accumulator is released after
  // pass 1 (its content lives on in registers as packed bf16), so the tensor pipe only idles during pass 1
  constexpr int kNBuf = (FUSE != 0 && 2 * NF + NF / 2 > kTmemCols) ? 1 : 2;
  pdl_launch_dependents();
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0u);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int b_bytes = (P.block_n / CG) * kBlockK * 2;   // this CTA's share of the weight tile
  const uint32_t a_base = base;
  const uint32_t b_base = base + P.stages * kABytes;
  constexpr int kA2Bytes = (FUSE != 0) ? (NF / 64) * kABytes : 0;
  const uint32_t a2_base = b_base + P.stages * b_bytes;            // fused: A operand of the mixing GEMM
  uint8_t* a2_ptr = smem + P.stages * (kABytes + b_bytes);
  uint8_t* tail = a2_ptr + kA2Bytes;
  float* s_bias = reinterpret_cast<float*>(tail);                  // 512 floats: bias (or beta for the fused fwd: +256)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 512 * sizeof(float));
  // bars: full[8] | empty[8] | tmem_full[2] | tmem_empty[2] | a2_full | norm_full | norm_empty
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t bar_a2 = bar_tempty + 16;
  const uint32_t bar_nfull = bar_a2 + 8;
  const uint32_t bar_nempty = bar_nfull + 8;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 7);
  // barriers the peer CTA of a pair arrives on live in the leader: shared::cluster addresses (CG = 1: the local ones)
  const uint32_t lead_full = (CG == 2) ? mapa_shared(bar_full, 0) : bar_full;
  const uint32_t lead_tempty = (CG == 2) ? mapa_shared(bar_tempty, 0) : bar_tempty;
  const uint32_t lead_a2 = (CG == 2) ? mapa_shared(bar_a2, 0) : bar_a2;
  const uint32_t lead_nempty = (CG == 2) ? mapa_shared(bar_nempty, 0) : bar_nempty;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kAccStride = (FUSE != 0) ? NF : 256;               // TMEM columns between the two accumulators
  constexpr int kNormCol = kNBuf * NF;                             // fused: mixing result (N/2 columns)
  constexpr int kNHalf = NF / 2;

  for (int i = threadIdx.x; i < P.Nn && i < 256; i += blockDim.x) {
    s_bias[i] = P.bias ? P.bias[i] : 0.0f;
  }
  if (FUSE == 0)
    for (int i = 256 + threadIdx.x; i < P.Nn && i < 512; i += blockDim.x) s_bias[i] = P.bias ? P.bias[i] : 0.0f;
  if (FUSE == 1)
    for (int i = threadIdx.x; i < NF; i += blockDim.x) s_bias[256 + i] = P.beta[i];

  if (warp == 0 && lane == 0) {
    for (int v = 0; v < kMaxViews; ++v) tma_prefetch_desc(&P.amap[v]);
    tma_prefetch_desc(&P.bmap);
    if (FUSE != 0) tma_prefetch_desc(&P.gmap);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 4 * G * CG);
    }
    mbar_init(bar_a2, 4 * G * CG);
    mbar_init(bar_nfull, 1);
    mbar_init(bar_nempty, 4 * G * CG);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_cg2(smem_u32(s_tmem), kTmemCols); tmem_relinquish_cg2(); }
    else { tmem_alloc(smem_u32(s_tmem), kTmemCols); tmem_relinquish(); }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  pdl_grid_dependency_wait();        // everything above touches only constants (bias) and on-chip state
  const uint32_t tmem_base = *s_tmem;
  const int nwaves = num_waves<CG>(P);
  // wait on a barrier the peer CTA arrives on / arrive on a leader barrier
  auto wait_x = [](uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); };
  auto arrive_lead = [](uint32_t local_bar, uint32_t lead_bar) { if (CG == 2) mbar_arrive_cluster(lead_bar); else mbar_arrive(local_bar); };
  // SGA_DBG & 8: per-role wait-cycle accounting, printed by the first and one middle CTA (development aid)
  // wait-cycle accounting (clock64 / printf) is a development aid: compiled in only with -DSGA_DBG (build.py --dbg),
  // otherwise dbg_t is a compile-time false and every instrumented statement reduces to the statement itself
#ifdef SGA_DBG
  const bool dbg_t = (P.dbg & 8) != 0;
#else
  constexpr bool dbg_t = false;
#endif
  long long dbg_w[4] = {0, 0, 0, 0};
  const long long dbg_t0 = dbg_t ? clock64() : 0;
  unsigned long long dbg_g0 = 0;
  if (dbg_t) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g0));
#define SGA_TIMED(slot, stmt)                                  \
  do {                                                         \
    if (dbg_t) { const long long _c = clock64(); stmt; dbg_w[slot] += clock64() - _c; } else { stmt; } \
  } while (0)

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (warp-uniform control flow keeps addresses / coordinates in uniform registers);
    // one elected lane issues the copies.  A single-lane loop costs ~350 clk of dependent-issue latency per K chunk,
    // which is as long as the 4 MMAs that consume the chunk.
    {
      int stage = 0;
      uint32_t phase = 0;
      auto mix_loads = [&](int h) {   // gamma rows [h*N/2, +N/2), all K chunks, through the regular B stages
        for (int kc = 0; kc < NF / 64; ++kc) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, kNHalf * 128);
            if (CG == 2)
              tma_load_2d_cg2(b_base + stage * b_bytes, &P.gmap, lead_full + 8 * stage, kc * kBlockK,
                              h * kNHalf + static_cast<int>(rank) * (kNHalf / 2));
            else
              tma_load_2d(b_base + stage * b_bytes, &P.gmap, bar_full + 8 * stage, kc * kBlockK, h * kNHalf);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
      };
      bool pending = false;           // a finished tile whose mixing GEMM has not been issued yet
      const uint32_t stage_tx = static_cast<uint32_t>(CG * (kABytes + b_bytes));
      const int b_row_off = static_cast<int>(rank) * (P.block_n / 2);
      for (int w = 0; w < nwaves; ++w) {
        TileCoord tc_;
        if (!decode_slot<CG>(P, w, rank, tc_)) continue;
        const TcJob& J = P.job[tc_.job];
        const int nk = J.ntaps * P.kchunks;
        const int ins0 = (P.mix_at0 * nk) / 8, ins1 = (P.mix_at1 * nk) / 8 < nk ? (P.mix_at1 * nk) / 8 : nk - 1;
        int k = 0;
        for (int t = 0; t < J.ntaps; ++t) {
          const Tap tap = P.taps[J.tap_begin + t];
          const CUtensorMap* am = &P.amap[tap.view];
          const int ax = tc_.w0 + tap.dx, ay = tc_.h0 + tap.dy, brow = tap.widx * P.Nn + tc_.n0 + (CG == 2 ? b_row_off : 0);
          for (int kc = 0; kc < P.kchunks; ++kc, ++k) {
            if (FUSE != 0 && pending) {
              if (k == ins0) mix_loads(0);
              if (k == ins1) mix_loads(1);
            }
            SGA_TIMED(0, mbar_wait(bar_empty + 8 * stage, phase ^ 1u));
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, stage_tx);
              if (CG == 2) {
                tma_load_4d_cg2(a_base + stage * kABytes, am, lead_full + 8 * stage, kc * kBlockK, ax, ay, tc_.img);
                tma_load_2d_cg2(b_base + stage * b_bytes, &P.bmap, lead_full + 8 * stage, kc * kBlockK, brow);
              } else {
                tma_load_4d(a_base + stage * kABytes, am, bar_full + 8 * stage, kc * kBlockK, ax, ay, tc_.img);
                tma_load_2d(b_base + stage * b_bytes, &P.bmap, bar_full + 8 * stage, kc * kBlockK, brow);
              }
            }
            if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          }
        }
        pending = true;
      }
      if (FUSE != 0 && pending) { mix_loads(0); mix_loads(1); }
      if (dbg_t && lane == 0 && (blockIdx.x < 2 || blockIdx.x == 74))
        printf("dbg cta %d producer: total %lld wait_empty %lld\n", blockIdx.x, clock64() - dbg_t0, dbg_w[0]);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop, one elected lane issues the MMAs and commits (tcgen05.commit tracks the issuing thread's MMAs:
    // elect.sync on a converged warp always picks the same lane).
    if (leader) {
      const uint32_t idesc = instr_desc_bf16(kBlockM * CG, P.block_n);
      const uint32_t idesc_mix = instr_desc_bf16(kBlockM * CG, (FUSE != 0) ? kNHalf : 16);
      auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
        if (CG == 2) umma_bf16_cg2(d, ad, bd, id, acc); else umma_bf16(d, ad, bd, id, acc);
      };
      auto commit = [](uint32_t bar) { if (CG == 2) umma_commit_cg2_both(bar); else umma_commit(bar); };
      // descriptors: only the 14-bit start-address field changes (stage, K = 16 slice), added to a prebuilt base
      const uint64_t adesc0 = smem_desc_sw128(a_base), bdesc0 = smem_desc_sw128(b_base), a2desc0 = smem_desc_sw128(a2_base);
      const uint32_t a_step = kABytes >> 4, b_step = static_cast<uint32_t>(b_bytes) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;        // tiles issued by this CTA
      int halves = 0;    // mixing halves issued
      auto mix_mma = [&](int h) {
        if (h == 0) SGA_TIMED(2, wait_x(bar_a2, (it - 1) & 1));      // A operand of tile it-1 written (pass 1)
        SGA_TIMED(3, wait_x(bar_nempty, (halves & 1) ^ 1u));         // previous half drained by the epilogue
        tcgen05_fence_after();
        for (int kc = 0; kc < NF / 64; ++kc) {
          mbar_wait(bar_full + 8 * stage, phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t ad = a2desc0 + static_cast<uint64_t>(kc * (kABytes >> 4));
            const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * b_step);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma(tmem_base + kNormCol, ad + 2 * kk, bd + 2 * kk, idesc_mix, (kc | kk) != 0 ? 1u : 0u);
            commit(bar_empty + 8 * stage);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) commit(bar_nfull);
        ++halves;
      };
      for (int w = 0; w < nwaves; ++w) {
        TileCoord tc_;
        if (!decode_slot<CG>(P, w, rank, tc_)) continue;
        const int nk = P.job[tc_.job].ntaps * P.kchunks;
        const int ins0 = (P.mix_at0 * nk) / 8, ins1 = (P.mix_at1 * nk) / 8 < nk ? (P.mix_at1 * nk) / 8 : nk - 1;
        const int buf = it % kNBuf;
        SGA_TIMED(1, wait_x(bar_tempty + 8 * buf, ((it / kNBuf) & 1) ^ 1u));
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kAccStride;
        for (int k = 0; k < nk; ++k) {
          if (FUSE != 0 && it > 0) {
            if (k == ins0) mix_mma(0);
            if (k == ins1) mix_mma(1);
          }
          SGA_TIMED(0, mbar_wait(bar_full + 8 * stage, phase));
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * a_step);
            const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * b_step);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) mma(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            commit(bar_empty + 8 * stage);        // frees the smem stage (in both CTAs) once these MMAs have read it
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) commit(bar_tfull + 8 * buf);            // accumulator complete
        ++it;
      }
      if (FUSE != 0 && it > 0) { mix_mma(0); mix_mma(1); }
      if (dbg_t && lane == 0 && (blockIdx.x < 2 || blockIdx.x == 74)) {
        unsigned long long g1 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        printf("dbg cta %d mma: total %lld (%.0f MHz) wait_full %lld wait_tempty %lld wait_a2 %lld wait_nempty %lld tiles %d\n",
               blockIdx.x, clock64() - dbg_t0, 1000.0 * static_cast<double>(clock64() - dbg_t0) / static_cast<double>(g1 - dbg_g0),
               dbg_w[0], dbg_w[1], dbg_w[2], dbg_w[3], it);
      }
    }
  } else {
    // ===================== epilogue (warps 2 .. 2+4G) =====================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;             // column group
    const int row = q * 32 + lane;               // tile row = pixel
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    int it = 0;
    int halves = 0;
    for (int w = 0; w < nwaves; ++w) {
      TileCoord tc_;
      if (!decode_slot<CG>(P, w, rank, tc_)) continue;
      const TcJob& J = P.job[tc_.job];
      const int buf = it % kNBuf;
      const int i = tc_.h0 + row / P.TW, j = tc_.w0 + row % P.TW;
      const bool valid = (i < J.Ho) && (j < J.Wo) && (tc_.img < P.N);
      const long long off0 = J.out_off + tc_.img * P.oN + i * J.oH + j * J.oW + tc_.n0;
      // fused backward: the saved x / s of this warp's two chunks are fetched before the accumulator is awaited
      uint32_t xin[2][16], sin[2][16];
      if (FUSE == 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = grp + h * G;
          if (valid) {
            ld_global_nc_v8(P.x_in + off0 + c * 32, &xin[h][0]);
            ld_global_nc_v8(P.x_in + off0 + c * 32 + 16, &xin[h][8]);
            ld_global_nc_v8(P.s_in + off0 + c * 32, &sin[h][0]);
            ld_global_nc_v8(P.s_in + off0 + c * 32 + 16, &sin[h][8]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) { xin[h][e] = 0u; sin[h][e] = 0x3F803F80u; }
          }
        }
      }
      SGA_TIMED(0, mbar_wait(bar_tfull + 8 * buf, (it / kNBuf) & 1));
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + buf * kAccStride + lane_off;
      if (FUSE == 0) {
        const int nfull = P.block_n >> 5;
        for (int c = grp; c < nfull; c += G) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          epilogue_store<32>(P, r, s_bias, tc_.n0 + c * 32, off0 + c * 32, valid);
        }
        if ((P.block_n & 31) != 0 && (nfull % G) == grp) {   // 16-column tail (block_n % 32 == 16)
          uint32_t r[16];
          tmem_ld16(taddr + nfull * 32, r);
          tmem_ld_wait();
          epilogue_store<16>(P, r, s_bias, tc_.n0 + nfull * 32, off0 + nfull * 32, valid);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) arrive_lead(bar_tempty + 8 * buf, lead_tempty + 8 * buf);
      } else {
        // ---------- pass 1: accumulator g, saved act / s -> g*s kept in registers as packed bf16, mixing operand -> smem
        uint32_t keep[2][16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = grp + h * G;               // 32-column chunk of the tile owned by this warp in half h
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          uint32_t a2w[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float v0 = __uint_as_float(r[2 * e]), v1 = __uint_as_float(r[2 * e + 1]);
            if (FUSE == 1) {
              // forward: x = acc + bias kept as packed bf16; the mixing sees the square of that stored x
              v0 += s_bias[c * 32 + 2 * e];
              v1 += s_bias[c * 32 + 2 * e + 1];
              const uint32_t xb = pack_bf16(v0, v1);
              keep[h][e] = xb;
              const float x0 = bf16_lo(xb), x1 = bf16_hi(xb);
              a2w[e] = pack_bf16(x0 * x0, x1 * x1);
              continue;
            }
            // the forward stores act = x*s and s: x = act / s
            const uint32_t aw = xin[h][e];
            const uint32_t sw = sin[h][e];
            const float s0 = bf16_lo(sw), s1 = bf16_hi(sw);
            const float r0 = fast_rcp(s0), r1 = fast_rcp(s1);
            const float x0 = bf16_lo(aw) * r0, x1 = bf16_hi(aw) * r1;
            xin[h][e] = pack_bf16(x0, x1);
            keep[h][e] = pack_bf16(v0 * s0, v1 * s1);
            a2w[e] = pack_bf16(v0 * x0 * r0, v1 * x1 * r1);
          }
          // 32 columns = 4 x 16-byte chunks of K-chunk (c/2), chunk index ((c&1)*4 + u) within the 128-byte row
          uint8_t* a2c = a2_ptr + (c >> 1) * kABytes;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(a2c + sw128_offset(row, (c & 1) * 4 + u)) =
                make_uint4(a2w[4 * u], a2w[4 * u + 1], a2w[4 * u + 2], a2w[4 * u + 3]);
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          arrive_lead(bar_a2, lead_a2);
          arrive_lead(bar_tempty + 8 * buf, lead_tempty + 8 * buf);
        }
        // ---------- pass 2: two halves of the mixing result; this warp owns chunk `grp` of each half
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = grp + h * G;
          SGA_TIMED(1, mbar_wait(bar_nfull, halves & 1));
          tcgen05_fence_after();
          uint32_t r[32];
          tmem_ld32(tmem_base + kNormCol + lane_off + grp * 32, r);
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) arrive_lead(bar_nempty, lead_nempty);   // the mixing result is in registers: release the TMEM region
          ++halves;
          uint32_t o[16];
          if (FUSE == 1) {
            // s = sqrt(beta + x^2 @ gamma) stored as bf16; act = x * (stored s), so that the backward's x = act / s inverts it
            uint32_t so[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float s0 = fast_sqrt(__uint_as_float(r[2 * e]) + s_bias[256 + c * 32 + 2 * e]);
              const float s1 = fast_sqrt(__uint_as_float(r[2 * e + 1]) + s_bias[256 + c * 32 + 2 * e + 1]);
              so[e] = pack_bf16(s0, s1);
              const uint32_t kw = keep[h][e];
              o[e] = pack_bf16(bf16_lo(kw) * bf16_lo(so[e]), bf16_hi(kw) * bf16_hi(so[e]));
            }
            if (valid) {
              st_global_v8(P.s_out + off0 + c * 32, &so[0]);
              st_global_v8(P.s_out + off0 + c * 32 + 16, &so[8]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float m0 = __uint_as_float(r[2 * e]), m1 = __uint_as_float(r[2 * e + 1]);
              const uint32_t kw = keep[h][e];
              const uint32_t xw = xin[h][e];
              o[e] = pack_bf16(bf16_lo(kw) + bf16_lo(xw) * m0, bf16_hi(kw) + bf16_hi(xw) * m1);
            }
          }
          if (valid) {
            st_global_v8(P.out_bf16 + off0 + c * 32, &o[0]);
            st_global_v8(P.out_bf16 + off0 + c * 32 + 16, &o[8]);
          }
        }
      }
      ++it;
    }
    if (dbg_t && threadIdx.x == 64 && (blockIdx.x < 2 || blockIdx.x == 74))
      printf("dbg cta %d epilogue: total %lld wait_tfull %lld wait_nfull %lld tiles %d\n", blockIdx.x, clock64() - dbg_t0,
             dbg_w[0], dbg_w[1], it);
  }

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer's shared memory / barriers are in use until here
  if (warp == 1) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int tma_encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SGA_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUresult r = fn(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,.. box %u,%u,..)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return SGA_ERR_CUDA;
  }
  return SGA_OK;
}


static int pick_block_n(int Nn) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (Nn % bn == 0) return bn;
  return 0;
}

typedef void (*TcKernel)(const TcParams);
template <int CG>
static TcKernel select_kernel_cg(int fuse, int nf) {
  if (fuse == 0) return k_conv_tc<0, 64, kPlainGroups, CG>;
  if (fuse == 1) return nf == 256 ? k_conv_tc<1, 256, 4, CG> : k_conv_tc<1, 192, 3, CG>;   // two-pass forward fusion
  if (nf == 256) return k_conv_tc<2, 256, 4, CG>;
  return nf == 192 ? k_conv_tc<2, 192, 3, CG> : (nf == 128 ? k_conv_tc<2, 128, 2, CG> : k_conv_tc<2, 64, 1, CG>);
}
static TcKernel select_kernel(int fuse, int nf, int cg) {
  return cg == 2 ? select_kernel_cg<2>(fuse, nf) : select_kernel_cg<1>(fuse, nf);
}
static int kernel_threads(int fuse, int nf) { return 64 + 128 * (fuse == 0 ? kPlainGroups : nf / 64); }

bool pdl_enabled() {
  static const bool on = getenv("SGA_NO_PDL") == nullptr;
  return on;
}

// one launch path for both modes: CTA pairs need the cluster dimension attribute
static int launch_tc(const TcConv* c, const TcParams& P, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(c->grid), 1, 1);
  cfg.blockDim = dim3(static_cast<unsigned>(kernel_threads(c->fuse, P.Nn)), 1, 1);
  cfg.dynamicSmemBytes = static_cast<size_t>(c->smem_bytes);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(c->cg);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_* in tc_common.cuh
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  SGA_CUDA(cudaLaunchKernelEx(&cfg, select_kernel(c->fuse, P.Nn, c->cg), P));
  g_launches += 1;
  return SGA_OK;
}

bool tc_conv_can_fuse(int cout) { return cout == 64 || cout == 128 || cout == 192 || cout == 256; }
// NF = 256 does not fit the split-warp forward kernel (TMEM): its forward fusion is the two-pass epilogue of k_conv_tc
// (SGA_FWD_TWO_PASS=1 selects it for NF = 192 as well: A/B switch against the split-warp kernel)
static bool fwd_two_pass(int cout) {
  static const bool force = getenv("SGA_FWD_TWO_PASS") != nullptr;
  return cout == 256 || (force && cout == 192);
}

int tc_conv_create(const ConvPlan& plan, const __nv_bfloat16* x, const __nv_bfloat16* w_kmajor, int fuse,
                   const __nv_bfloat16* mix_op, TcConv** out) {
  const ConvGeom& g = plan.g;
  SGA_CHECK(g.Cin % 8 == 0, "tcgen05 conv needs Cin %% 8 == 0 (got %d)", g.Cin);
  SGA_CHECK(g.Cout % 16 == 0 && g.Cout <= 512, "tcgen05 conv needs Cout %% 16 == 0 and <= 512 (got %d)", g.Cout);
  SGA_CHECK(fuse == 0 || (tc_conv_can_fuse(g.Cout) && mix_op != nullptr), "IGDN fusion needs Cout in {64,128,192,256}");
  for (int jn = 1; jn < plan.njobs; ++jn)
    SGA_CHECK(plan.job[jn].Ho == plan.job[0].Ho && plan.job[jn].Wo == plan.job[0].Wo, "jobs of a plan must share a grid");
  TcConv* c = new TcConv();
  memset(&c->P, 0, sizeof(c->P));
  c->fuse = fuse;
  c->cg = (getenv("SGA_NO_PAIR") == nullptr) ? 2 : 1;
  c->plan = plan;
  TcParams& P = c->P;
  P.N = g.N; P.K = g.Cin; P.Nn = g.Cout;
  P.kchunks = ceil_div(g.Cin, kBlockK);
  P.njobs = plan.njobs;
  {
    const char* e0 = getenv("SGA_MIX_AT0"); const char* e1 = getenv("SGA_MIX_AT1");
    P.mix_at0 = e0 ? atoi(e0) : 6; P.mix_at1 = e1 ? atoi(e1) : 8;
    // one accumulator buffer (NF = 256): the epilogue warps must finish pass 2 of tile i-1 before they can drain tile i,
    // and tile i+1 cannot start before that -- issue the mixing GEMM of tile i-1 at the very start of tile i's main loop
    if (fuse != 0 && g.Cout == 256 && !e0 && !e1) { P.mix_at0 = 0; P.mix_at1 = 1; }
    const char* ed = getenv("SGA_DBG");
    P.dbg = ed ? atoi(ed) : 0;
  }
  P.oN = static_cast<long long>(g.Hout) * g.Wout * g.Cout;
  // spatial tile: 128 pixels = TH x TW
  const int Wo0 = plan.job[0].Wo;
  P.TW = (Wo0 >= 16) ? 16 : 8;
  // experimental (SGA_HALO=2): halo-tile A reuse in the fused forward kernel; SGA_HALO=1 sets the descriptor base_offset
  // from the window start and is WRONG on this hardware (the operand swizzle follows absolute address bits) -- kept only
  // as the record of that experiment
  if (const char* eh = getenv("SGA_HALO")) P.halo = (fuse == 1) ? atoi(eh) : 0;
  if (P.halo) P.TW = 8;                               // one 8-row operand group per tile row
  P.TH = kBlockM / P.TW;
  P.tiles_h = ceil_div(plan.job[0].Ho, P.TH);
  P.tiles_w = ceil_div(plan.job[0].Wo, P.TW);
  P.n_sp = g.N * P.tiles_h * P.tiles_w;
  P.n_sp_pairs = (P.n_sp + 1) / 2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ctas = (c->cg == 2) ? sms / 2 : sms;                 // CTAs (or CTA pairs) that can be resident
  const int sp_units = (c->cg == 2) ? P.n_sp_pairs : P.n_sp;
  // N tile: the whole of Cout when it fits an MMA; small problems of the un-fused kernel take narrower tiles (down to
  // 64) until at least half of the CTAs have a unit -- they are latency bound, not tensor bound
  P.block_n = pick_block_n(g.Cout);
  if (fuse == 0 && getenv("SGA_NO_SMALL") == nullptr) {
    for (int bn = P.block_n; bn >= 64; bn -= 16) {
      if (g.Cout % bn != 0 || (bn / c->cg) % 8 != 0) continue;
      P.block_n = bn;
      if (sp_units * (g.Cout / bn) * plan.njobs * 2 >= ctas) break;
    }
  }
  P.n_tiles_n = g.Cout / P.block_n;
  P.n_spatial = P.n_sp * P.n_tiles_n;
  P.spread = (sp_units * P.n_tiles_n < ctas && getenv("SGA_NO_SMALL") == nullptr) ? 1 : 0;
  const int b_bytes = (P.block_n / c->cg) * kBlockK * 2;
  const int a2_bytes = fuse ? (g.Cout / 64) * kABytes : 0;
  const bool split_fwd = (fuse == 1) && !fwd_two_pass(g.Cout);
  const int fixed = split_fwd ? igdn_fwd_smem_fixed(g.Cout, c->cg) : 1024 + a2_bytes + 512 * 4 + 256;
  int stages = (227 * 1024 - fixed) / (kABytes + b_bytes);
  P.a_stages = 0;
  if (P.halo) {                                       // halo ring (A) + weight ring (B)
    P.a_stages = 2;
    stages = (227 * 1024 - fixed - P.a_stages * kHaloBytes) / b_bytes;
  }
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* es = getenv("SGA_DBG_STAGES")) { const int v = atoi(es); if (v >= 2 && v < stages) stages = v; }
  SGA_CHECK(stages >= 2, "not enough shared memory for the pipeline");
  P.stages = stages;
  c->smem_bytes = P.halo ? fixed + P.a_stages * kHaloBytes + stages * b_bytes : fixed + stages * (kABytes + b_bytes);
  int ntap = 0, ngroup = 0;
  for (int jn = 0; jn < plan.njobs; ++jn) {
    const ConvJobPlan& jp = plan.job[jn];
    TcJob& J = P.job[jn];
    J.tap_begin = ntap; J.ntaps = jp.ntaps;
    if (P.halo) {
      // group the taps by input view; the halo tile of a group starts at the smallest (dy, dx) of its taps
      J.group_begin = ngroup; J.ngroups = 0;
      int nh = ntap;
      for (int v = 0; v < jp.nviews; ++v) {
        int ox = 1 << 30, oy = 1 << 30, mx = -(1 << 30), my = -(1 << 30), cnt = 0;
        for (int t = 0; t < jp.ntaps; ++t) {
          if (jp.taps[t].view != v) continue;
          ++cnt;
          ox = jp.taps[t].dx < ox ? jp.taps[t].dx : ox; oy = jp.taps[t].dy < oy ? jp.taps[t].dy : oy;
          mx = jp.taps[t].dx > mx ? jp.taps[t].dx : mx; my = jp.taps[t].dy > my ? jp.taps[t].dy : my;
        }
        if (cnt == 0) continue;
        if (mx - ox + 8 > kHaloPitch || my - oy + 16 > kHaloRows || ngroup >= kMaxGroups) {
          delete c;
          set_error("tcgen05 conv: taps of view %d span more than the halo tile holds", v);
          return SGA_ERR_INVALID;
        }
        TcGroup& Gp = P.groups[ngroup++];
        Gp.view = v; Gp.ox = ox; Gp.oy = oy; Gp.tap_begin = nh; Gp.ntaps = cnt;
        for (int t = 0; t < jp.ntaps; ++t) {
          if (jp.taps[t].view != v) continue;
          TcHaloTap& T = P.htaps[nh++];
          T.widx = jp.taps[t].widx;
          const int ry = jp.taps[t].dy - oy, rx = jp.taps[t].dx - ox;
          T.a_off16 = static_cast<uint32_t>((ry * kHaloPitch + rx) * 128) >> 4;
          T.boff = (P.halo == 1) ? static_cast<uint32_t>(rx & 7) : 0u;
        }
        ++J.ngroups;
      }
    }
    for (int t = 0; t < jp.ntaps; ++t) P.taps[ntap++] = jp.taps[t];
    J.Ho = jp.Ho; J.Wo = jp.Wo;
    J.out_off = (jp.out_phase_stride == 2) ? (static_cast<long long>(jp.out_py) * g.Wout + jp.out_px) * g.Cout : 0;
    J.oH = static_cast<long long>(jp.out_phase_stride) * g.Wout * g.Cout;
    J.oW = static_cast<long long>(jp.out_phase_stride) * g.Cout;
  }
  // input views (identical for every job of a plan)
  const ConvJobPlan& j0 = plan.job[0];
  for (int v = 0; v < kMaxViews; ++v) {
    const int vv = (v < j0.nviews) ? v : 0;
    const int py = (j0.in_phase_stride == 2) ? j0.view_py[vv] : 0, px = (j0.in_phase_stride == 2) ? j0.view_px[vv] : 0;
    const int s = j0.in_phase_stride;
    const uint64_t dims[4] = {static_cast<uint64_t>(g.Cin), static_cast<uint64_t>((g.Win - px + s - 1) / s),
                              static_cast<uint64_t>((g.Hin - py + s - 1) / s), static_cast<uint64_t>(g.N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(s) * g.Cin * 2, static_cast<uint64_t>(s) * g.Win * g.Cin * 2,
                                 static_cast<uint64_t>(g.Hin) * g.Win * g.Cin * 2};
    const uint32_t box[4] = {kBlockK, static_cast<uint32_t>(P.halo ? kHaloPitch : P.TW),
                             static_cast<uint32_t>(P.halo ? kHaloRows : P.TH), 1};
    const __nv_bfloat16* basep = x + (static_cast<long long>(py) * g.Win + px) * g.Cin;
    int st = tma_encode_tiled(&P.amap[v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, basep, dims, strides, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != SGA_OK) { delete c; return st; }
  }
  {
    const int taps = g.kh * g.kw;
    const uint64_t dims[2] = {static_cast<uint64_t>(g.Cin), static_cast<uint64_t>(taps) * g.Cout};
    const uint64_t strides[1] = {static_cast<uint64_t>(g.Cin) * 2};
    const uint32_t box[2] = {kBlockK, static_cast<uint32_t>(P.block_n / c->cg)};
    int st = tma_encode_tiled(&P.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_kmajor, dims, strides, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != SGA_OK) { delete c; return st; }
  }
  if (fuse) {
    const uint64_t dims[2] = {static_cast<uint64_t>(g.Cout), static_cast<uint64_t>(g.Cout)};
    const uint64_t strides[1] = {static_cast<uint64_t>(g.Cout) * 2};
    // backward fusion: halves of the result (N/2 rows per MMA); forward fusion: 64-column parts
    const uint32_t box[2] = {kBlockK, static_cast<uint32_t>((split_fwd ? 64 : g.Cout / 2) / c->cg)};
    int st = tma_encode_tiled(&P.gmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, mix_op, dims, strides, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != SGA_OK) { delete c; return st; }
  }
  {
    const int units = sp_units * P.n_tiles_n * (P.spread ? P.njobs : 1);
    c->grid = c->cg * (units < ctas ? units : ctas);
  }
  if (split_fwd) SGA_TRY(igdn_fwd_set_attr(g.Cout, c->cg));
  else SGA_CUDA(cudaFuncSetAttribute(select_kernel(fuse, g.Cout, c->cg), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  *out = c;
  return SGA_OK;
}

void tc_conv_destroy(TcConv* c) { delete c; }

int tc_conv_launch(const TcConv* c, const float* bias, int relu, const __nv_bfloat16* relu_mask, __nv_bfloat16* out_bf16,
                   float* out_f32, cudaStream_t st) {
  SGA_CHECK(c->fuse == 0, "fused conv must be launched with tc_conv_launch_igdn_*");
  if (c->P.n_spatial == 0) return SGA_OK;
  TcParams P = c->P;
  P.bias = bias; P.relu = relu; P.mask = relu_mask; P.out_bf16 = out_bf16; P.out_f32 = out_f32;
  return launch_tc(c, P, st);
}

int tc_conv_launch_igdn_fwd(TcConv* c, const float* bias, const float* beta, __nv_bfloat16* s_out, __nv_bfloat16* act_out,
                            cudaStream_t st) {
  SGA_CHECK(c->fuse == 1, "conv was not created with forward IGDN fusion");
  if (c->P.n_spatial == 0) return SGA_OK;
  TcParams P = c->P;
  P.bias = bias; P.beta = beta; P.s_out = s_out; P.out_bf16 = act_out;
  if (fwd_two_pass(P.Nn)) return launch_tc(c, P, st);
  return launch_igdn_fwd_kernel(c, P, st);
}

int tc_conv_launch_igdn_bwd(const TcConv* c, const __nv_bfloat16* act_saved, const __nv_bfloat16* s_saved,
                            __nv_bfloat16* gx_out, cudaStream_t st) {
  SGA_CHECK(c->fuse == 2, "conv was not created with backward IGDN fusion");
  if (c->P.n_spatial == 0) return SGA_OK;
  TcParams P = c->P;
  P.x_in = act_saved; P.s_in = s_saved; P.out_bf16 = gx_out;
  return launch_tc(c, P, st);
}

__global__ void k_pack_weights(const float* __restrict__ w, int taps, int Cin, int Cout, int transposed, int dup,
                               __nv_bfloat16* __restrict__ o) {
  // forward kinds: o[t][n=co][k] = w[t][k % Cin][co] for k < dup*Cin (dup = 2: the hi|lo split operand of the first layers
  // sees the same weights twice); dgrads: o[t][n=ci][k=co] = w[t][ci][co]
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(taps) * Cin * Cout * dup;
  if (e >= total) return;
  if (transposed) {
    o[e] = __float2bfloat16_rn(w[e]);
  } else {
    const int K = Cin * dup;
    const int k = static_cast<int>(e % K);
    const long long r = e / K;
    const int n = static_cast<int>(r % Cout);
    const int t = static_cast<int>(r / Cout);
    o[e] = __float2bfloat16_rn(w[(static_cast<long long>(t) * Cin + (k % Cin)) * Cout + n]);
  }
}

int launch_pack_weights_bf16(const float* w_ref, int taps, int Cin, int Cout, bool transposed, __nv_bfloat16* out,
                             cudaStream_t st, int dup) {
  if (transposed) dup = 1;
  const long long total = static_cast<long long>(taps) * Cin * Cout * dup;
  k_pack_weights<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w_ref, taps, Cin, Cout, transposed ? 1 : 0, dup,
                                                                              out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga
