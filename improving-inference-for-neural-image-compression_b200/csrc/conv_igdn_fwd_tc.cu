// conv_igdn_fwd_tc.cu -- synthesis-transform layer forward: transposed conv (tcgen05 implicit GEMM, conv_tc.cu) fused
// with the inverse GDN that follows it (nn_models.py:48-59, tfc.GDN(inverse=True)):
//     x = conv(in) + bias;   s = sqrt(beta + x^2 @ gamma);   act = x * s        (stores s and act as bf16)
// x itself is not stored: the backward pass recovers it as act / s.
//
// The channel mixing is a second GEMM on the finished accumulator tile.  In the two-pass epilogue of k_conv_tc (still
// used for the backward fusion) the same warps square the tile, wait for the mixing GEMM and then finish it, so
// per tile  [pass 1 -> mixing MMA -> pass 2]  is one serial chain and the tensor pipe idles whenever that chain is longer
// than the main loop (it is, for the 4- and 6-tap output phases).  Here the two passes belong to different warps and
// overlap across tiles:
//   warp 0      TMA producer (A / weight ring; gamma is loaded once and stays resident in shared memory)
//   warp 1      MMA issuer: main loop of tile i, and -- whenever the operands are ready (mbarrier.test_wait, no
//               blocking before the end of the loop) -- the mixing GEMM of tile i-1 in 64-column parts into two
//               alternating 64-column TMEM buffers
//   warps 2-5   P1: accumulator (TMEM) -> (x + bias)^2 as the bf16 K-major A operand of the mixing GEMM (shared memory)
//   warps 6-13  P2: accumulator + mixing result (TMEM) -> s, act -> global memory; releases the accumulator
// CTA pairs (CG = 2) as in conv_tc.cu: tcgen05.mma.cta_group::2, each CTA holds half of the weight / gamma rows.
#include "conv_tc.cuh"

namespace sga {

constexpr int kFwdThreads = 64 + 128 + 256;

template <int NF, int CG>
__global__ void __launch_bounds__(kFwdThreads, 1) k_conv_igdn_fwd(const __grid_constant__ TcParams P) {
  constexpr int KC = NF / 64;                       // K chunks of the mixing GEMM
  constexpr int NP = NF / 64;                       // 64-column parts of its result
  constexpr int kGTile = (64 / CG) * 128;           // gamma rows of one part held by this CTA x one K chunk
  constexpr int kGBytes = NP * KC * kGTile;
  constexpr int kA2Bytes = KC * kABytes;
  pdl_launch_dependents();
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0u);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int b_bytes = (NF / CG) * kBlockK * 2;
  // operand ring: stages x (A tile | weight tile), or -- halo mode -- a ring of halo tiles followed by a ring of weight tiles
  const int ring = P.halo ? P.a_stages * kHaloBytes + P.stages * b_bytes : P.stages * (kABytes + b_bytes);
  const uint32_t a_base = base;
  const uint32_t b_base = base + (P.halo ? P.a_stages * kHaloBytes : P.stages * kABytes);
  const uint32_t a2_base = base + ring;
  uint8_t* a2_ptr = smem + ring;
  const uint32_t g_base = a2_base + kA2Bytes;
  float* s_par = reinterpret_cast<float*>(smem + ring + kA2Bytes + kGBytes);       // bias[NF] | beta[NF]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_par + 2 * NF);
  // full[8] | empty[8] | tfull[2] | tempty[2] | a2full | a2empty | nfull[2] | nempty[2] | gbar | afull[2] | aempty[2]
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t bar_a2full = bar_tempty + 16;
  const uint32_t bar_a2empty = bar_a2full + 8;
  const uint32_t bar_nfull = bar_a2empty + 8;
  const uint32_t bar_nempty = bar_nfull + 16;
  const uint32_t bar_g = bar_nempty + 16;
  const uint32_t bar_afull = bar_g + 8;
  const uint32_t bar_aempty = bar_afull + 16;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 15);
  const uint32_t lead_full = (CG == 2) ? mapa_shared(bar_full, 0) : bar_full;
  const uint32_t lead_afull = (CG == 2) ? mapa_shared(bar_afull, 0) : bar_afull;
  const uint32_t lead_tempty = (CG == 2) ? mapa_shared(bar_tempty, 0) : bar_tempty;
  const uint32_t lead_a2full = (CG == 2) ? mapa_shared(bar_a2full, 0) : bar_a2full;
  const uint32_t lead_nempty = (CG == 2) ? mapa_shared(bar_nempty, 0) : bar_nempty;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kNormCol = 2 * NF;             // two 64-column mixing-result buffers behind the two accumulators

  for (int i = threadIdx.x; i < NF; i += blockDim.x) {
    s_par[i] = P.bias ? P.bias[i] : 0.0f;
    s_par[NF + i] = P.beta[i];
  }
  if (warp == 0 && lane == 0) {
    for (int v = 0; v < kMaxViews; ++v) tma_prefetch_desc(&P.amap[v]);
    tma_prefetch_desc(&P.bmap);
    tma_prefetch_desc(&P.gmap);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 8 * CG);
      mbar_init(bar_nfull + 8 * b, 1);
      mbar_init(bar_nempty + 8 * b, 8 * CG);
    }
    mbar_init(bar_a2full, 4 * CG);
    mbar_init(bar_a2empty, 1);
    mbar_init(bar_g, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_afull + 8 * b, 1);
      mbar_init(bar_aempty + 8 * b, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_cg2(smem_u32(s_tmem), kTmemCols); tmem_relinquish_cg2(); }
    else { tmem_alloc(smem_u32(s_tmem), kTmemCols); tmem_relinquish(); }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  pdl_grid_dependency_wait();        // everything above touches only constants (bias, beta) and on-chip state
  const uint32_t tmem_base = *s_tmem;
  const int nwaves = num_waves<CG>(P);
  auto arrive_lead = [](uint32_t local_bar, uint32_t lead_bar) {
    if (CG == 2) mbar_arrive_cluster(lead_bar); else mbar_arrive(local_bar);
  };
  // wait-cycle accounting (clock64 / printf) is a development aid: compiled in only with -DSGA_DBG (build.py --dbg),
  // otherwise dbg_t is a compile-time false and every instrumented statement reduces to the statement itself
#ifdef SGA_DBG
  const bool dbg_t = (P.dbg & 8) != 0;
#else
  constexpr bool dbg_t = false;
#endif
  long long dbg_w[4] = {0, 0, 0, 0};
  const long long dbg_t0 = dbg_t ? clock64() : 0;
  const bool dbg_print = dbg_t && lane == 0 && (blockIdx.x == 0 || blockIdx.x == 74);
#define SGA_TIMED(slot, stmt)                                  \
  do {                                                         \
    if (dbg_t) { const long long _c = clock64(); stmt; dbg_w[slot] += clock64() - _c; } else { stmt; } \
  } while (0)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // gamma: rows [part*64 + rank*64/CG, +64/CG) of every part, all K chunks; stays resident
      mbar_arrive_expect_tx(bar_g, kGBytes);
      for (int part = 0; part < NP; ++part)
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(g_base + (part * KC + kc) * kGTile, &P.gmap, bar_g, kc * kBlockK,
                      part * 64 + static_cast<int>(rank) * (64 / CG));
    }
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage_tx = static_cast<uint32_t>(CG * (kABytes + b_bytes));
    const int b_row_off = static_cast<int>(rank) * (NF / 2);
    int astage = 0;
    uint32_t aphase = 0;
    for (int w = 0; w < nwaves; ++w) {
      TileCoord tc_;
      if (!decode_slot<CG>(P, w, rank, tc_)) continue;
      const TcJob& J = P.job[tc_.job];
      if (P.halo) {
        for (int gi = 0; gi < J.ngroups; ++gi) {
          const TcGroup grp_ = P.groups[J.group_begin + gi];
          const CUtensorMap* am = &P.amap[grp_.view];
          const int ax = tc_.w0 + grp_.ox, ay = tc_.h0 + grp_.oy;
          for (int kc = 0; kc < P.kchunks; ++kc) {
            mbar_wait(bar_aempty + 8 * astage, aphase ^ 1u);       // the halo tile of this (view, K chunk): all its taps share it
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(bar_afull + 8 * astage, CG * kHaloBytes);
              if (CG == 2) tma_load_4d_cg2(a_base + astage * kHaloBytes, am, lead_afull + 8 * astage, kc * kBlockK, ax, ay, tc_.img);
              else tma_load_4d(a_base + astage * kHaloBytes, am, bar_afull + 8 * astage, kc * kBlockK, ax, ay, tc_.img);
            }
            if (++astage == P.a_stages) { astage = 0; aphase ^= 1u; }
            for (int t = 0; t < grp_.ntaps; ++t) {
              const int brow = P.htaps[grp_.tap_begin + t].widx * P.Nn + (CG == 2 ? b_row_off : 0);
              SGA_TIMED(0, mbar_wait(bar_empty + 8 * stage, phase ^ 1u));
              if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, static_cast<uint32_t>(CG * b_bytes));
                if (CG == 2) tma_load_2d_cg2(b_base + stage * b_bytes, &P.bmap, lead_full + 8 * stage, kc * kBlockK, brow);
                else tma_load_2d(b_base + stage * b_bytes, &P.bmap, bar_full + 8 * stage, kc * kBlockK, brow);
              }
              if (++stage == P.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
        continue;
      }
      for (int t = 0; t < J.ntaps; ++t) {
        const Tap tap = P.taps[J.tap_begin + t];
        const CUtensorMap* am = &P.amap[tap.view];
        const int ax = tc_.w0 + tap.dx, ay = tc_.h0 + tap.dy, brow = tap.widx * P.Nn + (CG == 2 ? b_row_off : 0);
        for (int kc = 0; kc < P.kchunks; ++kc) {
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
    }
    if (dbg_print) printf("dbg fwd cta %d producer: total %lld wait_empty %lld\n", blockIdx.x, clock64() - dbg_t0, dbg_w[0]);
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (leader) {
      const uint32_t idesc = instr_desc_bf16(kBlockM * CG, NF);
      const uint32_t idesc_mix = instr_desc_bf16(kBlockM * CG, 64);
      auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
        if (CG == 2) umma_bf16_cg2(d, ad, bd, id, acc); else umma_bf16(d, ad, bd, id, acc);
      };
      auto commit = [](uint32_t bar) { if (CG == 2) umma_commit_cg2_both(bar); else umma_commit(bar); };
      const uint64_t adesc0 = smem_desc_sw128(a_base), bdesc0 = smem_desc_sw128(b_base), a2desc0 = smem_desc_sw128(a2_base),
                     gdesc0 = smem_desc_sw128(g_base);
      const uint64_t ahalo0 = smem_desc_sw128_sbo(a_base, kHaloPitch * 128);
      const uint32_t a_step = kABytes >> 4, b_step = static_cast<uint32_t>(b_bytes) >> 4;
      int astage = 0;
      uint32_t aphase = 0;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;          // tiles issued
      int mt = -1;         // tile whose mixing GEMM is in progress (it - 1)
      int mpart = NP;      // next part of that GEMM (NP = done)
      int nparts = 0;      // parts issued so far (selects the TMEM result buffer)
      // issue as many parts of the pending mixing GEMM as have their operands; `block`: wait for them
      auto mix_advance = [&](bool block) {
        while (mpart < NP) {
          const int nb = nparts & 1;
          const uint32_t npar = ((nparts >> 1) & 1) ^ 1u;
          if (!block) {
            bool ok = mbar_test(bar_nempty + 8 * nb, npar);
            if (mpart == 0) ok = ok && mbar_test(bar_a2full, mt & 1);
            if (!__all_sync(0xffffffffu, ok)) break;
          }
          if (mpart == 0) SGA_TIMED(2, mbar_wait(bar_a2full, mt & 1));      // (x + bias)^2 of tile mt is in shared memory
          SGA_TIMED(3, mbar_wait(bar_nempty + 8 * nb, npar));               // result buffer drained by P2
          tcgen05_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
              const uint64_t ad = a2desc0 + static_cast<uint64_t>(kc * (kABytes >> 4));
              const uint64_t bd = gdesc0 + static_cast<uint64_t>(((mpart * KC + kc) * kGTile) >> 4);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma(tmem_base + kNormCol + nb * 64, ad + 2 * kk, bd + 2 * kk, idesc_mix, (kc | kk) != 0 ? 1u : 0u);
            }
            commit(bar_nfull + 8 * nb);
            if (mpart == NP - 1) commit(bar_a2empty);      // the A operand may be overwritten by P1 of the next tile
          }
          ++mpart;
          ++nparts;
        }
      };
      mbar_wait(bar_g, 0);                                  // (the peer's half is covered by its P1 warps' first arrive)
      for (int w = 0; w < nwaves; ++w) {
        TileCoord tc_;
        if (!decode_slot<CG>(P, w, rank, tc_)) continue;
        const int nk = P.job[tc_.job].ntaps * P.kchunks;
        const int buf = it & 1;
        SGA_TIMED(1, mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1u));
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * NF;
        if (P.halo) {
          const TcJob& J = P.job[tc_.job];
          int k = 0;
          for (int gi = 0; gi < J.ngroups; ++gi) {
            const TcGroup grp_ = P.groups[J.group_begin + gi];
            for (int kc = 0; kc < P.kchunks; ++kc) {
              SGA_TIMED(0, mbar_wait(bar_afull + 8 * astage, aphase));
              for (int t = 0; t < grp_.ntaps; ++t, ++k) {
                const TcHaloTap tap = P.htaps[grp_.tap_begin + t];
                SGA_TIMED(0, mbar_wait(bar_full + 8 * stage, phase));
                tcgen05_fence_after();
                if (elect_one()) {
                  // window of this tap inside the halo tile: start row dy'*16 + dx', 8-row groups 2048 B apart
                  const uint64_t ad = ahalo0 + static_cast<uint64_t>(astage * (kHaloBytes >> 4) + tap.a_off16) +
                                      (static_cast<uint64_t>(tap.boff) << 49);
                  const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * b_step);
#pragma unroll
                  for (int kk = 0; kk < kBlockK / 16; ++kk) mma(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
                  commit(bar_empty + 8 * stage);
                }
                if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                if (mpart < NP) mix_advance(false);
              }
              if (elect_one()) commit(bar_aempty + 8 * astage);     // halo tile free after its last tap
              if (++astage == P.a_stages) { astage = 0; aphase ^= 1u; }
            }
          }
        } else
        for (int k = 0; k < nk; ++k) {
          SGA_TIMED(0, mbar_wait(bar_full + 8 * stage, phase));
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * a_step);
            const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * b_step);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) mma(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            commit(bar_empty + 8 * stage);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          if (mpart < NP) mix_advance(false);
        }
        if (elect_one()) commit(bar_tfull + 8 * buf);
        mix_advance(true);     // tile it-1 must be finished before its accumulator buffer is awaited for tile it+1
        mt = it;
        mpart = 0;
        ++it;
      }
      mix_advance(true);
      if (dbg_print) {
        printf("dbg fwd cta %d mma: total %lld wait_full %lld wait_tempty %lld wait_a2 %lld wait_nempty %lld tiles %d\n",
               blockIdx.x, clock64() - dbg_t0, dbg_w[0], dbg_w[1], dbg_w[2], dbg_w[3], it);
      }
    }
  } else if (warp < 6) {
    // ===================== P1: accumulator -> squared operand of the mixing GEMM =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    int it = 0;
    for (int w = 0; w < nwaves; ++w) {
      TileCoord tc_;
      if (!decode_slot<CG>(P, w, rank, tc_)) continue;
      const int buf = it & 1;
      SGA_TIMED(0, mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1));
      if (it > 0) SGA_TIMED(1, mbar_wait(bar_a2empty, (it - 1) & 1));   // mixing GEMM of the previous tile has read the operand
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + buf * NF + lane_off;
#pragma unroll
      for (int c2 = 0; c2 < KC; ++c2) {
        uint32_t r[2][32];
        tmem_ld32(taddr + c2 * 64, r[0]);
        tmem_ld32(taddr + c2 * 64 + 32, r[1]);
        tmem_ld_wait();
        uint8_t* a2c = a2_ptr + c2 * kABytes;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float* bp = s_par + c2 * 64 + h * 32;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 b0 = *reinterpret_cast<const float4*>(bp + 8 * u);
            const float4 b1 = *reinterpret_cast<const float4*>(bp + 8 * u + 4);
            const float x0 = __uint_as_float(r[h][8 * u + 0]) + b0.x, x1 = __uint_as_float(r[h][8 * u + 1]) + b0.y;
            const float x2 = __uint_as_float(r[h][8 * u + 2]) + b0.z, x3 = __uint_as_float(r[h][8 * u + 3]) + b0.w;
            const float x4 = __uint_as_float(r[h][8 * u + 4]) + b1.x, x5 = __uint_as_float(r[h][8 * u + 5]) + b1.y;
            const float x6 = __uint_as_float(r[h][8 * u + 6]) + b1.z, x7 = __uint_as_float(r[h][8 * u + 7]) + b1.w;
            *reinterpret_cast<uint4*>(a2c + sw128_offset(row, h * 4 + u)) =
                make_uint4(pack_bf16(x0 * x0, x1 * x1), pack_bf16(x2 * x2, x3 * x3), pack_bf16(x4 * x4, x5 * x5),
                           pack_bf16(x6 * x6, x7 * x7));
          }
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (it == 0) mbar_wait(bar_g, 0);    // this CTA's gamma rows are resident before the leader may issue the mixing
      // default (.release.cta) arrive as everywhere else: the operand sits in THIS CTA's shared memory and was made
      // visible to the async proxy by the fence above; CUTLASS signals its 2-SM mainloops the same way.  (A
      // .release.cluster arrive costs a MEMBAR.ALL.CTA per warp and tile: +3 % kernel time.)
      if (lane == 0) arrive_lead(bar_a2full, lead_a2full);
      ++it;
    }
    if (dbg_print && warp == 2)
      printf("dbg fwd cta %d p1: total %lld wait_tfull %lld wait_a2empty %lld tiles %d\n", blockIdx.x, clock64() - dbg_t0, dbg_w[0],
             dbg_w[1], it);
  } else {
    // ===================== P2: accumulator + mixing result -> s, act =====================
    const int q = warp & 3;
    const int cgp = (warp - 6) >> 2;             // 32-column half of each 64-column part
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    int it = 0, np = 0;
    for (int w = 0; w < nwaves; ++w) {
      TileCoord tc_;
      if (!decode_slot<CG>(P, w, rank, tc_)) continue;
      const TcJob& J = P.job[tc_.job];
      const int buf = it & 1;
      const int i = tc_.h0 + row / P.TW, j = tc_.w0 + row % P.TW;
      const bool valid = (i < J.Ho) && (j < J.Wo) && (tc_.img < P.N);
      const long long off0 = J.out_off + tc_.img * P.oN + i * J.oH + j * J.oW;
#pragma unroll
      for (int part = 0; part < NP; ++part, ++np) {
        const int nb = np & 1;
        SGA_TIMED(0, mbar_wait(bar_nfull + 8 * nb, (np >> 1) & 1));
        tcgen05_fence_after();
        const int col = part * 64 + cgp * 32;
        uint32_t m[32], a[32];
        tmem_ld32(tmem_base + kNormCol + nb * 64 + cgp * 32 + lane_off, m);
        tmem_ld32(tmem_base + buf * NF + col + lane_off, a);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          arrive_lead(bar_nempty + 8 * nb, lead_nempty + 8 * nb);
          if (part == NP - 1) arrive_lead(bar_tempty + 8 * buf, lead_tempty + 8 * buf);   // accumulator fully consumed
        }
        uint32_t so[16], oo[16];
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 bi = *reinterpret_cast<const float4*>(s_par + col + 4 * e4);
          const float4 be = *reinterpret_cast<const float4*>(s_par + NF + col + 4 * e4);
          const float x0 = __uint_as_float(a[4 * e4 + 0]) + bi.x, x1 = __uint_as_float(a[4 * e4 + 1]) + bi.y;
          const float x2 = __uint_as_float(a[4 * e4 + 2]) + bi.z, x3 = __uint_as_float(a[4 * e4 + 3]) + bi.w;
          const float s0 = fast_sqrt(__uint_as_float(m[4 * e4 + 0]) + be.x), s1 = fast_sqrt(__uint_as_float(m[4 * e4 + 1]) + be.y);
          const float s2 = fast_sqrt(__uint_as_float(m[4 * e4 + 2]) + be.z), s3 = fast_sqrt(__uint_as_float(m[4 * e4 + 3]) + be.w);
          so[2 * e4] = pack_bf16(s0, s1);
          so[2 * e4 + 1] = pack_bf16(s2, s3);
          // act uses the stored (bf16) s, so that the backward's x = act / s inverts it consistently
          oo[2 * e4] = pack_bf16(x0 * bf16_lo(so[2 * e4]), x1 * bf16_hi(so[2 * e4]));
          oo[2 * e4 + 1] = pack_bf16(x2 * bf16_lo(so[2 * e4 + 1]), x3 * bf16_hi(so[2 * e4 + 1]));
        }
        if (valid) {
          st_global_v8(P.s_out + off0 + col, &so[0]);
          st_global_v8(P.s_out + off0 + col + 16, &so[8]);
          st_global_v8(P.out_bf16 + off0 + col, &oo[0]);
          st_global_v8(P.out_bf16 + off0 + col + 16, &oo[8]);
        }
      }
      ++it;
    }
    if (dbg_print && warp == 6)
      printf("dbg fwd cta %d p2: total %lld wait_nfull %lld tiles %d\n", blockIdx.x, clock64() - dbg_t0, dbg_w[0], it);
  }
#undef SGA_TIMED

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int igdn_fwd_smem_fixed(int cout, int cg) {
  const int kc = cout / 64;
  return 1024 + kc * kABytes + kc * kc * (64 / cg) * 128 + 2 * cout * 4 + 256;
}

typedef void (*FwdKernel)(const TcParams);
static FwdKernel select_fwd(int nf, int cg) {
  if (cg == 2) return nf == 192 ? k_conv_igdn_fwd<192, 2> : (nf == 128 ? k_conv_igdn_fwd<128, 2> : k_conv_igdn_fwd<64, 2>);
  return nf == 192 ? k_conv_igdn_fwd<192, 1> : (nf == 128 ? k_conv_igdn_fwd<128, 1> : k_conv_igdn_fwd<64, 1>);
}

int igdn_fwd_set_attr(int cout, int cg) {
  SGA_CUDA(cudaFuncSetAttribute(select_fwd(cout, cg), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return SGA_OK;
}

int launch_igdn_fwd_kernel(const TcConv* c, const TcParams& P, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(c->grid), 1, 1);
  cfg.blockDim = dim3(kFwdThreads, 1, 1);
  cfg.dynamicSmemBytes = static_cast<size_t>(c->smem_bytes);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(c->cg);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  SGA_CUDA(cudaLaunchKernelEx(&cfg, select_fwd(P.Nn, c->cg), P));
  g_launches += 1;
  return SGA_OK;
}

}  // namespace sga
