// conv_tc.cu -- bf16 implicit-GEMM (de)convolution on the sm_100a tensor cores.
//
// Every (de)conv of the synthesis / hyper-synthesis transforms (nn_models.py:39-70, 140-170) and its dgrad is a
// sum over filter taps of  [128 pixels x Cin] * [Cin x Cout]  products (conv_plan.cu).  Per tap and 64-channel slice
//   A tile : TMA 4-D box {64 ch, TW, TH, 1} of the NHWC activation at the tap's spatial offset -- out-of-bounds
//            pixels/channels are zero-filled by the TMA unit, which implements "same_zeros" padding and ragged edges;
//            the box lands in shared memory as a K-major SWIZZLE_128B operand tile (128 rows x 128 B)
//   B tile : TMA 2-D box {64 ch, BLOCK_N} of the pre-packed K-major weights [tap][Cout][Cin]
//   MMA    : 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) accumulating in TMEM (fp32)
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
// (tcgen05.ld -> bias / ReLU / ReLU-mask -> bf16 and/or fp32 global stores).  Persistent CTAs (one per SM) walk a
// static tile list; the TMEM accumulator is double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "kernels.h"
#include "tc_common.cuh"

namespace sga {

using namespace tc;

constexpr int kTcThreads = 192;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;

struct TcJob {
  int tap_begin, ntaps;
  int Ho, Wo;             // pixel grid of the output view
  int tiles_h, tiles_w;
  int tile_begin;         // first global tile index of this job
  int pad_;
  long long out_off;      // element offset of the output view
  long long oH, oW;       // element strides of the output view
};

struct alignas(64) TcParams {
  CUtensorMap amap[kMaxViews];
  CUtensorMap bmap;
  TcJob job[4];
  Tap taps[kMaxTaps];
  int njobs, N, K, Nn;
  int block_n, n_tiles_n, kchunks, stages;
  int TH, TW, total_tiles, relu;
  long long oN;
  const float* bias;
  const __nv_bfloat16* mask;
  __nv_bfloat16* out_bf16;
  float* out_f32;
};

struct TileCoord { int job, img, h0, w0, n0; };

__device__ __forceinline__ TileCoord decode_tile(const TcParams& P, int tile) {
  int jb = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q)
    if (q < P.njobs && tile >= P.job[q].tile_begin) jb = q;
  const TcJob& J = P.job[jb];
  int t = tile - J.tile_begin;
  const int nt = t % P.n_tiles_n; t /= P.n_tiles_n;
  const int tw = t % J.tiles_w; t /= J.tiles_w;
  const int th = t % J.tiles_h;
  const int img = t / J.tiles_h;
  return TileCoord{jb, img, th * P.TH, tw * P.TW, nt * P.block_n};
}

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
    uint4* op = reinterpret_cast<uint4*>(P.out_bf16 + off);
#pragma unroll
    for (int q = 0; q < NV / 8; ++q)
      op[q] = make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                         pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
  }
  if (P.out_f32) {
    float4* op = reinterpret_cast<float4*>(P.out_f32 + off);
#pragma unroll
    for (int q = 0; q < NV / 4; ++q) op[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) k_conv_tc(const __grid_constant__ TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int b_bytes = P.block_n * kBlockK * 2;
  const uint32_t a_base = base;
  const uint32_t b_base = base + P.stages * kABytes;
  uint8_t* tail = smem + P.stages * (kABytes + b_bytes);
  float* s_bias = reinterpret_cast<float*>(tail);                       // 512 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 512 * sizeof(float));
  // bars: full[S] | empty[S] | tmem_full[2] | tmem_empty[2]
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < P.Nn && i < 512; i += blockDim.x) s_bias[i] = P.bias ? P.bias[i] : 0.0f;

  if (warp == 0 && lane == 0) {
    for (int v = 0; v < kMaxViews; ++v) tma_prefetch_desc(&P.amap[v]);
    tma_prefetch_desc(&P.bmap);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(s_tmem), kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int ksteps_per_tap = P.kchunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TileCoord tc_ = decode_tile(P, tile);
        const TcJob& J = P.job[tc_.job];
        for (int t = 0; t < J.ntaps; ++t) {
          const Tap tap = P.taps[J.tap_begin + t];
          for (int kc = 0; kc < ksteps_per_tap; ++kc) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            mbar_arrive_expect_tx(bar_full + 8 * stage, kABytes + b_bytes);
            tma_load_4d(a_base + stage * kABytes, &P.amap[tap.view], bar_full + 8 * stage, kc * kBlockK,
                        tc_.w0 + tap.dx, tc_.h0 + tap.dy, tc_.img);
            tma_load_2d(b_base + stage * b_bytes, &P.bmap, bar_full + 8 * stage, kc * kBlockK,
                        tap.widx * P.Nn + tc_.n0);
            if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = instr_desc_bf16(kBlockM, P.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
        const TileCoord tc_ = decode_tile(P, tile);
        const int nk = P.job[tc_.job].ntaps * ksteps_per_tap;
        const int buf = it & 1;
        mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        for (int k = 0; k < nk; ++k) {
          mbar_wait(bar_full + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t a_addr = a_base + stage * kABytes, b_addr = b_base + stage * b_bytes;
#pragma unroll
          for (int kk = 0; kk < kBlockK / 16; ++kk) {
            umma_bf16(d_tmem, smem_desc_sw128(a_addr + kk * 32), smem_desc_sw128(b_addr + kk * 32), idesc,
                      (k | kk) != 0 ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * stage);   // frees the smem stage once these MMAs have read it
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_tfull + 8 * buf);       // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;               // tile row = pixel
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const TileCoord tc_ = decode_tile(P, tile);
      const TcJob& J = P.job[tc_.job];
      const int buf = it & 1;
      const int i = tc_.h0 + row / P.TW, j = tc_.w0 + row % P.TW;
      const bool valid = (i < J.Ho) && (j < J.Wo);
      const long long off0 = J.out_off + tc_.img * P.oN + i * J.oH + j * J.oW + tc_.n0;
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + buf * 256 + (static_cast<uint32_t>(q * 32) << 16);
      int col = 0;
      for (; col + 32 <= P.block_n; col += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + col, r);
        tmem_ld_wait();
        epilogue_store<32>(P, r, s_bias, tc_.n0 + col, off0 + col, valid);
      }
      if (col < P.block_n) {                     // 16-column tail (block_n % 32 == 16)
        uint32_t r[16];
        tmem_ld16(taddr + col, r);
        tmem_ld_wait();
        epilogue_store<16>(P, r, s_bias, tc_.n0 + col, off0 + col, valid);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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

struct TcConv {
  TcParams P;
  int smem_bytes;
  int grid;
};

static int pick_block_n(int Nn) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (Nn % bn == 0) return bn;
  return 0;
}

int tc_conv_create(const ConvPlan& plan, const __nv_bfloat16* x, const __nv_bfloat16* w_kmajor, TcConv** out) {
  const ConvGeom& g = plan.g;
  SGA_CHECK(g.Cin % 8 == 0, "tcgen05 conv needs Cin %% 8 == 0 (got %d)", g.Cin);
  SGA_CHECK(g.Cout % 16 == 0 && g.Cout <= 512, "tcgen05 conv needs Cout %% 16 == 0 and <= 512 (got %d)", g.Cout);
  TcConv* c = new TcConv();
  memset(&c->P, 0, sizeof(c->P));
  TcParams& P = c->P;
  P.N = g.N; P.K = g.Cin; P.Nn = g.Cout;
  P.block_n = pick_block_n(g.Cout);
  P.n_tiles_n = g.Cout / P.block_n;
  P.kchunks = ceil_div(g.Cin, kBlockK);
  P.njobs = plan.njobs;
  P.oN = static_cast<long long>(g.Hout) * g.Wout * g.Cout;
  // spatial tile: 128 pixels = TH x TW
  const int Wo0 = plan.job[0].Wo;
  P.TW = (Wo0 >= 16) ? 16 : 8;
  P.TH = kBlockM / P.TW;
  const int b_bytes = P.block_n * kBlockK * 2;
  int stages = (227 * 1024 - 1024 - 512 * 4 - 256) / (kABytes + b_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  P.stages = stages;
  c->smem_bytes = 1024 + stages * (kABytes + b_bytes) + 512 * 4 + 256;
  int ntap = 0, ntile = 0;
  for (int jn = 0; jn < plan.njobs; ++jn) {
    const ConvJobPlan& jp = plan.job[jn];
    TcJob& J = P.job[jn];
    J.tap_begin = ntap; J.ntaps = jp.ntaps;
    for (int t = 0; t < jp.ntaps; ++t) P.taps[ntap++] = jp.taps[t];
    J.Ho = jp.Ho; J.Wo = jp.Wo;
    J.tiles_h = ceil_div(jp.Ho, P.TH); J.tiles_w = ceil_div(jp.Wo, P.TW);
    J.tile_begin = ntile;
    ntile += g.N * J.tiles_h * J.tiles_w * P.n_tiles_n;
    J.out_off = (jp.out_phase_stride == 2) ? (static_cast<long long>(jp.out_py) * g.Wout + jp.out_px) * g.Cout : 0;
    J.oH = static_cast<long long>(jp.out_phase_stride) * g.Wout * g.Cout;
    J.oW = static_cast<long long>(jp.out_phase_stride) * g.Cout;
  }
  P.total_tiles = ntile;
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
    const uint32_t box[4] = {kBlockK, static_cast<uint32_t>(P.TW), static_cast<uint32_t>(P.TH), 1};
    const __nv_bfloat16* basep = x + (static_cast<long long>(py) * g.Win + px) * g.Cin;
    int st = tma_encode_tiled(&P.amap[v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, basep, dims, strides, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != SGA_OK) { delete c; return st; }
  }
  {
    const int taps = g.kh * g.kw;
    const uint64_t dims[2] = {static_cast<uint64_t>(g.Cin), static_cast<uint64_t>(taps) * g.Cout};
    const uint64_t strides[1] = {static_cast<uint64_t>(g.Cin) * 2};
    const uint32_t box[2] = {kBlockK, static_cast<uint32_t>(P.block_n)};
    int st = tma_encode_tiled(&P.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_kmajor, dims, strides, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != SGA_OK) { delete c; return st; }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  c->grid = (ntile < sms) ? ntile : sms;
  static bool attr_set = false;
  if (!attr_set) {
    SGA_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  *out = c;
  return SGA_OK;
}

void tc_conv_destroy(TcConv* c) { delete c; }

int tc_conv_launch(const TcConv* c, const float* bias, int relu, const __nv_bfloat16* relu_mask, __nv_bfloat16* out_bf16,
                   float* out_f32, cudaStream_t st) {
  if (c->P.total_tiles == 0) return SGA_OK;
  TcParams P = c->P;
  P.bias = bias; P.relu = relu; P.mask = relu_mask; P.out_bf16 = out_bf16; P.out_f32 = out_f32;
  k_conv_tc<<<c->grid, kTcThreads, c->smem_bytes, st>>>(P);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

__global__ void k_pack_weights(const float* __restrict__ w, int taps, int Cin, int Cout, int transposed,
                               __nv_bfloat16* __restrict__ o) {
  // forward kinds: o[t][n=co][k=ci] = w[t][ci][co]; dgrads: o[t][n=ci][k=co] = w[t][ci][co]
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(taps) * Cin * Cout;
  if (e >= total) return;
  if (transposed) {
    o[e] = __float2bfloat16_rn(w[e]);
  } else {
    const int k = static_cast<int>(e % Cin);
    const long long r = e / Cin;
    const int n = static_cast<int>(r % Cout);
    const int t = static_cast<int>(r / Cout);
    o[e] = __float2bfloat16_rn(w[(static_cast<long long>(t) * Cin + k) * Cout + n]);
  }
}

int launch_pack_weights_bf16(const float* w_ref, int taps, int Cin, int Cout, bool transposed, __nv_bfloat16* out,
                             cudaStream_t st) {
  const long long total = static_cast<long long>(taps) * Cin * Cout;
  k_pack_weights<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w_ref, taps, Cin, Cout, transposed ? 1 : 0,
                                                                              out);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga
