// igdn_tc.cu -- inverse GDN (tfc.GDN(inverse=True), nn_models.py:51,55,59) forward and backward on bf16
// activations, with the CxC channel mixing on the tensor cores.
//   forward : s = sqrt(beta + x^2 @ gamma),  out = x * s                     (saves s for the backward)
//   backward: g_x = g * s + x * ((g * x / s) @ gamma^T)
// One persistent CTA per SM: gamma (K-major bf16, SWIZZLE_128B) stays resident in shared memory; per 128-pixel tile
//   stage 1 (8 worker warps, coalesced 16-byte loads): build the A operand (x^2 or g*x/s) directly in the swizzled
//            K-major layout in shared memory (double-buffered), fence.proxy.async, arrive on a2_full
//   stage 2 (warp 0, one thread): C/16 tcgen05.mma into a double-buffered TMEM accumulator, commit -> acc_full
//   stage 3 (workers, one pixel row per thread, half of the channels per warp): tcgen05.ld + point-wise finish + store
// Stage 1 of tile t+1 is issued before stage 3 of tile t, so the MMA and the global-load latency overlap.
#include "kernels.h"
#include "tc_common.cuh"

namespace sga {

using namespace tc;

constexpr int kIgdnWorkers = 256;
constexpr int kIgdnThreads = 32 + kIgdnWorkers;

struct alignas(64) IgdnParams {
  CUtensorMap bmap;           // [n rows][k cols] bf16 operand of the 1x1 mixing
  const __nv_bfloat16* x;     // pre-activation (forward input)
  const __nv_bfloat16* g;     // backward: upstream gradient
  const __nv_bfloat16* s;     // backward: saved sqrt(norm)
  const float* beta;          // forward
  __nv_bfloat16* out;         // forward: x*s ; backward: g_x
  __nv_bfloat16* s_out;       // forward
  long long pixels;
  int C, block_n, n_tiles_n, bwd;
  int x_is_act;               // backward: `x` holds the saved activation act = x*s (the fused forward does not store x)
  int abufs;                  // A-operand buffers: 2, or 1 when only that lets the whole channel range fit (C = 256)
};

__device__ __forceinline__ float igdn_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
    f[2 * e] = __low2float(b2);
    f[2 * e + 1] = __high2float(b2);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

__global__ void __launch_bounds__(kIgdnThreads, 1) k_igdn_tc(const __grid_constant__ IgdnParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int kchunks = P.C / 64;
  const int b_chunk_bytes = P.block_n * 128;
  const int a_chunk_bytes = 128 * 128;
  const uint32_t b_base = base;
  const uint32_t a_base0 = base + kchunks * b_chunk_bytes;
  const int a_buf_bytes = kchunks * a_chunk_bytes;
  uint8_t* a_ptr0 = smem + kchunks * b_chunk_bytes;
  uint8_t* tail = a_ptr0 + P.abufs * a_buf_bytes;
  float* s_beta = reinterpret_cast<float*>(tail);                       // 256 floats (this CTA's n-tile)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 256 * sizeof(float));
  const uint32_t bar_b = smem_u32(bars);          // gamma loaded
  const uint32_t bar_a2 = bar_b + 8;              // [2] A operand ready (count = workers)
  const uint32_t bar_acc = bar_a2 + 16;           // [2] accumulator ready
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x % P.n_tiles_n;                 // this CTA's channel tile
  const int n0 = nt * P.block_n;
  const int cta_in_group = blockIdx.x / P.n_tiles_n;
  const int group_ctas = gridDim.x / P.n_tiles_n;
  const long long num_tiles = (P.pixels + 127) / 128;

  if (!P.bwd)
    for (int i = threadIdx.x; i < P.block_n; i += blockDim.x) s_beta[i] = P.beta[n0 + i];

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&P.bmap);
      mbar_init(bar_b, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_a2 + 8 * b, kIgdnWorkers);
        mbar_init(bar_acc + 8 * b, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      // resident B operand
      mbar_arrive_expect_tx(bar_b, kchunks * b_chunk_bytes);
      for (int kc = 0; kc < kchunks; ++kc) tma_load_2d(b_base + kc * b_chunk_bytes, &P.bmap, bar_b, kc * 64, n0);
      mbar_wait(bar_b, 0);
      const uint32_t idesc = instr_desc_bf16(128, P.block_n);
      int it = 0;
      for (long long tile = cta_in_group; tile < num_tiles; tile += group_ctas, ++it) {
        const int buf = it & 1;
        mbar_wait(bar_a2 + 8 * buf, (it >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        const uint32_t a_addr = a_base0 + (P.abufs == 2 ? buf : 0) * a_buf_bytes;
        for (int kc = 0; kc < kchunks; ++kc) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(d_tmem, smem_desc_sw128(a_addr + kc * a_chunk_bytes + kk * 32),
                      smem_desc_sw128(b_base + kc * b_chunk_bytes + kk * 32), idesc, (kc | kk) != 0 ? 1u : 0u);
        }
        umma_commit(bar_acc + 8 * buf);
      }
    }
  } else {
    const int wt = threadIdx.x - 32;            // worker thread 0..255
    const int ww = wt >> 5;                     // worker warp 0..7
    const int q = warp & 3;                     // TMEM lane quarter accessible to this warp
    const int half = ww >> 2;                   // which half of the block_n columns this warp finishes
    const int row = q * 32 + lane;
    const int vec_per_row = P.C / 8;            // 16-byte vectors per pixel
    const int vecs = 128 * vec_per_row;
    const int cols_half = P.block_n / 2;

    auto stage1 = [&](long long tile, int buf) {
      uint8_t* a_buf = a_ptr0 + (P.abufs == 2 ? buf : 0) * a_buf_bytes;
      const long long pix0 = tile * 128;
      for (int v0 = wt; v0 < vecs; v0 += kIgdnWorkers * 4) {
        uint4 xv[4], gv[4], sv[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int v = v0 + u * kIgdnWorkers;
          const int r = v / vec_per_row, c16 = v % vec_per_row;
          ok[u] = (v < vecs) && (pix0 + r < P.pixels);
          if (ok[u]) {
            const long long off = (pix0 + r) * P.C + c16 * 8;
            xv[u] = __ldg(reinterpret_cast<const uint4*>(P.x + off));
            if (P.bwd) {
              gv[u] = __ldg(reinterpret_cast<const uint4*>(P.g + off));
              sv[u] = __ldg(reinterpret_cast<const uint4*>(P.s + off));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int v = v0 + u * kIgdnWorkers;
          if (v >= vecs) continue;
          const int r = v / vec_per_row, c16 = v % vec_per_row;
          float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (ok[u]) {
            float xf[8];
            unpack8(xv[u], xf);
            if (P.bwd) {
              float gf[8], sf[8];
              unpack8(gv[u], gf);
              unpack8(sv[u], sf);
              // rcp.approx + multiplies, as in the fused epilogues (three IEEE divisions per element made this kernel
              // ALU-bound: 4.4 ms for the 604 M elements of IGDN_2 at C = 256)
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float rs = igdn_rcp(sf[e]);
                const float xq = P.x_is_act ? xf[e] * rs : xf[e];
                t[e] = gf[e] * xq * rs;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) t[e] = xf[e] * xf[e];
            }
          }
          const int kc = c16 >> 3, c = c16 & 7;
          *reinterpret_cast<uint4*>(a_buf + kc * a_chunk_bytes + sw128_offset(r, c)) = pack8(t);
        }
      }
      fence_proxy_async_smem();                 // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(bar_a2 + 8 * buf);
    };

    auto stage3 = [&](long long tile, int buf) {
      const long long pix = tile * 128 + row;
      const bool valid = pix < P.pixels;
      const uint32_t taddr = tmem_base + buf * 256 + (static_cast<uint32_t>(q * 32) << 16) + half * cols_half;
      for (int col = 0; col < cols_half; col += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + col, r);
        const int nloc = half * cols_half + col;
        const long long off = pix * P.C + n0 + nloc;
        uint4 xv[4], gv[4], sv[4];
        if (valid) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            xv[u] = __ldg(reinterpret_cast<const uint4*>(P.x + off) + u);
            if (P.bwd) {
              gv[u] = __ldg(reinterpret_cast<const uint4*>(P.g + off) + u);
              sv[u] = __ldg(reinterpret_cast<const uint4*>(P.s + off) + u);
            }
          }
        }
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float xf[8], o[8];
            unpack8(xv[u], xf);
            if (P.bwd) {
              float gf[8], sf[8];
              unpack8(gv[u], gf);
              unpack8(sv[u], sf);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float xq = P.x_is_act ? xf[e] * igdn_rcp(sf[e]) : xf[e];
                o[e] = gf[e] * sf[e] + xq * __uint_as_float(r[u * 8 + e]);
              }
              reinterpret_cast<uint4*>(P.out + off)[u] = pack8(o);
            } else {
              float sq[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                sq[e] = sqrtf(__uint_as_float(r[u * 8 + e]) + s_beta[nloc + u * 8 + e]);
                o[e] = xf[e] * sq[e];
              }
              reinterpret_cast<uint4*>(P.out + off)[u] = pack8(o);
              reinterpret_cast<uint4*>(P.s_out + off)[u] = pack8(sq);
            }
          }
        }
      }
      tcgen05_fence_before();
    };

    int it = 0;
    long long tile = cta_in_group;
    if (tile < num_tiles) stage1(tile, 0);
    for (; tile < num_tiles; tile += group_ctas, ++it) {
      const int buf = it & 1;
      const long long next = tile + group_ctas;
      if (P.abufs == 2 && next < num_tiles) stage1(next, buf ^ 1);
      mbar_wait(bar_acc + 8 * buf, (it >> 1) & 1);
      tcgen05_fence_after();
      // one A buffer: the MMAs of this tile have finished reading it (bar_acc), so the next tile's operand is built now and
      // its MMAs run while stage 3 drains this tile's accumulator (the accumulators stay double-buffered)
      if (P.abufs == 1 && next < num_tiles) stage1(next, buf ^ 1);
      stage3(tile, buf);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct IgdnTc {
  IgdnParams P;
  int smem_bytes, grid;
};

// gamma_op: bf16 [C][C] K-major operand (forward: gamma^T, backward: gamma)
int igdn_tc_create(const __nv_bfloat16* gamma_op, int C, long long pixels, int bwd, IgdnTc** out) {
  SGA_CHECK(C % 64 == 0 && C <= 512, "tcgen05 IGDN needs C %% 64 == 0 and <= 512 (got %d)", C);
  IgdnTc* c = new IgdnTc();
  memset(&c->P, 0, sizeof(c->P));
  IgdnParams& P = c->P;
  P.C = C; P.pixels = pixels; P.bwd = bwd;
  // largest channel tile (multiple of 64, <= 256, dividing C) whose operands fit in shared memory
  // Prefer covering all channels in one tile (every pixel tile is then read and its A operand built once), with a single
  // A buffer if that is what it takes (C = 256: 128 KB of gamma + 64 KB of operand); narrower tiles only as a last resort.
  int bn = 0, abufs = 2;
  for (int cand = 256; cand >= 64 && bn == 0; cand -= 64) {
    if (C % cand) continue;
    for (int ab = 2; ab >= 1; --ab) {
      const long long need = 1024LL + (C / 64) * (cand * 128LL) + static_cast<long long>(ab) * (C / 64) * 16384 + 1024 + 256;
      if (need <= 227 * 1024) { bn = cand; abufs = ab; break; }
    }
  }
  SGA_CHECK(bn > 0, "IGDN operands do not fit in shared memory for C=%d", C);
  P.block_n = bn; P.n_tiles_n = C / bn; P.abufs = abufs;
  c->smem_bytes = 1024 + (C / 64) * (bn * 128) + abufs * (C / 64) * 16384 + 1024 + 256;
  const uint64_t dims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(C)};
  const uint64_t strides[1] = {static_cast<uint64_t>(C) * 2};
  const uint32_t box[2] = {64, static_cast<uint32_t>(bn)};
  int st = tma_encode_tiled(&P.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gamma_op, dims, strides, box,
                            CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != SGA_OK) { delete c; return st; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (pixels + 127) / 128;
  long long groups = sms / P.n_tiles_n;
  if (groups > tiles) groups = tiles;
  if (groups < 1) groups = 1;
  c->grid = static_cast<int>(groups) * P.n_tiles_n;
  static bool attr_set = false;
  if (!attr_set) {
    SGA_CUDA(cudaFuncSetAttribute(k_igdn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  *out = c;
  return SGA_OK;
}

void igdn_tc_destroy(IgdnTc* c) { delete c; }

int igdn_tc_forward(const IgdnTc* c, const __nv_bfloat16* x, const float* beta, __nv_bfloat16* out, __nv_bfloat16* s_out,
                    cudaStream_t st) {
  if (c->P.pixels == 0) return SGA_OK;
  IgdnParams P = c->P;
  P.x = x; P.beta = beta; P.out = out; P.s_out = s_out;
  k_igdn_tc<<<c->grid, kIgdnThreads, c->smem_bytes, st>>>(P);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

int igdn_tc_backward(const IgdnTc* c, const __nv_bfloat16* g, const __nv_bfloat16* x, const __nv_bfloat16* s,
                     __nv_bfloat16* gx, cudaStream_t st, int x_is_act) {
  if (c->P.pixels == 0) return SGA_OK;
  IgdnParams P = c->P;
  P.x = x; P.g = g; P.s = s; P.out = gx; P.x_is_act = x_is_act;
  k_igdn_tc<<<c->grid, kIgdnThreads, c->smem_bytes, st>>>(P);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga
