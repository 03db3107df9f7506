// codec.cu -- entropy coding of the quantised latents: the bitstream side of the reference's mbt2018.py
// (`string = conditional_bottleneck.compress(y)`, `side_string = entropy_bottleneck.compress(z)`, mbt2018.py:84-85, packed
// at :211-221, decoded at :248-295).  The reference delegates this to tensorflow-compression's range-coding ops (C++
// kernels of a third-party package, not in /root/reference); sga.py itself only ESTIMATES the rate.  What is built here:
//
//   device  k_gauss_tables : quantised CDFs (16-bit precision) of  N(delta, s_i^2) * U(-1/2, 1/2)  on the integers for the 64
//                            scales s_i = exp(linspace(log .11, log 256, 64)) (sga.py:129 / mbt2018.py:74) x a grid of mean
//                            offsets delta in [-1/2, 1/2] (step 1/128 for small scales .. 1/8 for large ones): the
//                            Gaussian-conditional model evaluated where the estimate evaluates it
//           k_eb_tables    : the same for the factorised prior of z, one table per channel (sga.py:100-104)
//           k_bin_y/k_bin_z: per element, the table it is coded with and its symbol (the "binning")
//   host    a 32-bit range coder (carry-propagating byte-wise renormalisation, 16-bit frequencies) with an escape symbol per
//           table followed by an Elias-gamma code for values outside the table's support
//
// The stream is this repo's own: byte compatibility with tfc's PackedTensors / range coder cannot be checked offline
// (PARITY_ASSUMPTIONS.md P16).  What is guaranteed -- and tested -- is decode(encode(y_hat, z_hat)) == (y_hat, z_hat) bit
// for bit and an actual size within 1 % of the estimated one.
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "kernels.h"

namespace sga {

constexpr int kCdfBits = 16;
constexpr uint32_t kCdfTotal = 1u << kCdfBits;
constexpr int kScaleLevels = 64;        // SCALES_LEVELS (sga.py:26)
// Mean offsets delta = mu - rint(mu) are quantised on a grid whose step shrinks with the scale: what matters is delta /
// sigma (a 1/64 error at sigma = 0.11 moves a 50/50 split by 6 %, about 1 % in bits; measured +1.4 % total at 1/32).
__host__ __device__ inline int codec_offset_res(int i) { return i < 23 ? 128 : (i < 40 ? 32 : 8); }   // sigma < 1.9 / < 15.5
constexpr int kMaxTables = 64 * 129;
constexpr double kScaleMin = 0.11, kScaleMax = 256.0;   // SCALES_MIN / SCALES_MAX (sga.py:24-25)
constexpr int kEbRadius = 96;           // support of a factorised-prior table: median +- 96, escape beyond

__host__ __device__ inline double codec_scale(int i) {
  return exp(log(kScaleMin) + (log(kScaleMax) - log(kScaleMin)) * i / (kScaleLevels - 1));
}
// half-width of a Gaussian table: covers all but ~2^-16 of the mass on each side, at least 2
__host__ __device__ inline int codec_radius(double s) { return static_cast<int>(ceil(4.4 * s + 0.5)) + 2; }

// Integerise a pmf: symbols 0..L-1 plus the escape symbol L (tail mass); every symbol gets frequency >= 1, the sum is
// exactly 2^16 (the excess / deficit goes to the most probable symbol).  cdf has L + 2 entries.  One block per table,
// the serial parts run on thread 0 (tables have at most a few thousand entries and are built once per model).
__device__ void integerise_table(const double* __restrict__ pmf, int L, uint32_t* __restrict__ cdf) {
  if (threadIdx.x != 0) return;
  double tot = 0.0;
  for (int k = 0; k < L; ++k) tot += pmf[k];
  const double tail = fmax(1.0 - tot, 0.0);
  uint32_t sum = 0, best = 0;
  int arg = 0;
  for (int k = 0; k <= L; ++k) {
    const double p = (k < L) ? pmf[k] : tail;
    uint32_t f = static_cast<uint32_t>(p * (kCdfTotal - (L + 1)));     // leaves room for the +1 of every symbol
    f += 1;
    cdf[k] = f;                                                        // frequencies first
    sum += f;
    if (k < L && f > best) { best = f; arg = k; }
  }
  cdf[arg] += kCdfTotal - sum;                                         // sum <= 2^16 by construction
  uint32_t run = 0;
  for (int k = 0; k <= L; ++k) { const uint32_t f = cdf[k]; cdf[k] = run; run += f; }
  cdf[L + 1] = run;                                                    // == 2^16
}

__device__ __forceinline__ double std_cdf_d(double x) { return 0.5 * erfc(-0.70710678118654752440 * x); }

// table t = tab_base[scale i] + offset index j (j = 0 .. res_i, delta = j / res_i - 1/2); symbols k = -R..R map to
// indices 0..2R
__global__ void k_gauss_tables(const int* __restrict__ tab_base, const int* __restrict__ tab_off, uint32_t* __restrict__ cdf,
                               double* __restrict__ scratch, int scratch_stride) {
  const int t = blockIdx.x;
  int i = 0;
  while (i + 1 < kScaleLevels && tab_base[i + 1] <= t) ++i;
  const int j = t - tab_base[i];
  const double s = codec_scale(i);
  const double delta = static_cast<double>(j) / codec_offset_res(i) - 0.5;
  const int R = codec_radius(s), L = 2 * R + 1;
  double* pmf = scratch + static_cast<size_t>(t) * scratch_stride;
  for (int q = threadIdx.x; q < L; q += blockDim.x) {
    const double k = q - R;
    // the estimate's form (utils.py:98-101): evaluated on |k - delta| so that the upper tail never cancels
    const double v = fabs(k - delta);
    pmf[q] = std_cdf_d((0.5 - v) / s) - std_cdf_d((-0.5 - v) / s);
  }
  __syncthreads();
  integerise_table(pmf, L, cdf + tab_off[t]);
}

// factorised prior: table per channel centred on `center[c]` (the integer nearest to the mode of the density)
struct EbP { float p[44]; };
__device__ double eb_logit_d(const float* __restrict__ pk, int C, int c, double u) {
  double h[3];
  for (int r = 0; r < 3; ++r) {
    h[r] = static_cast<double>(pk[r * C + c]) * u + pk[(3 + r) * C + c];
    h[r] += static_cast<double>(pk[(6 + r) * C + c]) * tanh(h[r]);
  }
  for (int layer = 0; layer < 2; ++layer) {
    const int mo = 9 + 15 * layer, bo = 18 + 15 * layer, fo = 21 + 15 * layer;
    double h2[3];
    for (int r = 0; r < 3; ++r) {
      double a = pk[(bo + r) * C + c];
      for (int k = 0; k < 3; ++k) a += static_cast<double>(pk[(mo + r * 3 + k) * C + c]) * h[k];
      h2[r] = a + static_cast<double>(pk[(fo + r) * C + c]) * tanh(a);
    }
    for (int r = 0; r < 3; ++r) h[r] = h2[r];
  }
  double a = pk[42 * C + c];
  for (int k = 0; k < 3; ++k) a += static_cast<double>(pk[(39 + k) * C + c]) * h[k];
  return a;
}
__device__ double eb_pmf_d(const float* __restrict__ pk, int C, int c, double k) {
  const double lo = eb_logit_d(pk, C, c, k - 0.5), up = eb_logit_d(pk, C, c, k + 0.5);
  const double sg = (lo + up > 0.0) ? -1.0 : ((lo + up < 0.0) ? 1.0 : 0.0);
  const double a = 1.0 / (1.0 + exp(-sg * up)), b = 1.0 / (1.0 + exp(-sg * lo));
  return fabs(a - b);
}
__global__ void k_eb_tables(const float* __restrict__ pk, int C, int* __restrict__ center, uint32_t* __restrict__ cdf,
                            double* __restrict__ scratch) {
  const int c = blockIdx.x;
  constexpr int L = 2 * kEbRadius + 1;
  __shared__ int s_center;
  if (threadIdx.x == 0) {
    // median of the prior: bisection of logit(u) = 0 on [-1e4, 1e4] (the logit is increasing in u)
    double a = -1e4, b = 1e4;
    for (int it = 0; it < 80; ++it) {
      const double m = 0.5 * (a + b);
      if (eb_logit_d(pk, C, c, m) < 0.0) a = m; else b = m;
    }
    s_center = static_cast<int>(rint(0.5 * (a + b)));
    center[c] = s_center;
  }
  __syncthreads();
  double* pmf = scratch + static_cast<size_t>(c) * L;
  for (int q = threadIdx.x; q < L; q += blockDim.x) pmf[q] = eb_pmf_d(pk, C, c, static_cast<double>(s_center + q - kEbRadius));
  __syncthreads();
  integerise_table(pmf, L, cdf + static_cast<size_t>(c) * (L + 2));
}

// y: table and symbol of every element.  mode 0 (SGA result, sga.py:240): y_hat = rint(y) is an integer, coded relative to
// base = rint(mu) under the table with offset delta = mu - base; mode 1 (mbt2018.py baseline, tfc's mean-centred
// quantisation): symbol = rint(y - mu), table offset 0, y_hat = symbol + mu.
__global__ void k_bin_y(const float* __restrict__ y, const float* __restrict__ h2, int N, int Hy, int Wy, int Hh, int Wh,
                        int C, int mode, const int* __restrict__ tab_base, int* __restrict__ tab, int* __restrict__ sym,
                        float* __restrict__ base_out, float* __restrict__ yhat_out) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long numel = static_cast<long long>(N) * Hy * Wy * C;
  if (e >= numel) return;
  const int c = static_cast<int>(e % C);
  long long pix = e / C;
  const int j = static_cast<int>(pix % Wy); pix /= Wy;
  const int i = static_cast<int>(pix % Hy);
  const int n = static_cast<int>(pix / Hy);
  const long long hp = ((static_cast<long long>(n) * Hh + i) * Wh + j) * 2 * C;
  const float mu = h2[hp + c];
  const float sigma = fmaxf(expf(h2[hp + C + c]), 0.11f);
  // nearest scale in the log domain
  const float step = (logf(256.0f) - logf(0.11f)) / (kScaleLevels - 1);
  int si = static_cast<int>(rintf((logf(sigma) - logf(0.11f)) / step));
  si = min(max(si, 0), kScaleLevels - 1);
  int k = 0, oj;
  float base;
  if (mode == 0) {
    base = rintf(mu);
    oj = static_cast<int>(rintf((mu - base) * static_cast<float>(codec_offset_res(si))));   // in [-res/2, res/2]
    if (y) {
      k = static_cast<int>(rintf(y[e]) - base);
      if (yhat_out) yhat_out[e] = rintf(y[e]);
    }
  } else {
    base = mu;
    oj = 0;
    if (y) {
      k = static_cast<int>(rintf(y[e] - mu));
      if (yhat_out) yhat_out[e] = static_cast<float>(k) + mu;
    }
  }
  tab[e] = tab_base[si] + oj + codec_offset_res(si) / 2;
  if (y) sym[e] = k;
  base_out[e] = base;
}

__global__ void k_bin_z(const float* __restrict__ z, long long numel, int C, const int* __restrict__ center,
                        int* __restrict__ tab, int* __restrict__ sym) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= numel) return;
  const int c = static_cast<int>(e % C);
  tab[e] = c;
  sym[e] = static_cast<int>(rintf(z[e])) - center[c];
}

// ------------------------------------------------------------------------------------------------------------------
// host: table set + range coder
// ------------------------------------------------------------------------------------------------------------------
struct TableSet {
  std::vector<int> off;        // start of table t in cdf
  std::vector<int> len;        // L_t (regular symbols); cdf holds L_t + 2 entries
  std::vector<int> sym_min;    // value of symbol index 0
  std::vector<uint32_t> cdf;
};

}  // namespace sga

using namespace sga;

struct sga_codec {
  TableSet y, z;
  int C = 0;
  int* d_center = nullptr;     // device: centre of every channel's z table
  int* d_tab_base = nullptr;   // device: first y table of every scale
};

namespace sga {

struct RcEncoder {
  std::vector<uint8_t>& out;
  uint64_t low = 0;
  uint32_t range = 0xFFFFFFFFu;
  uint8_t cache = 0;
  uint64_t cache_size = 1;
  explicit RcEncoder(std::vector<uint8_t>& o) : out(o) {}
  void shift_low() {
    if (static_cast<uint32_t>(low) < 0xFF000000u || (low >> 32) != 0) {
      uint8_t carry = static_cast<uint8_t>(low >> 32);
      uint8_t temp = cache;
      do {
        out.push_back(static_cast<uint8_t>(temp + carry));
        temp = 0xFF;
      } while (--cache_size != 0);
      cache = static_cast<uint8_t>(low >> 24);
    }
    ++cache_size;
    low = (low & 0x00FFFFFFu) << 8;
  }
  void encode(uint32_t cum, uint32_t freq, int bits) {
    range >>= bits;
    low += static_cast<uint64_t>(cum) * range;
    range *= freq;
    while (range < (1u << 24)) { range <<= 8; shift_low(); }
  }
  void finish() { for (int i = 0; i < 5; ++i) shift_low(); }
};

struct RcDecoder {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t range = 0xFFFFFFFFu, code = 0;
  RcDecoder(const uint8_t* b, size_t n) : p(b), end(b + n) {
    next();                      // the encoder's first byte is the initial (zero) cache
    for (int i = 0; i < 4; ++i) code = (code << 8) | next();
  }
  uint8_t next() { return p < end ? *p++ : 0; }
  uint32_t target(int bits) { range >>= bits; const uint32_t v = code / range; return v < (1u << bits) ? v : (1u << bits) - 1; }
  void consume(uint32_t cum, uint32_t freq) {
    code -= cum * range;
    range *= freq;
    while (range < (1u << 24)) { range <<= 8; code = (code << 8) | next(); }
  }
};

// escape values: sign bit + Elias-gamma of the distance beyond the table's support, as equiprobable bits
static void put_bits(RcEncoder& e, uint32_t v, int n) { for (int b = n - 1; b >= 0; --b) e.encode((v >> b) & 1u, 1, 1); }
static uint32_t get_bits(RcDecoder& d, int n) {
  uint32_t v = 0;
  for (int b = 0; b < n; ++b) { const uint32_t t = d.target(1); d.consume(t, 1); v = (v << 1) | t; }
  return v;
}
static void put_gamma(RcEncoder& e, uint32_t v) {   // v >= 1
  int nb = 0;
  while ((v >> (nb + 1)) != 0) ++nb;
  put_bits(e, 0, nb);
  put_bits(e, v, nb + 1);
}
static uint32_t get_gamma(RcDecoder& d) {
  int nb = 0;
  while (nb < 32) { const uint32_t t = d.target(1); d.consume(t, 1); if (t) break; ++nb; }
  uint32_t v = 1;
  for (int b = 0; b < nb; ++b) { const uint32_t t = d.target(1); d.consume(t, 1); v = (v << 1) | t; }
  return v;
}

static int rc_encode(const TableSet& T, long long n, const int* tab, const int* sym, std::vector<uint8_t>& out) {
  RcEncoder e(out);
  for (long long i = 0; i < n; ++i) {
    const int t = tab[i];
    if (t < 0 || t >= static_cast<int>(T.off.size())) { set_error("range coder: table index %d out of range", t); return SGA_ERR_INVALID; }
    const uint32_t* cdf = T.cdf.data() + T.off[t];
    const int L = T.len[t];
    const long long q = static_cast<long long>(sym[i]) - T.sym_min[t];
    if (q >= 0 && q < L) {
      e.encode(cdf[q], cdf[q + 1] - cdf[q], kCdfBits);
    } else {
      e.encode(cdf[L], cdf[L + 1] - cdf[L], kCdfBits);
      const bool neg = q < 0;
      put_bits(e, neg ? 1u : 0u, 1);
      put_gamma(e, static_cast<uint32_t>(neg ? -q : q - L + 1));
    }
  }
  e.finish();
  return SGA_OK;
}

static int rc_decode(const TableSet& T, long long n, const int* tab, const uint8_t* bytes, size_t nbytes, int* sym) {
  RcDecoder d(bytes, nbytes);
  for (long long i = 0; i < n; ++i) {
    const int t = tab[i];
    if (t < 0 || t >= static_cast<int>(T.off.size())) { set_error("range coder: table index %d out of range", t); return SGA_ERR_INVALID; }
    const uint32_t* cdf = T.cdf.data() + T.off[t];
    const int L = T.len[t];
    const uint32_t v = d.target(kCdfBits);
    int lo = 0, hi = L + 1;                         // largest q with cdf[q] <= v
    while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (cdf[m] <= v) lo = m; else hi = m; }
    d.consume(cdf[lo], cdf[lo + 1] - cdf[lo]);
    if (lo < L) {
      sym[i] = lo + T.sym_min[t];
    } else {
      const bool neg = get_bits(d, 1) != 0;
      const long long g = get_gamma(d);
      sym[i] = static_cast<int>(neg ? T.sym_min[t] - g : T.sym_min[t] + L - 1 + g);
    }
  }
  return SGA_OK;
}

// pk: packed factorised-prior parameters of the context (elementwise.cu k_eb_pack)
int codec_build_tables(const float* eb_pack, int C, sga_codec* cd, cudaStream_t st) {
  // ---- Gaussian tables
  int tab_base[kScaleLevels + 1];
  tab_base[0] = 0;
  for (int i = 0; i < kScaleLevels; ++i) tab_base[i + 1] = tab_base[i] + codec_offset_res(i) + 1;
  const int nt = tab_base[kScaleLevels];
  TableSet& Y = cd->y;
  Y.off.resize(nt); Y.len.resize(nt); Y.sym_min.resize(nt);
  int total = 0, maxL = 0;
  for (int i = 0; i < kScaleLevels; ++i)
    for (int t = tab_base[i]; t < tab_base[i + 1]; ++t) {
      const int R = codec_radius(codec_scale(i)), L = 2 * R + 1;
      Y.off[t] = total; Y.len[t] = L; Y.sym_min[t] = -R;
      total += L + 2;
      if (L > maxL) maxL = L;
    }
  Y.cdf.resize(total);
  int* d_off = nullptr; uint32_t* d_cdf = nullptr; double* d_scr = nullptr;
  SGA_CUDA(cudaMalloc(&cd->d_tab_base, (kScaleLevels + 1) * sizeof(int)));
  SGA_CUDA(cudaMemcpyAsync(cd->d_tab_base, tab_base, (kScaleLevels + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  SGA_CUDA(cudaMalloc(&d_off, nt * sizeof(int)));
  SGA_CUDA(cudaMalloc(&d_cdf, static_cast<size_t>(total) * sizeof(uint32_t)));
  SGA_CUDA(cudaMalloc(&d_scr, static_cast<size_t>(nt) * maxL * sizeof(double)));
  SGA_CUDA(cudaMemcpyAsync(d_off, Y.off.data(), nt * sizeof(int), cudaMemcpyHostToDevice, st));
  k_gauss_tables<<<nt, 128, 0, st>>>(cd->d_tab_base, d_off, d_cdf, d_scr, maxL);
  SGA_LAUNCH_CHECK();
  SGA_CUDA(cudaMemcpyAsync(Y.cdf.data(), d_cdf, static_cast<size_t>(total) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  // ---- factorised-prior tables
  constexpr int Lz = 2 * kEbRadius + 1;
  TableSet& Z = cd->z;
  Z.off.resize(C); Z.len.assign(C, Lz); Z.sym_min.assign(C, -kEbRadius);     // symbols are taken relative to the centre
  for (int c = 0; c < C; ++c) Z.off[c] = c * (Lz + 2);
  Z.cdf.resize(static_cast<size_t>(C) * (Lz + 2));
  uint32_t* d_zcdf = nullptr; double* d_zscr = nullptr;
  SGA_CUDA(cudaMalloc(&cd->d_center, C * sizeof(int)));
  SGA_CUDA(cudaMalloc(&d_zcdf, Z.cdf.size() * sizeof(uint32_t)));
  SGA_CUDA(cudaMalloc(&d_zscr, static_cast<size_t>(C) * Lz * sizeof(double)));
  k_eb_tables<<<C, 128, 0, st>>>(eb_pack, C, cd->d_center, d_zcdf, d_zscr);
  SGA_LAUNCH_CHECK();
  SGA_CUDA(cudaMemcpyAsync(Z.cdf.data(), d_zcdf, Z.cdf.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SGA_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_off); cudaFree(d_cdf); cudaFree(d_scr); cudaFree(d_zcdf); cudaFree(d_zscr);
  cd->C = C;
  return SGA_OK;
}

sga_codec* codec_new() { return new sga_codec(); }
int codec_num_channels(const sga_codec* cd) { return cd ? cd->C : 0; }

int launch_bin_y(const sga_codec* cd, const float* y, const float* h2, int N, int Hy, int Wy, int Hh, int Wh, int C, int mode,
                 int* tab, int* sym, float* base, float* yhat, cudaStream_t st) {
  const long long numel = static_cast<long long>(N) * Hy * Wy * C;
  if (numel <= 0) return SGA_OK;
  SGA_CHECK(cd->d_tab_base != nullptr, "codec has no device tables (host-only codec)");
  k_bin_y<<<static_cast<unsigned>((numel + 255) / 256), 256, 0, st>>>(y, h2, N, Hy, Wy, Hh, Wh, C, mode, cd->d_tab_base, tab, sym,
                                                                       base, yhat);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

int launch_bin_z(const float* z, long long numel, int C, const sga_codec* cd, int* tab, int* sym, cudaStream_t st) {
  if (numel <= 0) return SGA_OK;
  k_bin_z<<<static_cast<unsigned>((numel + 255) / 256), 256, 0, st>>>(z, numel, C, cd->d_center, tab, sym);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}

}  // namespace sga

extern "C" {

/* centre of every channel's z table (host copy): z_hat = centre[c] + symbol */
int sga_codec_z_centers(const sga_codec* cd, int32_t* out, int32_t n) {
  if (!cd || !out) { set_error("null argument"); return SGA_ERR_INVALID; }
  SGA_CHECK(n == cd->C && cd->d_center != nullptr, "codec has %d channels (host-only codecs have no centres)", cd->C);
  SGA_CUDA(cudaMemcpy(out, cd->d_center, n * sizeof(int), cudaMemcpyDeviceToHost));
  return SGA_OK;
}

int sga_codec_destroy(sga_codec* cd) {
  if (!cd) return SGA_OK;
  if (cd->d_center) cudaFree(cd->d_center);
  if (cd->d_tab_base) cudaFree(cd->d_tab_base);
  delete cd;
  return SGA_OK;
}

// Host-only table set (tests of the coder without a GPU; also what a decoder-only build would load): `which` 0 = y tables,
// 1 = z tables.  freqs: concatenated per-table frequencies of the L_t regular symbols followed by the escape symbol; every
// table must sum to 2^16 with all entries >= 1.
int sga_codec_create_host(int32_t n_tables_y, const int32_t* len_y, const int32_t* min_y, const uint32_t* freq_y,
                          int32_t n_tables_z, const int32_t* len_z, const int32_t* min_z, const uint32_t* freq_z,
                          sga_codec** out) {
  if (!out) { set_error("null argument"); return SGA_ERR_INVALID; }
  sga_codec* cd = new sga_codec();
  for (int w = 0; w < 2; ++w) {
    TableSet& T = w ? cd->z : cd->y;
    const int nt = w ? n_tables_z : n_tables_y;
    const int32_t* len = w ? len_z : len_y;
    const int32_t* mn = w ? min_z : min_y;
    const uint32_t* fr = w ? freq_z : freq_y;
    long long pos = 0;
    for (int t = 0; t < nt; ++t) {
      T.off.push_back(static_cast<int>(T.cdf.size()));
      T.len.push_back(len[t]);
      T.sym_min.push_back(mn[t]);
      uint32_t run = 0;
      for (int k = 0; k <= len[t]; ++k) {
        if (fr[pos + k] == 0) { delete cd; set_error("table %d: zero frequency", t); return SGA_ERR_INVALID; }
        T.cdf.push_back(run);
        run += fr[pos + k];
      }
      T.cdf.push_back(run);
      if (run != kCdfTotal) { delete cd; set_error("table %d sums to %u, expected %u", t, run, kCdfTotal); return SGA_ERR_INVALID; }
      pos += len[t] + 1;
    }
  }
  *out = cd;
  return SGA_OK;
}

// HOST arrays.  Returns the number of bytes written through *nbytes; fails if `capacity` is too small.
int sga_rc_encode(const sga_codec* cd, int32_t which, int64_t n, const int32_t* tab, const int32_t* sym, uint8_t* out,
                  uint64_t capacity, uint64_t* nbytes) {
  if (!cd || !nbytes || (n > 0 && (!tab || !sym))) { set_error("null argument"); return SGA_ERR_INVALID; }
  std::vector<uint8_t> buf;
  buf.reserve(static_cast<size_t>(n / 2 + 64));
  SGA_TRY(rc_encode(which ? cd->z : cd->y, n, tab, sym, buf));
  *nbytes = buf.size();
  SGA_CHECK(out != nullptr && buf.size() <= capacity, "range coder: output needs %zu bytes, capacity %llu", buf.size(),
            static_cast<unsigned long long>(capacity));
  memcpy(out, buf.data(), buf.size());
  return SGA_OK;
}

int sga_rc_decode(const sga_codec* cd, int32_t which, int64_t n, const int32_t* tab, const uint8_t* bytes, uint64_t nbytes,
                  int32_t* sym_out) {
  if (!cd || (n > 0 && (!tab || !sym_out)) || (nbytes > 0 && !bytes)) { set_error("null argument"); return SGA_ERR_INVALID; }
  return rc_decode(which ? cd->z : cd->y, n, tab, bytes, static_cast<size_t>(nbytes), sym_out);
}

}  // extern "C"
