/* sga_b200.h -- C ABI of libsga_b200.so: the Blackwell (sm_100a) implementation of the
 * Stochastic-Gumbel-Annealing hot path of mandt-lab/improving-inference-for-neural-image-compression.
 *
 * The reference has no FFI; the seam this library replaces is the per-iteration
 *     sess.run([rd_gradients, rd_loss, train_mse, train_bpp, psnr], feed {y, z, x, T})   sga.py:212-214
 *     y, z = adam_optimizer.update([y, z], grads)                                         sga.py:215 (adam.py:20-59)
 * plus the one-off init  sess.run([y_init, z_init])  (sga.py:207) and the final hard-rounded
 * evaluation (sga.py:240-247).  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; tensors are NHWC, contiguous, float32
 *  - all work is enqueued on the cudaStream_t passed in (as void*); no hidden synchronisation,
 *    no allocation after sga_create + sga_bind_weights
 *  - every function returns 0 on success, a negative sga_status otherwise; sga_last_error()
 *    returns a thread-local description.  Nothing throws.
 *  - one context per GPU; a context is not thread-safe; contexts are independent
 */
#ifndef SGA_B200_H_
#define SGA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SGA_OK = 0,
  SGA_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  SGA_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed */
  SGA_ERR_STATE = -3,     /* e.g. weights not bound */
  SGA_ERR_NOMEM = -4
} sga_status;

/* arithmetic of the (de)conv stacks inside the loop */
typedef enum {
  SGA_OPERAND_F32 = 0,   /* fp32 SIMT kernels: init (sga.py:207), final eval (sga.py:244-247), checking */
  SGA_OPERAND_BF16 = 1   /* bf16 operands, fp32 accumulate in TMEM: tcgen05 implicit-GEMM kernels */
} sga_operand;

/* which per-iteration graph the context evaluates */
typedef enum {
  SGA_VARIANT_SGA = 0,        /* sga.py: Gumbel soft round of y and z, factorised-prior likelihood of z */
  SGA_VARIANT_BITS_BACK = 1   /* bb_sga.py: z_tilde = eps*exp(logvar/2)+mean, prior density, bits-back term; h_a has 2C outputs */
} sga_variant;

typedef struct {
  int32_t n_images;        /* images resident on this GPU (the shard) */
  int32_t height, width;   /* image size (sga.py:51) */
  int32_t num_filters;     /* --num_filters C (sga.py:70-73) */
  int32_t n_ref;           /* size of the REFERENCE batch these images belong to: both loss terms are
                              means over it (sga.py:147,150), so every gradient carries 1/n_ref */
  int32_t operand;         /* sga_operand used by sga_step / sga_grad */
  float lmbda;             /* --lambda (sga.py:157-163); <= 0 means rate only */
  float clip_eps;          /* epsilon of the atanh clip, 1e-5 (sga.py:30) */
  double lr, beta1, beta2, adam_eps;  /* adam.py:9-10 ; sga.py:190 uses lr = 0.005.  Python floats in the reference: kept
                              as doubles so that lr_t and float32(1 - beta) are formed exactly as numpy forms them */
  uint64_t seed;           /* key of the Philox4x32-10 Gumbel noise stream */
  int32_t first_image;     /* index of this shard's first image inside the reference batch: the noise counter is the
                              element index inside the WHOLE batch, so results do not depend on the sharding */
  int32_t variant;         /* sga_variant: which reference script's graph (0 = sga.py, 1 = bb_sga.py) */
} sga_config;

/* Weights in the reference's variable layout (float32 device pointers; kernels (kh,kw,Cin,Cout),
 * gamma (in,out)); index = layer number inside the transform (nn_models.py). */
typedef struct {
  const float* ga_kernel[4]; const float* ga_bias[4]; const float* ga_beta[3]; const float* ga_gamma[3];
  const float* gs_kernel[4]; const float* gs_bias[4]; const float* gs_beta[3]; const float* gs_gamma[3];
  const float* ha_kernel[3]; const float* ha_bias[2];
  const float* hs_kernel[3]; const float* hs_bias[3];
  const float* eb_matrix[4]; const float* eb_bias[4]; const float* eb_factor[3];
} sga_weights;

typedef struct sga_ctx sga_ctx;

/* stats_out layout (floats): [0] rd_loss [1] train_mse [2] train_bpp, then for image n (0-based):
 *   [3+5n] psnr  [4+5n] y_bpp  [5+5n] z_bpp  [6+5n] sse (float reconstruction, sum over h,w,c)
 *   [7+5n] mse of the rounded reconstruction (0..255 scale)                       -> 3 + 5*n_images floats
 * rd_loss / train_mse / train_bpp are this shard's contribution: sum over local images / n_ref. */
#define SGA_STATS_HEADER 3
#define SGA_STATS_PER_IMAGE 5

const char* sga_last_error(void);
int sga_version(void);

int sga_create(const sga_config* cfg, sga_ctx** out);
int sga_destroy(sga_ctx* ctx);
/* bytes of device workspace owned by the context */
size_t sga_workspace_bytes(const sga_ctx* ctx);
/* latent shapes: y is [n, hy, wy, C], z is [n, hz, wz, C] */
int sga_latent_shape(const sga_ctx* ctx, int32_t* hy, int32_t* wy, int32_t* hz, int32_t* wz);

/* Copies / re-lays-out the weights into the context (transposed + bf16 copies for the tcgen05 path). */
int sga_bind_weights(sga_ctx* ctx, const sga_weights* w, void* stream);

/* y_init = g_a(x), z_init = h_a(y_init)   (sga.py:77-78,207; nn_models.py:5-36,73-103).  Always fp32. */
int sga_encode_init(sga_ctx* ctx, const float* x, float* y_out, float* z_out, void* stream);

/* One evaluation of the reference's per-iteration graph (sga.py:212-214): gradients of rd_loss w.r.t.
 * y and z and the scalar statistics.  `step` selects the noise counter; if noise_y/noise_z are non-null
 * they hold injected uniforms in (0,1), shape [2][numel] (DOWN logits first), and replace the Philox draw. */
int sga_grad(sga_ctx* ctx, const float* x, const float* y, const float* z, float temperature, int32_t step,
             const float* noise_y, const float* noise_z, float* grad_y, float* grad_z, float* stats_out,
             void* stream);

/* sga_grad followed by the Adam update of adam.py:36-57 on y, z (in place) with moments m_*, v_*.
 * `step` is also adam's `iterations` (0-based). */
int sga_step(sga_ctx* ctx, const float* x, float* y, float* z, float* m_y, float* v_y, float* m_z, float* v_z,
             float temperature, int32_t step, const float* noise_y, const float* noise_z, float* stats_out,
             void* stream);

/* The loop of sga.py:210-215 for n_steps consecutive iterations, device-resident: iteration k (0-based inside the call)
 * is sga_step with temperature temperatures[k] (HOST array of n_steps floats, the caller's annealing schedule,
 * sga.py:211) and step = first_step + k, Philox noise only.  On bf16 contexts every step is issued as a two-stream
 * schedule (hyper path beside the synthesis path) from inside the library, nothing is synchronised with the host; with
 * the environment variable SGA_GRAPH=1 the step is captured once per set of tensors as a CUDA graph and replayed with the
 * per-iteration scalars read from a device table.  stats_out (may be NULL) receives the
 * statistics of the LAST iteration.  The caller's stream is honoured (work is ordered after everything already enqueued on
 * it and later work on it waits for the loop); the default stream is legal. */
int sga_loop(sga_ctx* ctx, const float* x, float* y, float* z, float* m_y, float* v_y, float* m_z, float* v_z,
             const float* temperatures, int32_t first_step, int32_t n_steps, float* stats_out, void* stream);

/* Final evaluation with hard-rounded latents (sga.py:240-247), always fp32: round-half-even of y, z fed as
 * y_tilde / z_tilde.  stats_out as above (rd_loss etc. computed on the rounded latents). */
int sga_eval_hard(sga_ctx* ctx, const float* x, const float* y, const float* z, float* stats_out, void* stream);

/* Everything downstream of GIVEN (y_tilde, z_tilde), always fp32, nothing rounded: the eval tensors of mbt2018.py:171-185
 * for the baseline's mean-centred quantisation; with x = NULL only the reconstruction is produced (the decoder's synthesis,
 * mbt2018.py:275-283; read it with sga_reconstruction). */
int sga_eval_tilde(sga_ctx* ctx, const float* x, const float* y_tilde, const float* z_tilde, float* stats_out, void* stream);

/* ---- entropy coding of the quantised latents (reference mbt2018.py:84-85 `string` / `side_string`, :211-221, :248-295;
 * the reference calls tensorflow-compression's range-coding ops there).  A codec holds the quantised CDF tables: for y the
 * Gaussian-conditional model on the 64 scales of sga.py:129 x 33 mean offsets, for z one table per channel from the
 * factorised prior; both are built on the device.  Binning (table index + symbol per element) runs on the device, the
 * range coder itself on the host. */
typedef struct sga_codec sga_codec;
int sga_codec_create(sga_ctx* ctx, sga_codec** out, void* stream);
int sga_codec_destroy(sga_codec* codec);
/* Host-only table set (coder tests, decoder-only use).  freq_*: per table len+1 frequencies (regular symbols, then the
 * escape symbol), every table summing to 65536 with no zero entry; min_*: value of symbol index 0. */
int sga_codec_create_host(int32_t n_tables_y, const int32_t* len_y, const int32_t* min_y, const uint32_t* freq_y,
                          int32_t n_tables_z, const int32_t* len_z, const int32_t* min_z, const uint32_t* freq_z,
                          sga_codec** out);
int sga_codec_z_centers(const sga_codec* codec, int32_t* out, int32_t n_channels);
/* z: hyper-latents, rounded inside (z_hat = rint(z)); y: latents or NULL (decoder).  mode 0: y_hat = rint(y) (sga.py:240),
 * mode 1: y_hat = rint(y - mu) + mu (mbt2018.py baseline).  Device int32 / float arrays with one entry per latent element;
 * y_hat = base_y + sym_y, z_hat = centre[channel] + sym_z.  yhat, tab_z, sym_z may be NULL. */
int sga_codec_bin(sga_ctx* ctx, const sga_codec* codec, const float* y, const float* z, int32_t mode, int32_t* tab_y,
                  int32_t* sym_y, float* base_y, float* yhat, int32_t* tab_z, int32_t* sym_z, void* stream);
/* The range coder, HOST arrays.  which: 0 = y tables, 1 = z tables.  *nbytes receives the stream length (also when the
 * capacity is too small, so the caller can retry). */
int sga_rc_encode(const sga_codec* codec, int32_t which, int64_t n, const int32_t* tab, const int32_t* sym, uint8_t* out,
                  uint64_t capacity, uint64_t* nbytes);
int sga_rc_decode(const sga_codec* codec, int32_t which, int64_t n, const int32_t* tab, const uint8_t* bytes,
                  uint64_t nbytes, int32_t* sym_out);

/* Copies the cropped reconstruction x_tilde [n,H,W,3] (float, un-clipped) of the most recent fp32 forward
 * (sga_eval_hard, or sga_grad/sga_step with operand f32) -- used for MS-SSIM (sga.py:175) and PNG dumps. */
int sga_reconstruction(sga_ctx* ctx, float* xt_out, void* stream);

/* ---- bits-back variant (reference bb_sga.py) -- contexts created with variant = SGA_VARIANT_BITS_BACK -------------
 * The weights struct is the same; ha_kernel[2] is (5,5,C,2C) (bb_sga.py:69) and eb_* hold the BMSHJ2018Prior
 * variables (bb_sga.py:63; same per-channel network, learned_prior.py:43-66).  sga_encode_init is called with
 * z_out = NULL.  Statistics: [0] rd_loss [1] train_mse [2] train_bpp (= mean of y_bpp + z_bpp - bpp_back), then per
 * image [3+6n] psnr [4+6n] y_bpp [5+6n] z_bpp [6+6n] sse [7+6n] rounded mse [8+6n] bpp_back  -> 3 + 6*n_images floats. */
#define SGA_BB_STATS_PER_IMAGE 6

/* (z_mean_init, z_logvar_init) = split(h_a(y_tilde))   bb_sga.py:99, 204, 247.  Always fp32. */
int sga_bb_hyper_init(sga_ctx* ctx, const float* y_tilde, float* z_mean_out, float* z_logvar_out, void* stream);

/* One evaluation of bb_sga.py's per-iteration graph: gradients w.r.t. [y, z_mean, z_logvar] of rd_loss
 * (bb_sga.py:156, 212-216), or -- rate_only != 0 -- w.r.t. [z_mean, z_logvar] of train_bpp with `y` holding the fixed
 * y_tilde = round(y) (bb_sga.py:157, 250-252; x may be NULL, grad_y is not written).  eps_z: injected N(0,1) noise
 * shaped like z_mean, or NULL -> Philox(seed, step, tensor 3) + Box-Muller. */
int sga_bb_grad(sga_ctx* ctx, const float* x, const float* y, const float* z_mean, const float* z_logvar,
                float temperature, int32_t step, const float* noise_y, const float* eps_z, int32_t rate_only,
                float* grad_y, float* grad_z_mean, float* grad_z_logvar, float* stats_out, void* stream);

/* sga_bb_grad followed by adam.py's update (one optimiser over the tensors, bb_sga.py:208,217 / :248,253) with
 * learning rate `lr` (0.005 in the R-D loop, 0.003 in the rate loop; bb_sga.py:186,192).  `step` keys the noise,
 * `adam_iterations` is the optimiser's 0-based iteration count (it restarts with the second loop). */
int sga_bb_step(sga_ctx* ctx, const float* x, float* y, float* z_mean, float* z_logvar, float* m_y, float* v_y,
                float* m_zm, float* v_zm, float* m_zl, float* v_zl, float temperature, int32_t step,
                int32_t adam_iterations, const float* noise_y, const float* eps_z, int32_t rate_only, double lr,
                float* stats_out, void* stream);

/* eval_tensors of bb_sga.py:176-177 for y_tilde = round(y) (bb_sga.py:271-272), always fp32. */
int sga_bb_eval(sga_ctx* ctx, const float* x, const float* y_tilde, const float* z_mean, const float* z_logvar,
                int32_t step, const float* eps_z, float* stats_out, void* stream);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------
 * sga_launch_count: kernels launched by this library in the calling thread since load.
 * sga_profile_enable(ctx, 1): bracket every launch group of sga_grad/sga_step with CUDA events on the launching
 * stream; sga_profile_read synchronises the recorded events, adds the elapsed milliseconds into ms[label] and the
 * number of groups into count[label] (label < n_labels, see SGA_PROF_*), and clears the record. */
#define SGA_PROF_ELEMENTWISE 0
#define SGA_PROF_IGDN_FWD 1
#define SGA_PROF_IGDN_BWD 2
#define SGA_PROF_L3_FWD 3
#define SGA_PROF_L3_DGRAD 4
#define SGA_PROF_CONV_HS 8      /* +0..2 forward layer i, +3..5 dgrad of layer i */
#define SGA_PROF_CONV_GS 16     /* +0..2 forward layer i, +3..5 dgrad of layer i */
#define SGA_PROF_LABELS 24
long long sga_launch_count(void);
int sga_profile_enable(sga_ctx* ctx, int32_t enable);
int sga_profile_read(sga_ctx* ctx, float* ms, int32_t* count, int32_t n_labels);

/* ---- operator-level entry points (used by the parity tests; same kernels the step uses) -------------- */

/* a2: Gumbel soft rounding (sga.py:88-98,111-121).  u = [2][numel] uniforms or NULL -> Philox(seed, step, tensor_id). */
int sga_op_soft_round(const float* v, const float* u, int64_t numel, float temperature, float clip_eps,
                      uint64_t seed, int32_t step, int32_t tensor_id, float* v_tilde, float* dvt_dv, void* stream);
/* a3: factorised-prior likelihood (tfc EntropyBottleneck._likelihood; learned_prior.py:78-121,296-347).
 * params: matrix_i/bias_i/factor_i device pointers in the reference layout. p_out and dp_out may be NULL. */
int sga_op_eb_likelihood(const float* z_tilde, int64_t numel, int32_t channels, const float* const matrix[4],
                         const float* const bias[4], const float* const factor[3], float* p_out, float* dp_out,
                         void* stream);
/* a6: box-convolved Gaussian likelihood (utils.py:80-102) with scale bound 0.11 and its partials. */
int sga_op_gauss_likelihood(const float* y_tilde, const float* mu, const float* log_sigma, int64_t numel,
                            float* p_out, float* dp_dy, float* dp_dsigma_bounded, void* stream);
/* a12: Adam (adam.py:36-57), float32 state, lr_t computed on the host in double. */
int sga_op_adam(float* p, float* m, float* v, const float* g, int64_t numel, double lr, double beta1, double beta2,
                double eps, int32_t iterations, void* stream);
/* a4/a7/a14: one (de)conv of nn_models.py. kind: 0 = up2 (corr=False, strides_up=2), 1 = dgrad of up2,
 * 2 = stride-1 true conv (corr=False), 3 = dgrad of 2, 4 = corr stride 2 (strides_down=2), 5 = corr stride 1.
 * kernel (kh,kw,Cin,Cout) in the reference layout; x [n,h,w,cin_of_op]; out [n,ho,wo,cout_of_op];
 * bias may be NULL; relu != 0 applies ReLU.  operand selects fp32 SIMT or bf16 tcgen05. */
int sga_op_conv(int32_t kind, int32_t operand, const float* x, const float* kernel, const float* bias,
                int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t cin, int32_t cout, int32_t relu,
                float* out, void* stream);
/* IGDN forward (x*sqrt(beta + x^2 @ gamma)) and its backward; inverse = 0 gives forward GDN (x*rsqrt). */
int sga_op_gdn(int32_t operand, int32_t inverse, const float* x, const float* beta, const float* gamma,
               int64_t pixels, int32_t channels, float* out, float* s_out, void* stream);
int sga_op_igdn_bwd(int32_t operand, const float* g, const float* x, const float* s, const float* gamma,
                    int64_t pixels, int32_t channels, float* gx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGA_B200_H_ */
