/*
 * dcvic_b200.h -- C ABI of the B200-native DC-VIC hot path (libdcvic_b200.so).
 *
 * This is the drop-in boundary: everything the reference's Python modules on the hot path
 * compute is reachable through these entry points with plain pointers and sizes (no torch
 * types).  The Python host side (dc_vic_b200/*.py) binds them with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer adds.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *   - the caller owns and allocates every buffer, including `workspace` (size it with the
 *     matching *_workspace_bytes(); the first call on a workspace needs it zero-filled,
 *     later calls leave it in a reusable state); one workspace per in-flight call;
 *   - all work is enqueued on `stream`; no host synchronisation, no allocation, no global
 *     state => thread-safe per stream and CUDA-graph capturable;
 *   - return value: DCVIC_OK (0) or a negative DCVIC_ERR_* code; nothing is enqueued when
 *     an argument check fails.  There is NO CPU fallback in this library.
 *   - all tensors are FP32, contiguous, NCHW unless stated; indices are int64 (torch.long).
 *
 * Each entry point cites the reference interface it replaces (paths into iwa-shi/DC_VIC).
 */
#ifndef DCVIC_B200_H_
#define DCVIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dcvic_stream_t; /* == cudaStream_t */

enum {
  DCVIC_OK = 0,
  DCVIC_ERR_BAD_ARG = -1,     /* null pointer / non-positive size / misaligned buffer */
  DCVIC_ERR_UNSUPPORTED = -2, /* shape outside what the kernels cover (e.g. e_dim > 1024) */
  DCVIC_ERR_WORKSPACE = -3,   /* workspace too small */
  DCVIC_ERR_CUDA = -4,        /* cudaGetLastError() != cudaSuccess after enqueue */
  DCVIC_ERR_DEVICE = -5       /* device is not sm_100 (tcgen05 path requested) */
};

/* flags for dcvic_vq_forward */
enum {
  DCVIC_VQ_REUSE_PREP = 1,  /* workspace already holds the prepared codebook of this `codebook` */
  DCVIC_VQ_FORCE_EXACT = 2, /* skip the tcgen05 candidate search, use the FP32 SIMT scan */
  DCVIC_VQ_FORCE_TENSOR = 4, /* fail with DCVIC_ERR_UNSUPPORTED instead of falling back to SIMT */
  /* measurement aids (bench.py / ncu): run only one stage of the two-stage paths; outputs of the
   * skipped stage are left as the previous call on the same workspace produced them */
  DCVIC_VQ_STAGE_SEARCH_ONLY = 8,  /* codebook prep + candidate search, no finish */
  DCVIC_VQ_STAGE_FINISH_ONLY = 16, /* finish (re-rank + gather + STE + loss) from the workspace's candidates */
  DCVIC_VQ_STAGE_PREP_ONLY = 64,   /* codebook prep only */
  DCVIC_VQ_TWO_KERNELS = 128,      /* wide path: separate search and finish kernels instead of the single-pass kernel */
  DCVIC_VQ_RAGGED_HW = 32   /* dcvic_vq_path only: H*W is not a multiple of 4 (or z is not 16-byte aligned), which
                               the tcgen05 search does not take; dcvic_vq_forward sets it by itself */
};

const char* dcvic_version(void);
const char* dcvic_error_string(int code);
/* Which search kernel dcvic_vq_forward would pick: 0 = narrow fused SIMT (e_dim 4/8),
 * 1 = FP32 SIMT scan, 2 = tcgen05 FP16 candidate search + FP32 re-rank. */
int dcvic_vq_path(int D, int K, int flags);

/* ------------------------------------------------------------------ VQ quantizer ----
 * Replaces VectorQuantizer2.forward (taming/modules/vqvae/quantize.py:271-312) and, with
 * onehot/perplexity non-null, VectorQuantizer.forward (:34-90).
 *   z_nchw   [B,D,H,W]   in      codebook [K,D] in (embedding.weight)
 *   zq_nchw  [B,D,H,W]   out     value of  z + (E[idx] - z)  (the straight-through forward value)
 *   idx      [B*H*W]     out     int64, token order (b,h,w)
 *   loss     [1]         out     mean((zq-z)^2) + beta*mean((zq-z)^2)   (legacy!=0)
 *                                beta*mean(..) + mean(..)               (legacy==0)
 *   onehot   [B*H*W,K]   out, nullable   V1 min_encodings
 *   perplexity [1]       out, nullable   V1 exp(-sum(p*log(p+1e-10))), p = mean(onehot,0)
 * Distances are formed in FP32 as (sum z^2 + sum e^2) - 2 z.e ; ties resolve to the lowest
 * index (torch.argmin).
 */
size_t dcvic_vq_workspace_bytes(int B, int D, int H, int W, int K);
int dcvic_vq_forward(const float* z_nchw, const float* codebook, int B, int D, int H, int W, int K,
                     float beta, int legacy, float* zq_nchw, int64_t* idx, float* loss,
                     float* onehot, float* perplexity, int flags, void* workspace, size_t ws_bytes,
                     dcvic_stream_t stream);

/* Autograd of the above (implicit in the reference via torch autograd):
 *   dz = g_zq + g_loss * 2/(N*D) * (1 [+beta swapped if !legacy]) * (z - E[idx])
 *   dE[k] = g_loss * coef * 2/(N*D) * sum_{i: idx_i = k} (E[k] - z_i)   (g_zq never reaches E)
 * g_zq nullable (treated as 0), g_loss [1] device pointer, nullable (treated as 0).
 * dE [K,D] is OVERWRITTEN (zero-filled first), nullable.  dz nullable.
 */
int dcvic_vq_backward(const float* g_zq, const float* g_loss, const float* z_nchw, const float* codebook,
                      const int64_t* idx, int B, int D, int H, int W, int K, float beta, int legacy,
                      float* dz, float* dE, dcvic_stream_t stream);

/* get_codebook_entry (quantize.py:314-329) / vq_indices_to_latent
 * (src/models/comp_model/hyperprior_vic_model.py:165-168): out = E[idx], laid out NCHW
 * [B,D,HW] (to_nchw != 0) or token-major [B*HW, D] (to_nchw == 0).  Returns
 * DCVIC_ERR_BAD_ARG via the status word only for host-checkable errors; out-of-range
 * indices are clamped and counted in *bad_count (device int32, nullable). */
int dcvic_codebook_gather(const int64_t* idx, const float* codebook, int B, int HW, int D, int K,
                          int to_nchw, float* out, int* bad_count, dcvic_stream_t stream);

/* F.one_hot(idx, K).permute(0,3,1,2).float()  (hyperprior_vic_model.py:268-271):
 * out [B,K,HW] fp32. */
int dcvic_onehot_nchw(const int64_t* idx, int B, int HW, int K, float* out, dcvic_stream_t stream);

/* Decoder-side token path (src/models/comp_model/hyperprior_dc_vic_model.py:250-260): argmax over the vq_estimator
 * logits [B,K,HW] (first maximal index on ties, NaN counts as maximal: torch.argmax), optional accuracy against the
 * encoder's indices (match_count = #{idx == gt_idx}, device int32, zeroed here) and optional codebook gather
 * latent [B,D,HW] = E[idx] ('b h w c -> b c h w').  idx int64 [B,HW]. */
int dcvic_token_decode(const float* logits, const float* codebook, const int64_t* gt_idx, int B, int K, int HW, int D,
                       int64_t* idx, float* latent, int* match_count, dcvic_stream_t stream);

/* ------------------------------------------------------- GaussianConditional -------
 * Replaces compressai==1.2.4 GaussianConditional.forward/_likelihood/quantize as called by
 * src/models/subnet/entropy_model/gaussian_conditional.py:9-24 and
 * ste_gaussian_conditional.py:9-23.
 *   y, mu, sigma: per-sample contiguous blocks of n floats; sample b starts at
 *   ptr + b*{y,mu,sigma}_bstride (lets mu/sigma alias params.chunk(2,1) without a copy).
 *   mu nullable (GaussianScaleConditional).  noise nullable: null => eval ("dequantize"),
 *   non-null => training: outputs = y + noise (noise is U(-.5,.5) drawn by the caller).
 *   y_hat_mode: 0 = CompressAI output (y+noise | round(y-mu)+mu)
 *               1 = DC-VIC STE output  (round(y-mu)+mu in both modes; value of ste_round)
 *               | DCVIC_GC_PRECISE (2): evaluate the likelihood with library erfcf and IEEE divisions in compressai's op
 *                 order instead of the streaming kernel's 1.2e-7-relative erfc; for CDF-table construction
 *                 (GaussianConditional.update), where pmf_to_quantized_cdf amplifies ulp-level differences
 *   y_hat, lik  [B,n] contiguous, each nullable.
 *   bits [B] nullable: -sum(log2(lik)) per sample (likelihood_to_bit,
 *   hyperprior_vic_model.py:80-82 summed per sample); needs workspace.
 */
enum { DCVIC_GC_PRECISE = 2 };
size_t dcvic_gc_workspace_bytes(int64_t B, int64_t n);
int dcvic_gc_forward(const float* y, const float* mu, const float* sigma, const float* noise, int64_t B,
                     int64_t n, int64_t y_bstride, int64_t mu_bstride, int64_t sigma_bstride,
                     float scale_bound, float lik_bound, int y_hat_mode, float* y_hat, float* lik,
                     float* bits, void* workspace, size_t ws_bytes, dcvic_stream_t stream);

/* Noisy + quantized likelihood in one pass over y/mu/sigma (the two calls per CHARM slice,
 * src/models/subnet/context_model/minnen20_charm_context_model.py:96-101).
 * y_hat = STE value round(y-mu)+mu; lik_noisy from y+noise; lik_q from round(y-mu)+mu. */
int dcvic_gc_forward_dual(const float* y, const float* mu, const float* sigma, const float* noise, int64_t B,
                          int64_t n, int64_t y_bstride, int64_t mu_bstride, int64_t sigma_bstride,
                          float scale_bound, float lik_bound, float* y_hat, float* lik_noisy, float* lik_q,
                          float* bits_noisy, float* bits_q, void* workspace, size_t ws_bytes,
                          dcvic_stream_t stream);

/* Backward of the training-mode likelihood (noise != null path): given g_lik [B,n] returns
 * d/dy (== d/d outputs), d/dmu, d/dsigma incl. both LowerBound gradient rules
 * (compressai/ops/bound_ops.py).  g_yhat is NOT handled here (identity / STE, done by the
 * caller).  mu, d_mu nullable.  Outputs contiguous [B,n]. */
int dcvic_gc_backward(const float* g_lik, const float* y, const float* mu, const float* sigma,
                      const float* noise, int64_t B, int64_t n, int64_t y_bstride, int64_t mu_bstride,
                      int64_t sigma_bstride, float scale_bound, float lik_bound, float* d_y, float* d_mu,
                      float* d_sigma, dcvic_stream_t stream);

/* GaussianConditional.build_indexes (compressai 1.2.4; callers
 * minnen20_charm_context_model.py:164,199): idx = #{ table[j] < max(sigma, bound), j < T-1 }.
 * table [T] ascending (device). out int32 [n]. */
int dcvic_gc_build_indexes(const float* sigma, int64_t n, const float* table, int T, float scale_bound,
                           int32_t* out, dcvic_stream_t stream);

/* Compress-side step of one CHARM slice in one pass (replaces, per slice,
 * entropy_model_y(y, [mu, sigma], is_train=False) at minnen20_charm_context_model.py:146 plus the slice's share of
 * build_indexes (:164) and of quantize(y, "symbols", means) inside compress (:165)):
 *   y_hat = round(y - mu) + mu, lik = max(likelihood(y_hat), lik_bound), symbols = int32(round(y - mu)),
 *   indexes = build_indexes(sigma).  Tensors are [B, n] with per-batch strides (channel slices of larger tensors);
 * outputs are dense [B, n]; any output may be NULL (at least one must not be). */
int dcvic_gc_codec_step(const float* y, const float* mu, const float* sigma, int64_t B, int64_t n,
                        int64_t y_bstride, int64_t mu_bstride, int64_t sigma_bstride, const float* table, int T,
                        float scale_bound, float lik_bound, float* y_hat, float* lik, int32_t* symbols,
                        int32_t* indexes, dcvic_stream_t stream);

/* --------------------------------------------------------- EntropyBottleneck --------
 * Replaces compressai==1.2.4 EntropyBottleneck.forward/_likelihood/_logits_cumulative with
 * filters=(3,3,3,3) as called by src/models/subnet/entropy_model/entropy_bottleneck.py:13-28.
 *   x [B,C,HW] NCHW (no permute copies).  params: HOST array of 15 DEVICE pointers in the
 *   order _matrix0.._matrix4, _bias0.._bias4, _factor0.._factor3, quantiles (raw, i.e.
 *   before softplus/tanh; shapes as in CompressAI: [C,3,1],[C,3,3]x3,[C,1,3] / [C,3,1]x4,
 *   [C,1,1] / [C,3,1]x4 / [C,1,3]).
 *   noise nullable (null => eval).  x_hat_mode as y_hat_mode above (median instead of mu).
 *   bits [B] nullable.
 */
size_t dcvic_eb_workspace_bytes(int B, int C, int HW);
int dcvic_eb_forward(const float* x, const float* noise, const float* const* params /*host[15]*/, int B, int C,
                     int HW, float lik_bound, int x_hat_mode, float* x_hat, float* lik, float* bits,
                     void* workspace, size_t ws_bytes, dcvic_stream_t stream);
/* Backward of the training-mode likelihood wrt x and the 14 network parameters.
 * grads: HOST array of 14 DEVICE pointers (same order, quantiles excluded), each
 * OVERWRITTEN; d_x [B,C,HW] nullable. */
int dcvic_eb_backward(const float* g_lik, const float* x, const float* noise, const float* const* params,
                      int B, int C, int HW, float lik_bound, float* d_x, float* const* grads /*host[14]*/,
                      void* workspace, size_t ws_bytes, dcvic_stream_t stream);

/* ----------------------------------------------------------------- rate ------------
 * likelihood_to_bit (hyperprior_vic_model.py:80-82) and the per-sample form
 * (src/trainer/dual_cond_rate_distortion_vq_code_trainer.py:100-108):
 * bits[b] = -sum_i log(lik[b,i]) / ln 2.   lik [B,n] contiguous, bits [B]. */
size_t dcvic_rate_workspace_bytes(int64_t B, int64_t n);
int dcvic_rate_bits(const float* lik, int64_t B, int64_t n, float* bits, void* workspace, size_t ws_bytes,
                    dcvic_stream_t stream);
/* d bits[b] / d lik = -1/(lik ln2): d_lik[b,i] = -g_bits[b] / (lik[b,i] * ln 2). */
int dcvic_rate_bits_backward(const float* lik, const float* g_bits, int64_t B, int64_t n, float* d_lik,
                             dcvic_stream_t stream);

/* ste_round (src/models/subnet/entropy_model/ste_round.py:4-5): out = (rint(x)-x)+x. */
int dcvic_ste_round(const float* x, int64_t n, float* out, dcvic_stream_t stream);

/* HOST function (model-setup time, CPU in the reference as well): compressai==1.2.4
 * _CXX.pmf_to_quantized_cdf used by EntropyBottleneck.update / GaussianConditional.update
 * (hyperprior_dc_vic_model.py:65-68).  pmf host[n] -> cdf host[n+1], cdf[n] = 1<<precision,
 * every symbol width >= 1. */
int dcvic_pmf_to_quantized_cdf(const float* pmf, int n, int precision, int32_t* cdf);

/* dcvic_token_decode with the rest of the decoder-side token path fused in:
 *  - pq_weight [D_out, D] / pq_bias [D_out] (nullable): ldm VQModel.post_quant_conv, the 1x1 conv applied to the looked-up
 *    latent (hyperprior_dc_vic_model.py:259-260); latent is then [B, D_out, HW];
 *  - loss_sums (device double[2], nullable; needs gt_idx): sum over tokens of the code cross entropy and of the focal
 *    term (1 - p_t)^gamma CE (src/losses/cross_entropy_loss.py:9-52; the losses are these sums / (B*HW) * loss_weight);
 *  - lse [B*HW] (nullable): log-sum-exp of every token's logits, input of dcvic_token_ce_backward.
 * dcvic_token_ce_backward: d_logits [B,K,HW] = g_loss[0] * scale * coef_t * (softmax - onehot(target)), coef_t = 1 for
 * gamma = 0 (plain CE), the focal factor otherwise; scale = loss_weight / (B*HW) for the mean reduction. */
int dcvic_token_decode_ex(const float* logits, const float* codebook, const int64_t* gt_idx, int B, int K, int HW, int D,
                          const float* pq_weight, const float* pq_bias, int D_out, float gamma, int64_t* idx,
                          float* latent, int* match_count, float* lse, double* loss_sums, dcvic_stream_t stream);
int dcvic_token_ce_backward(const float* logits, const int64_t* gt_idx, const float* lse, const float* g_loss, int B,
                            int K, int HW, float gamma, float scale, float* d_logits, dcvic_stream_t stream);

/* The same construction for a whole table ON THE DEVICE (EntropyBottleneck.update / GaussianConditional.update ->
 * EntropyModel._pmf_to_cdf): pmf [rows, width] FP32 as the likelihood kernel left it, tail_mass [rows], lengths [rows];
 * cdf [rows, width + 2] int32 (zero beyond lengths[r] + 2); *status (device int32, zero on entry) receives a negative
 * code if a row is invalid.  One thread per row. */
int dcvic_pmf_to_quantized_cdf_rows(const float* pmf, int rows, int width, const float* tail_mass, const int32_t* lengths,
                                    int precision, int32_t* cdf, int32_t* status, dcvic_stream_t stream);

/* Range-ANS coder: compressai.ans' rans64 streams (CompressAI 1.2.4 rans_interface.cpp; call sites
 * src/models/comp_model/hyperprior_dc_vic_model.py:308-328,378-387, minnen20_charm_context_model.py:175-202), bit-exact
 * with oracle/rans_oracle.c.  All pointers are device pointers.
 * encode: n_streams independent streams; stream s codes symbols[starts[s] .. starts[s+1]) with the CDF rows selected by
 *   indexes[] (cdf [rows, width] int32, cdf_sizes / offsets per row).  Its 32-bit words are written BACKWARDS from
 *   out_words[out_starts[s+1]] (the caller sizes the slot: 4 words per symbol + 8 always suffice); n_words[s] receives
 *   how many: the stream is out_words[out_starts[s+1] - n_words[s] .. out_starts[s+1]).
 * decode: advances ONE stream by n symbols; state[4] (int64, zero before the first call) carries the coder state and
 *   the read position between calls, as RansDecoder.set_stream / decode_stream do. */
int dcvic_rans_encode(const int32_t* symbols, const int32_t* indexes, const int64_t* starts, int n_streams,
                      const int32_t* cdf, int rows, int width, const int32_t* cdf_sizes, const int32_t* offsets,
                      uint32_t* out_words, const int64_t* out_starts, int32_t* n_words, dcvic_stream_t stream);
int dcvic_rans_decode(const uint32_t* words, int64_t n_words, int64_t* state, const int32_t* indexes, int64_t n,
                      const int32_t* cdf, int rows, int width, const int32_t* cdf_sizes, const int32_t* offsets,
                      int32_t* out_symbols, dcvic_stream_t stream);

/* Tiling driver for images beyond 1024 px (src/models/comp_model/hyperprior_vic_model.py:190-246
 * `_vq_encode_split`, :413-473 `decode_split`): windows of all tiles in one batch, keep-windows stitched on the device.
 * gather: out [(T*N), C, ph, pw] <- in [N, C, H, W] at origins[t] = {y0, x0} (device int32 [T][2]).
 * stitch: out [N, C, H, W] <- tiles [(T*N), C, ph, pw]; windows[t] = {y0, x0, top, bottom, left, right} (device int32
 *   [T][6]) in OUTPUT coordinates: rows [top, bottom) x columns [left, right) of the output come from tile t, whose
 *   own origin in the output is (y0, x0).  vec_ok != 0: every origin / window column is a multiple of 4. */
int dcvic_tile_gather(const float* in, int N, int C, int H, int W, const int32_t* origins, int T, int ph, int pw,
                      int vec_ok, float* out, dcvic_stream_t stream);
int dcvic_tile_stitch(const float* tiles, int N, int C, int ph, int pw, const int32_t* windows, int T, int vec_ok,
                      float* out, int H, int W, dcvic_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DCVIC_B200_H_ */
