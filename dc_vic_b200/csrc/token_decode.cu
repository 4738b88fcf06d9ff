// Decoder-side token path: argmax over the vq_estimator logits, accuracy against the encoder's
// indices and the codebook gather, in one pass over the logits.
// Reference: src/models/comp_model/hyperprior_dc_vic_model.py:250-260
//     out_vq_indices = torch.argmax(out_vq_logits, dim=1)              # [B,H,W], first maximal index on ties
//     vq_accuracy    = (out_vq_indices == gt_vq_indices).float().mean()
//     vq_latent      = vq_indices_to_latent(out_vq_indices)            # embedding + 'b h w c -> b c h w'
// HBM-bound: 4K bytes of logits per token in, 8 + 4D bytes out.  One thread owns 4 consecutive tokens (16-byte loads
// along the token axis of the NCHW logits, 8 channels in flight), so a warp reads 512 contiguous bytes per channel.
#include "common.cuh"
#include <float.h>

namespace dcvic {

// torch.argmax semantics: NaN counts as the maximum; the first maximal index wins.
__device__ __forceinline__ void argmax_step(float v, int k, float& best, int& bi) {
  if (v > best || (v != v && best == best)) {
    best = v;
    bi = k;
  }
}

// Extras of the extended entry point (all optional):
//   * post_quant_conv (ldm VQModel's 1x1 conv after the lookup, hyperprior_dc_vic_model.py:259): latent[o] =
//     pq_b[o] + sum_i pq_w[o][i] E[idx][i] - the conv of a gathered row is a row of a transformed codebook, so it
//     costs D_out x D_in FMAs per token here instead of a second pass over the latent;
//   * the code cross-entropy / focal loss of src/losses/cross_entropy_loss.py:9-52 on the same logits: an online
//     log-sum-exp rides along the arg-max scan; per token CE = lse - logit[target], focal = (1 - p_t)^gamma CE,
//     summed into loss_sums[0..1] (double), lse kept for the backward.
struct TokenExtras {
  const float* pq_w;     // [D_out][D] or null
  const float* pq_b;     // [D_out] or null
  int D_out;
  float gamma;           // focal exponent
  float* lse;            // [B*HW] or null
  double* loss_sums;     // [2]: sum CE, sum focal (zeroed by the launcher) or null
};

__device__ __forceinline__ void lse_step(float v, float& mx, float& se) {
  if (v > mx) { se = se * __expf(mx - v) + 1.f; mx = v; }
  else se += __expf(v - mx);
}

template <bool VEC>
__global__ void __launch_bounds__(128) token_decode_kernel(const float* __restrict__ logits,
                                                            const float* __restrict__ E,
                                                            const int64_t* __restrict__ gt, int B, int K, int HW, int D,
                                                            int64_t* __restrict__ idx, float* __restrict__ latent,
                                                            int* __restrict__ match_count, TokenExtras ex) {
  constexpr int TPT = VEC ? 4 : 1;   // tokens per thread
  const int b = blockIdx.y;
  const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * TPT;
  int matches = 0;
  if (p0 < HW) {
    const float* lp = logits + (size_t)b * K * HW + p0;
    float best[TPT], mx[TPT], se[TPT];
    int bi[TPT];
    const bool want_loss = ex.loss_sums != nullptr || ex.lse != nullptr;
#pragma unroll
    for (int i = 0; i < TPT; ++i) { best[i] = -INFINITY; bi[i] = 0; mx[i] = -INFINITY; se[i] = 0.f; }
    int k = 0;
    for (; k + 8 <= K; k += 8) {
      if (VEC) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ldg_stream(reinterpret_cast<const float4*>(lp + (size_t)(k + u) * HW));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          argmax_step(v[u].x, k + u, best[0], bi[0]);
          argmax_step(v[u].y, k + u, best[TPT > 1 ? 1 : 0], bi[TPT > 1 ? 1 : 0]);
          argmax_step(v[u].z, k + u, best[TPT > 2 ? 2 : 0], bi[TPT > 2 ? 2 : 0]);
          argmax_step(v[u].w, k + u, best[TPT > 3 ? 3 : 0], bi[TPT > 3 ? 3 : 0]);
          if (want_loss) {
            lse_step(v[u].x, mx[0], se[0]);
            lse_step(v[u].y, mx[TPT > 1 ? 1 : 0], se[TPT > 1 ? 1 : 0]);
            lse_step(v[u].z, mx[TPT > 2 ? 2 : 0], se[TPT > 2 ? 2 : 0]);
            lse_step(v[u].w, mx[TPT > 3 ? 3 : 0], se[TPT > 3 ? 3 : 0]);
          }
        }
      } else {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(lp + (size_t)(k + u) * HW);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          argmax_step(v[u], k + u, best[0], bi[0]);
          if (want_loss) lse_step(v[u], mx[0], se[0]);
        }
      }
    }
    for (; k < K; ++k) {
#pragma unroll
      for (int i = 0; i < TPT; ++i) {
        const float v = __ldg(lp + (size_t)k * HW + i);
        argmax_step(v, k, best[i], bi[i]);
        if (want_loss) lse_step(v, mx[i], se[i]);
      }
    }
    // first index of a run of equal maxima: argmax_step only replaces on a strictly larger value, and the very first
    // element always replaces -inf unless it is -inf itself (then index 0 is already right)
    const size_t t0 = (size_t)b * HW + p0;
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      idx[t0 + i] = (int64_t)bi[i];
      if (gt) matches += (gt[t0 + i] == (int64_t)bi[i]);
    }
    if (want_loss) {
      float ce_sum = 0.f, fo_sum = 0.f;
#pragma unroll
      for (int i = 0; i < TPT; ++i) {
        const float l = mx[i] + __logf(se[i]);
        if (ex.lse) ex.lse[t0 + i] = l;
        if (ex.loss_sums && gt) {
          const long long tg = gt[t0 + i];
          const float ce = l - __ldg(lp + (size_t)tg * HW + i);
          const float pt = __expf(-ce);
          ce_sum += ce;
          fo_sum += (ex.gamma == 0.f ? 1.f : powf(fmaxf(1.f - pt, 0.f), ex.gamma)) * ce;
        }
      }
      if (ex.loss_sums) {
        const double a = warp_sum((double)ce_sum), f = warp_sum((double)fo_sum);
        if ((threadIdx.x & 31) == 0) { atomicAdd(ex.loss_sums, a); atomicAdd(ex.loss_sums + 1, f); }
      }
    }
    if (latent && ex.pq_w) {
      // lookup + 1x1 post_quant_conv: D_out outputs per token from the D values of its codebook row
      for (int o = 0; o < ex.D_out; ++o) {
        float acc[TPT];
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
          float a = ex.pq_b ? __ldg(ex.pq_b + o) : 0.f;
          for (int c = 0; c < D; ++c) a = fmaf(__ldg(ex.pq_w + (size_t)o * D + c), __ldg(E + (size_t)bi[i] * D + c), a);
          acc[i] = a;
        }
        float* op = latent + ((size_t)b * ex.D_out + o) * HW + p0;
        if (VEC) stg_stream(reinterpret_cast<float4*>(op), make_float4(acc[0], acc[TPT > 1 ? 1 : 0], acc[TPT > 2 ? 2 : 0],
                                                                      acc[TPT > 3 ? 3 : 0]));
        else op[0] = acc[0];
      }
    } else if (latent) {
      for (int c = 0; c < D; ++c) {
        float* o = latent + ((size_t)b * D + c) * HW + p0;
        if (VEC) {
          stg_stream(reinterpret_cast<float4*>(o), make_float4(__ldg(E + (size_t)bi[0] * D + c),
                                                              __ldg(E + (size_t)bi[TPT > 1 ? 1 : 0] * D + c),
                                                              __ldg(E + (size_t)bi[TPT > 2 ? 2 : 0] * D + c),
                                                              __ldg(E + (size_t)bi[TPT > 3 ? 3 : 0] * D + c)));
        } else {
          o[0] = __ldg(E + (size_t)bi[0] * D + c);
        }
      }
    }
  }
  if (match_count) {
    matches = (int)warp_sum((float)matches);   // <= 128 per warp: exact in float
    if ((threadIdx.x & 31) == 0 && matches) atomicAdd(match_count, matches);
  }
}

// d loss / d logits of the mean code CE / focal loss: g * coef_t * (softmax_k - [k == target]) / n_tokens with
// coef_t = (1 - p_t)^gamma + gamma CE_t p_t (1 - p_t)^(gamma - 1)   (= 1 for the plain cross entropy).
__global__ void __launch_bounds__(256) token_ce_backward_kernel(const float* __restrict__ logits,
                                                                 const int64_t* __restrict__ gt,
                                                                 const float* __restrict__ lse,
                                                                 const float* __restrict__ g_loss, int B, int K, int HW,
                                                                 float gamma, float scale, float* __restrict__ d_logits) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const size_t t = (size_t)b * HW + p;
  const float l = lse[t];
  const long long tg = gt[t];
  const float* lp = logits + (size_t)b * K * HW + p;
  float* dp = d_logits + (size_t)b * K * HW + p;
  const float ce = l - __ldg(lp + (size_t)tg * HW);
  float coef = 1.f;
  if (gamma != 0.f) {
    const float pt = __expf(-ce), om = fmaxf(1.f - pt, 0.f);
    coef = powf(om, gamma) + gamma * ce * pt * (gamma == 1.f ? 1.f : powf(om, gamma - 1.f));
  }
  const float g = __ldg(g_loss) * scale * coef;
  for (int k = 0; k < K; ++k) {
    const float pk = __expf(__ldg(lp + (size_t)k * HW) - l);
    dp[(size_t)k * HW] = g * (pk - (k == tg ? 1.f : 0.f));
  }
}

}  // namespace dcvic

using namespace dcvic;

static int token_decode_launch(const float* logits, const float* codebook, const int64_t* gt_idx, int B, int K, int HW,
                               int D, int64_t* idx, float* latent, int* match_count, TokenExtras ex,
                               dcvic_stream_t stream);

extern "C" int dcvic_token_decode(const float* logits, const float* codebook, const int64_t* gt_idx, int B, int K, int HW,
                                  int D, int64_t* idx, float* latent, int* match_count, dcvic_stream_t stream) {
  return token_decode_launch(logits, codebook, gt_idx, B, K, HW, D, idx, latent, match_count,
                             TokenExtras{nullptr, nullptr, 0, 0.f, nullptr, nullptr}, stream);
}

extern "C" int dcvic_token_decode_ex(const float* logits, const float* codebook, const int64_t* gt_idx, int B, int K,
                                     int HW, int D, const float* pq_weight, const float* pq_bias, int D_out, float gamma,
                                     int64_t* idx, float* latent, int* match_count, float* lse, double* loss_sums,
                                     dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(!pq_weight || (D_out > 0 && latent));
  DCVIC_CHECK_ARG(!loss_sums || gt_idx);
  if (loss_sums && cudaMemsetAsync(loss_sums, 0, 2 * sizeof(double), (cudaStream_t)stream) != cudaSuccess)
    return DCVIC_ERR_CUDA;
  return token_decode_launch(logits, codebook, gt_idx, B, K, HW, D, idx, latent, match_count,
                             TokenExtras{pq_weight, pq_bias, D_out, gamma, lse, loss_sums}, stream);
}

extern "C" int dcvic_token_ce_backward(const float* logits, const int64_t* gt_idx, const float* lse, const float* g_loss,
                                       int B, int K, int HW, float gamma, float scale, float* d_logits,
                                       dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(logits && gt_idx && lse && g_loss && d_logits);
  DCVIC_CHECK_ARG(B > 0 && K > 0 && HW > 0 && B <= 65535);
  token_ce_backward_kernel<<<dim3(ceil_div_i(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(logits, gt_idx, lse, g_loss, B,
                                                                                            K, HW, gamma, scale, d_logits);
  return dcvic_launch_status();
}

static int token_decode_launch(const float* logits, const float* codebook, const int64_t* gt_idx, int B, int K, int HW,
                               int D, int64_t* idx, float* latent, int* match_count, TokenExtras ex,
                               dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(logits && idx);
  DCVIC_CHECK_ARG(B > 0 && K > 0 && HW > 0);
  DCVIC_CHECK_ARG(!latent || (codebook && D > 0));
  DCVIC_CHECK_ARG(!gt_idx || match_count);
  if (B > 65535) return DCVIC_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (match_count && cudaMemsetAsync(match_count, 0, sizeof(int), s) != cudaSuccess) return DCVIC_ERR_CUDA;
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) &&
                   (!latent || (reinterpret_cast<uintptr_t>(latent) & 15) == 0);
  if (vec) {
    dim3 grid(ceil_div_i(HW, 128 * 4), B);
    token_decode_kernel<true><<<grid, 128, 0, s>>>(logits, codebook, gt_idx, B, K, HW, D, idx, latent, match_count, ex);
  } else {
    dim3 grid(ceil_div_i(HW, 128), B);
    token_decode_kernel<false><<<grid, 128, 0, s>>>(logits, codebook, gt_idx, B, K, HW, D, idx, latent, match_count, ex);
  }
  return dcvic_launch_status();
}
