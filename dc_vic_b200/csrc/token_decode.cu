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

template <bool VEC>
__global__ void __launch_bounds__(128) token_decode_kernel(const float* __restrict__ logits,
                                                            const float* __restrict__ E,
                                                            const int64_t* __restrict__ gt, int B, int K, int HW, int D,
                                                            int64_t* __restrict__ idx, float* __restrict__ latent,
                                                            int* __restrict__ match_count) {
  constexpr int TPT = VEC ? 4 : 1;   // tokens per thread
  const int b = blockIdx.y;
  const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * TPT;
  int matches = 0;
  if (p0 < HW) {
    const float* lp = logits + (size_t)b * K * HW + p0;
    float best[TPT];
    int bi[TPT];
#pragma unroll
    for (int i = 0; i < TPT; ++i) { best[i] = -INFINITY; bi[i] = 0; }
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
        }
      } else {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(lp + (size_t)(k + u) * HW);
#pragma unroll
        for (int u = 0; u < 8; ++u) argmax_step(v[u], k + u, best[0], bi[0]);
      }
    }
    for (; k < K; ++k) {
#pragma unroll
      for (int i = 0; i < TPT; ++i) argmax_step(__ldg(lp + (size_t)k * HW + i), k, best[i], bi[i]);
    }
    // first index of a run of equal maxima: argmax_step only replaces on a strictly larger value, and the very first
    // element always replaces -inf unless it is -inf itself (then index 0 is already right)
    const size_t t0 = (size_t)b * HW + p0;
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      idx[t0 + i] = (int64_t)bi[i];
      if (gt) matches += (gt[t0 + i] == (int64_t)bi[i]);
    }
    if (latent) {
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

}  // namespace dcvic

using namespace dcvic;

extern "C" int dcvic_token_decode(const float* logits, const float* codebook, const int64_t* gt_idx, int B, int K, int HW,
                                  int D, int64_t* idx, float* latent, int* match_count, dcvic_stream_t stream) {
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
    token_decode_kernel<true><<<grid, 128, 0, s>>>(logits, codebook, gt_idx, B, K, HW, D, idx, latent, match_count);
  } else {
    dim3 grid(ceil_div_i(HW, 128), B);
    token_decode_kernel<false><<<grid, 128, 0, s>>>(logits, codebook, gt_idx, B, K, HW, D, idx, latent, match_count);
  }
  return dcvic_launch_status();
}
