// Tiling driver for images beyond 1024 px: the reference cuts them into 512 x 512 windows with
// stride 256, runs the VQGAN encoder (or the decoder) on one window at a time in a Python loop and stitches the
// central part of every result into the output - on the CPU for the decoder
// (src/models/comp_model/hyperprior_vic_model.py:190-246 `_vq_encode_split`, :413-473 `decode_split`).
// Here the windows of ALL tiles are gathered into one batch by one copy kernel, the network runs once on the batch,
// and one copy kernel writes every tile's keep-window into the full-size output: pure data movement, HBM-bound,
// 128-bit when the geometry allows (it does for the reference's sizes: images are padded to multiples of 64).
#include "common.cuh"

namespace dcvic {

// out[(t * N + n), c, y, x] = in[n, c, y0[t] + y, x0[t] + x]
template <int VEC>
__global__ void __launch_bounds__(256) tile_gather_kernel(const float* __restrict__ in, int N, int C, int H, int W,
                                                           const int32_t* __restrict__ origin /*[T][2] y0, x0*/, int T,
                                                           int ph, int pw, float* __restrict__ out) {
  const long long rows = (long long)T * N * C * ph;
  const int per_row = pw / VEC;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int y = (int)(r % ph);
    const long long q = r / ph;
    const int c = (int)(q % C), n = (int)((q / C) % N), t = (int)(q / ((long long)C * N));
    const int y0 = origin[2 * t], x0 = origin[2 * t + 1];
    const float* src = in + (((size_t)n * C + c) * H + (y0 + y)) * W + x0;
    float* dst = out + (size_t)r * pw;
    for (int i = threadIdx.x; i < per_row; i += blockDim.x) {
      if (VEC == 4) stg_stream(reinterpret_cast<float4*>(dst) + i, ldg_stream(reinterpret_cast<const float4*>(src) + i));
      else dst[i] = src[i];
    }
  }
}

// out[n, c, t:b, l:r] = tiles[(k * N + n), c, t - y0s : b - y0s, l - x0s : r - x0s] for every tile k;
// win[k] = {y0s, x0s, t, b, l, r} in output coordinates (the keep-windows partition the output)
template <int VEC>
__global__ void __launch_bounds__(256) tile_stitch_kernel(const float* __restrict__ tiles, int N, int C, int ph, int pw,
                                                           const int32_t* __restrict__ win, int T,
                                                           float* __restrict__ out, int H, int W) {
  const long long rows = (long long)T * N * C * ph;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int y = (int)(r % ph);
    const long long q = r / ph;
    const int c = (int)(q % C), n = (int)((q / C) % N), k = (int)(q / ((long long)C * N));
    const int32_t* w = win + 6 * k;
    const int oy = w[0] + y;                       // output row of this tile row
    if (oy < w[2] || oy >= w[3]) continue;
    const int l = w[4], rr = w[5];
    const float* src = tiles + (size_t)r * pw + (l - w[1]);
    float* dst = out + (((size_t)n * C + c) * H + oy) * W + l;
    const int cnt = (rr - l) / VEC;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      if (VEC == 4) stg_stream(reinterpret_cast<float4*>(dst) + i, ldg_stream(reinterpret_cast<const float4*>(src) + i));
      else dst[i] = src[i];
    }
  }
}

}  // namespace dcvic

using namespace dcvic;

extern "C" int dcvic_tile_gather(const float* in, int N, int C, int H, int W, const int32_t* origins, int T, int ph,
                                 int pw, int vec_ok, float* out, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(in && origins && out);
  DCVIC_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && T > 0 && ph > 0 && pw > 0 && ph <= H && pw <= W);
  const long long rows = (long long)T * N * C * ph;
  const int grid = (int)(rows < 16 * kNumSMs ? rows : 16 * kNumSMs);
  const bool vec = vec_ok && pw % 4 == 0 && W % 4 == 0 && !((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15);
  if (vec) tile_gather_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(in, N, C, H, W, origins, T, ph, pw, out);
  else tile_gather_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(in, N, C, H, W, origins, T, ph, pw, out);
  return dcvic_launch_status();
}

extern "C" int dcvic_tile_stitch(const float* tiles, int N, int C, int ph, int pw, const int32_t* windows, int T,
                                 int vec_ok, float* out, int H, int W, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(tiles && windows && out);
  DCVIC_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && T > 0 && ph > 0 && pw > 0);
  const long long rows = (long long)T * N * C * ph;
  const int grid = (int)(rows < 16 * kNumSMs ? rows : 16 * kNumSMs);
  const bool vec = vec_ok && pw % 4 == 0 && W % 4 == 0 && !((reinterpret_cast<uintptr_t>(tiles) | reinterpret_cast<uintptr_t>(out)) & 15);
  if (vec) tile_stitch_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(tiles, N, C, ph, pw, windows, T, out, H, W);
  else tile_stitch_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(tiles, N, C, ph, pw, windows, T, out, H, W);
  return dcvic_launch_status();
}
