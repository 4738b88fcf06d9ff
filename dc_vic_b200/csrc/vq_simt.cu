// FP32 SIMT kernels of the VQ quantizer path (sm_100a):
//   * codebook preparation (|e|^2, max|e|, FP16 copy with the -|e|^2/2 fold for the tensor search)
//   * narrow fused forward   (e_dim 4 / 8: the codebooks DC-VIC itself uses, K=256..16384)
//   * exact FP32 search      (any e_dim <= 1024, any K): register-tiled distance scan + argmin
//   * finish                 (FP32 re-rank of candidates + gather + straight-through value + loss)
//   * V1 extras              (one-hot rows, perplexity)
//   * backward               (dz, dE scatter-add)
//   * gather / one-hot feature
// Reference semantics: taming/modules/vqvae/quantize.py:34-107, :271-329 (iwa-shi/DC_VIC).
// Distances are formed as (sum z^2 + sum e^2) - 2 z.e in FP32; ties resolve to the lowest index.
#include "vq_common.cuh"
#include <float.h>

namespace dcvic {

// ------------------------------------------------------------------ codebook prepare
// One warp per code.  ee[k] = sum_c fl(e^2) (lane-strided partials + xor tree); FP16 codebook in CHUNK-MAJOR layout
// cb16[chunk][k][64] (chunk = 64 channels: every chunk is a dense [K x 128 B] matrix, so one 3-D TMA box brings all
// chunks of a code range): columns 0..D-1 = fp16(e), columns D..D+2 (the first three of the pad chunk) = three-way
// FP16 split of -ee[k]/2 (3 x 11 significant bits; exact down to FP16's 2^-24 grid), rest of the pad zero.
// emax[0] = max_k |e_k| (rounded up), emax[1] != 0 if the codebook does not fit FP16's range: per-CTA values go to
// `scratch`, the CTA that finishes last reduces them (no atomics on floats, no memset before the kernel;
// counters[kCtrPrep] is zero on entry and reset here).
// Launched with programmatic dependent launch: it waits for its predecessor in the stream (which may be the
// producer of the codebook or of z) before it reads anything, then lets the tensor search start - which may load z
// while the codebook is being prepared.
__global__ void __launch_bounds__(256) vq_prepare_kernel(const float* __restrict__ E, int K, int D,
                                                          float* __restrict__ ee, float* __restrict__ scratch,
                                                          float* __restrict__ emax, __half* __restrict__ cb16,
                                                          float* __restrict__ eperm, unsigned* __restrict__ counters) {
  __shared__ float s_m[8];
  __shared__ float s_r[8];
  __shared__ int s_u[8];
  __shared__ int s_last;
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int k = blockIdx.x * 8 + wid;
  float wm = 0.f, wr = 0.f;
  bool wu = false;
  if (k < K) {
    const float* row = E + (size_t)k * D;
    auto at = [&](int c) { return ((size_t)(c >> 6) * K + k) * 64 + (c & 63); };   // chunk-major position
    float acc = 0.f, res = 0.f;
    bool unsafe = false;
    for (int c = lane; c < D; c += 32) {
      const float v = row[c];
      acc = __fadd_rn(acc, __fmul_rn(v, v));
      unsafe |= !(fabsf(v) < 6.0e4f);
      const __half hv = __float2half_rn(v);
      const float dv = v - __half2float(hv);          // exact: the FP16 rounding residual of this element
      res = fmaf(dv, dv, res);
      if (cb16) cb16[at(c)] = hv;
      if (eperm) eperm[(size_t)k * D + vq_eperm_dest(c)] = v;
    }
    acc = warp_sum(acc);
    res = warp_sum(res);
    wr = sqrtf(res) * 1.000001f;                      // |e_k - fp16(e_k)|, rounded up
    unsafe |= !(acc < 1.2e5f);                  // -ee/2 must be representable as well
    unsafe = __any_sync(0xffffffffu, unsafe);
    if (lane == 0) ee[k] = acc;
    wm = sqrtf(acc) * 1.0000002f;
    wu = unsafe;
    if (cb16) {
      const float v = unsafe ? 0.f : -0.5f * acc;
      const __half h = __float2half_rn(v);
      const float r1 = v - __half2float(h);
      const __half m = __float2half_rn(r1);
      const float r2 = r1 - __half2float(m);
      const __half l = __float2half_rn(r2);
      for (int c = lane; c < kCb16Pad; c += 32)
        cb16[at(D + c)] = c == 0 ? h : (c == 1 ? m : (c == 2 ? l : __float2half_rn(0.f)));
    }
  }
  if (lane == 0) { s_m[wid] = wm; s_r[wid] = wr; s_u[wid] = wu ? 1 : 0; }
  __syncthreads();
  const unsigned grid = gridDim.x;
  if (threadIdx.x == 0) {
    float m = 0.f, r = 0.f;
    int u = 0;
    for (int w = 0; w < 8; ++w) { m = fmaxf(m, s_m[w]); r = fmaxf(r, s_r[w]); u |= s_u[w]; }   // NaN-free: fmaxf drops
    scratch[blockIdx.x] = m;                                                // a NaN norm, and such a row is flagged unsafe
    scratch[grid + blockIdx.x] = u ? 1.f : 0.f;
    scratch[2 * grid + blockIdx.x] = r;
    __threadfence();
    s_last = atomicAdd(counters + kCtrPrep, 1u) == grid - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float m = 0.f, u = 0.f, r = 0.f;
  for (unsigned i = threadIdx.x; i < grid; i += 256) {
    m = fmaxf(m, __ldcg(scratch + i));
    u = fmaxf(u, __ldcg(scratch + grid + i));
    r = fmaxf(r, __ldcg(scratch + 2 * grid + i));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    u = fmaxf(u, __shfl_xor_sync(0xffffffffu, u, o));
    r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  }
  __syncthreads();
  if (lane == 0) { s_m[wid] = m; s_r[wid] = r; s_u[wid] = u != 0.f; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int uu = 0;
    m = 0.f;
    r = 0.f;
    for (int w = 0; w < 8; ++w) { m = fmaxf(m, s_m[w]); r = fmaxf(r, s_r[w]); uu |= s_u[w]; }
    emax[0] = m;
    emax[1] = __uint_as_float(uu ? 1u : 0u);
    emax[2] = r;                                   // max_k |e_k - fp16(e_k)|: the single-pass kernel's margin
    counters[kCtrPrep] = 0u;
  }
}

int vq_prepare_codebook(const float* codebook, int K, int D, float* ee, float* scratch, float* emax, __half* cb16,
                        float* eperm, unsigned* counters, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ceil_div_i(K, 8));
  cfg.blockDim = dim3(256);
  cfg.stream = s;
  cudaLaunchAttribute pdl[1];
  pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = pdl;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, vq_prepare_kernel, codebook, K, D, ee, scratch, emax, cb16, eperm, counters) != cudaSuccess)
    return DCVIC_ERR_CUDA;
  return dcvic_launch_status();
}


// ------------------------------------------------------------------ narrow fused forward
// One thread per token, token row in registers, codebook tile (+ |e|^2) in shared memory and
// read as warp-wide broadcasts.  Reads z NCHW directly, writes z_q NCHW + idx + loss.
template <int D>
__global__ void __launch_bounds__(128) vq_narrow_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                         int N, int HW, int K, int KT, float beta, int legacy,
                                                         float* __restrict__ zq, int64_t* __restrict__ idx,
                                                         float* __restrict__ loss, double* __restrict__ partials,
                                                         unsigned* __restrict__ counters) {
  extern __shared__ __align__(16) float smem_f[];
  float* sE = smem_f;            // [KT][D]
  float* sEE = smem_f + KT * D;  // [KT]
  __shared__ double scratch[32];

  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < N;
  const size_t base = valid ? ((size_t)(t / HW) * D * HW + (size_t)(t % HW)) : 0;
  float zr[D];
  float zz = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    zr[c] = valid ? z[base + (size_t)c * HW] : 0.f;
    zz = __fadd_rn(zz, __fmul_rn(zr[c], zr[c]));
  }
  float best = FLT_MAX;
  int bi = 0;
  for (int k0 = 0; k0 < K; k0 += KT) {
    const int kt = min(KT, K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kt * D; i += blockDim.x) sE[i] = E[(size_t)k0 * D + i];
    __syncthreads();
    for (int k = threadIdx.x; k < kt; k += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) a = __fadd_rn(a, __fmul_rn(sE[k * D + c], sE[k * D + c]));
      sEE[k] = a;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < kt; ++k) {
      float dot = __fmul_rn(zr[0], sE[k * D]);
#pragma unroll
      for (int c = 1; c < D; ++c) dot = fmaf(zr[c], sE[k * D + c], dot);
      const float d = fmaf(-2.f, dot, __fadd_rn(zz, sEE[k]));
      if (d < best) {
        best = d;
        bi = k0 + k;
      }
    }
  }
  float sq = 0.f;
  if (valid) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float e = E[(size_t)bi * D + c];
      const float diff = __fsub_rn(e, zr[c]);
      zq[base + (size_t)c * HW] = __fadd_rn(zr[c], diff);
      sq = fmaf(diff, diff, sq);
    }
    idx[t] = (int64_t)bi;
  }
  const double bsum = block_sum((double)sq, scratch);
  double total;
  if (publish_and_elect_last(bsum, partials, counters + kCtrLoss, gridDim.x, blockIdx.x, scratch, &total)) {
    if (threadIdx.x == 0) write_loss(total, (long long)N * D, beta, legacy, loss);
  }
}

// Large codebooks (K >= 2048: the 16384x4 / 16384x8 VQGANs): one thread per token leaves most of the machine idle
// at DC-VIC's token counts (45k tokens for a 2K image = 9 warps per SM, each scanning 16k codes serially).  Here
// S = 8 threads share a token and take every 8th code (8 consecutive codes = 128 contiguous bytes of shared memory per
// token group at e_dim 4), then reduce (distance, index) lexicographically with shuffles -- same distances, same
// lowest-index tie-break, 8x the threads in flight.
template <int D, int S>
__global__ void __launch_bounds__(256) vq_narrow_split_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                               int N, int HW, int K, int KT, float beta, int legacy,
                                                               float* __restrict__ zq, int64_t* __restrict__ idx,
                                                               float* __restrict__ loss, double* __restrict__ partials,
                                                               unsigned* __restrict__ counters) {
  extern __shared__ __align__(16) float smem_f[];
  float* sE = smem_f;            // [KT][D]
  float* sEE = smem_f + KT * D;  // [KT]
  __shared__ double scratch[32];
  constexpr int TOK = 256 / S;   // tokens per CTA
  const int sub = threadIdx.x % S;
  const int t = blockIdx.x * TOK + threadIdx.x / S;
  const bool valid = t < N;
  const size_t base = valid ? ((size_t)(t / HW) * D * HW + (size_t)(t % HW)) : 0;
  float zr[D];
  float zz = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    zr[c] = valid ? z[base + (size_t)c * HW] : 0.f;
    zz = __fadd_rn(zz, __fmul_rn(zr[c], zr[c]));
  }
  float best = FLT_MAX;
  int bi = 0x7fffffff;
  for (int k0 = 0; k0 < K; k0 += KT) {
    const int kt = min(KT, K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kt * D; i += blockDim.x) sE[i] = E[(size_t)k0 * D + i];
    __syncthreads();
    for (int k = threadIdx.x; k < kt; k += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) a = __fadd_rn(a, __fmul_rn(sE[k * D + c], sE[k * D + c]));
      sEE[k] = a;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = sub; k < kt; k += S) {
      float dot = __fmul_rn(zr[0], sE[k * D]);
#pragma unroll
      for (int c = 1; c < D; ++c) dot = fmaf(zr[c], sE[k * D + c], dot);
      const float d = fmaf(-2.f, dot, __fadd_rn(zz, sEE[k]));
      if (d < best) {            // k ascends within a thread: the first minimum is the lowest index
        best = d;
        bi = k0 + k;
      }
    }
  }
#pragma unroll
  for (int o = S / 2; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  float sq = 0.f;
  if (valid) {
    bi = min(max(bi, 0), K - 1);
    // lane `sub` writes channel `sub` (+S, ...) of the token
#pragma unroll
    for (int c = 0; c < D; ++c)
      if (c % S == sub) {
        const float e = E[(size_t)bi * D + c];
        const float diff = __fsub_rn(e, zr[c]);
        zq[base + (size_t)c * HW] = __fadd_rn(zr[c], diff);
        sq = fmaf(diff, diff, sq);
      }
    if (sub == 0) idx[t] = (int64_t)bi;
  }
  const double bsum = block_sum((double)sq, scratch);
  double total;
  if (publish_and_elect_last(bsum, partials, counters + kCtrLoss, gridDim.x, blockIdx.x, scratch, &total)) {
    if (threadIdx.x == 0) write_loss(total, (long long)N * D, beta, legacy, loss);
  }
}

int vq_narrow_forward(const float* z, const float* E, int B, int D, int HW, int K, float beta, int legacy, float* zq,
                      int64_t* idx, float* loss, double* partials, unsigned* counters, cudaStream_t s) {
  const int N = B * HW;
  const int KT = min(K, D == 4 ? 4096 : 2048);
  const size_t smem = (size_t)KT * (D + 1) * sizeof(float);
  if (K >= 2048) {               // large codebook: 8 threads per token (32 tokens per 256-thread CTA)
    const int g8 = ceil_div_i(N, 32);
    if (D == 4) {
      cudaFuncSetAttribute(vq_narrow_split_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      vq_narrow_split_kernel<4, 8><<<g8, 256, smem, s>>>(z, E, N, HW, K, KT, beta, legacy, zq, idx, loss, partials,
                                                          counters);
    } else if (D == 8) {
      cudaFuncSetAttribute(vq_narrow_split_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      vq_narrow_split_kernel<8, 8><<<g8, 256, smem, s>>>(z, E, N, HW, K, KT, beta, legacy, zq, idx, loss, partials,
                                                          counters);
    } else {
      return DCVIC_ERR_UNSUPPORTED;
    }
    return dcvic_launch_status();
  }
  const int grid = ceil_div_i(N, 128);
  if (D == 4) {
    cudaFuncSetAttribute(vq_narrow_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vq_narrow_kernel<4><<<grid, 128, smem, s>>>(z, E, N, HW, K, KT, beta, legacy, zq, idx, loss, partials, counters);
  } else if (D == 8) {
    cudaFuncSetAttribute(vq_narrow_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vq_narrow_kernel<8><<<grid, 128, smem, s>>>(z, E, N, HW, K, KT, beta, legacy, zq, idx, loss, partials, counters);
  } else {
    return DCVIC_ERR_UNSUPPORTED;
  }
  return dcvic_launch_status();
}

// ------------------------------------------------------------------ exact FP32 search
// 128 tokens x 128 codes per CTA tile, 16-wide e_dim chunks through shared memory, 8x8 register
// micro-tile per thread, dot products accumulated with sequential FMAs over e_dim, running
// (min, argmin) per token across code tiles.  This is the reference-order FP32 path: it is the
// product path for shapes the tensor search does not cover and the check for the ones it does.
constexpr int XT = 128, XC = 128, XD = 16;

__global__ void __launch_bounds__(256) vq_exact_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                        const float* __restrict__ ee, int N, int D, int HW, int K,
                                                        int* __restrict__ cand) {
  __shared__ __align__(16) float As[XD][XT];   // [c][token]
  __shared__ __align__(16) float Bs[XD][XC + 4];  // [c][code]
  __shared__ float s_zz[XT];
  __shared__ size_t s_base[XT];
  __shared__ float s_bd[XT][17];
  __shared__ int s_bk[XT][17];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int t0 = blockIdx.x * XT;

  if (tid < XT) {
    const int t = t0 + tid;
    size_t base = 0;
    float zz = 0.f;
    if (t < N) {
      base = (size_t)(t / HW) * D * HW + (size_t)(t % HW);
      for (int c = 0; c < D; ++c) {
        const float v = z[base + (size_t)c * HW];
        zz = __fadd_rn(zz, __fmul_rn(v, v));
      }
    }
    s_base[tid] = base;
    s_zz[tid] = zz;
  }
  __syncthreads();

  float best[8];
  int bk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    best[i] = FLT_MAX;
    bk[i] = 0x7fffffff;
  }

  for (int k0 = 0; k0 < K; k0 += XC) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < D; c0 += XD) {
      // z chunk: 16 x 128 floats, coalesced along tokens
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int li = tid + r * 256;
        const int c = li >> 7, tk = li & 127;
        const bool ok = (t0 + tk < N) && (c0 + c < D);
        As[c][tk] = ok ? z[s_base[tk] + (size_t)(c0 + c) * HW] : 0.f;
      }
      // codebook chunk: 128 codes x 16 floats, stored transposed
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int li = tid + r * 256;
        const int code = li >> 4, c = li & 15;
        const bool ok = (k0 + code < K) && (c0 + c < D);
        Bs[c][code] = ok ? E[(size_t)(k0 + code) * D + c0 + c] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < XD; ++c) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[c][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[c][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[c][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[c][64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
    // fold this code tile into the running minima (ascending code order within the thread)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int code = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (code < K) {
        const float e2 = ee[code];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int tk = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
          const float d = fmaf(-2.f, acc[i][j], __fadd_rn(s_zz[tk], e2));
          if (d < best[i] || (d == best[i] && code < bk[i])) {
            best[i] = d;
            bk[i] = code;
          }
        }
      }
    }
  }
  // cross-thread argmin per token (16 threads share a token)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int tk = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    s_bd[tk][tx] = best[i];
    s_bk[tk][tx] = bk[i];
  }
  __syncthreads();
  if (tid < XT && t0 + tid < N) {
    float d = s_bd[tid][0];
    int k = s_bk[tid][0];
    for (int j = 1; j < 16; ++j) {
      const float dj = s_bd[tid][j];
      const int kj = s_bk[tid][j];
      if (dj < d || (dj == d && kj < k)) {
        d = dj;
        k = kj;
      }
    }
    cand[t0 + tid] = k;
  }
}

int vq_exact_search(const float* z, const float* E, const float* ee, int B, int D, int HW, int K, int* cand,
                    cudaStream_t s) {
  const int N = B * HW;
  vq_exact_kernel<<<ceil_div_i(N, XT), 256, 0, s>>>(z, E, ee, N, D, HW, K, cand);
  return dcvic_launch_status();
}

// ------------------------------------------------------------------ finish
// 32 tokens x e_dim per CTA staged (transposed, padded) in shared memory so that both the NCHW
// side (lanes = tokens) and the codebook side (lanes = channels) are coalesced.
//   cand != nullptr : the search already decided (exact FP32 scan): idx = cand[t]
//   else            : meta/list from the tensor search (vq_common.cuh).  Per token: threshold =
//                     max(m0, m1) - margin; every flagged code of every chunk whose maximum reaches the
//                     threshold is re-ranked with the reference-order FP32 distance
//                     (sum z^2 + sum e^2) - 2 z.e, ties to the lowest index.  Tokens whose list
//                     overflowed (or that flag more than kCandMax codes) are scanned against the whole
//                     codebook by the full CTA.
// Phases: (0) stage z, |z|^2 per token  (1) one thread per token expands its lists into (token, code)
// pairs  (2) warps take pairs round-robin: one FP32 dot each  (3) one thread per token picks the
// minimum  (3b) full scans  (4) gather, straight-through value, loss partial, coalesced stores.
constexpr int kPairCap = kFinishTokens * kCandMax;

__global__ void __launch_bounds__(256) vq_finish_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                         const float* __restrict__ ee,
                                                         const float* __restrict__ emax_ptr,
                                                         const int* __restrict__ cand,
                                                         const VqMeta* __restrict__ meta,
                                                         const uint2* __restrict__ list, int N, int D, int HW, int K,
                                                         float beta, int legacy, float* __restrict__ zq,
                                                         int64_t* __restrict__ idx, float* __restrict__ loss,
                                                         double* __restrict__ partials,
                                                         unsigned* __restrict__ counters) {
  extern __shared__ __align__(16) float buf[];  // [D][33]
  __shared__ double scratch[32];
  __shared__ float s_zz[kFinishTokens];
  __shared__ int s_best[kFinishTokens];        // decided code, or -1 while undecided
  __shared__ int s_first[kFinishTokens + 1];   // pair range of each token
  __shared__ unsigned short s_pair_tok[kPairCap];
  __shared__ unsigned short s_pair_k[kPairCap];
  __shared__ float s_pair_d[kPairCap];
  __shared__ float s_wd[8];
  __shared__ int s_wk[8];
  __shared__ int s_nc[kFinishTokens];
  __shared__ unsigned short s_ck[kFinishTokens * kCandMax];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t0 = blockIdx.x * kFinishTokens;
  if (threadIdx.x < kFinishTokens) s_nc[threadIdx.x] = 0;
  const int tl = t0 + lane;
  const bool valid = tl < N;
  const size_t base = valid ? ((size_t)(tl / HW) * D * HW + (size_t)(tl % HW)) : 0;

  // (0) stage the token tile
  for (int c = wid; c < D; c += 8) buf[c * 33 + lane] = valid ? z[base + (size_t)c * HW] : 0.f;
  __syncthreads();

  if (cand) {
    if (wid == 0) s_best[lane] = valid ? min(max(cand[tl], 0), K - 1) : 0;
    __syncthreads();
  } else {
    for (int tok = wid; tok < kFinishTokens; tok += 8) {
      float zzp = 0.f;
      for (int c = lane; c < D; c += 32) {
        const float v = buf[c * 33 + tok];
        zzp = __fadd_rn(zzp, __fmul_rn(v, v));
      }
      zzp = warp_sum(zzp);
      if (lane == 0) s_zz[tok] = zzp;
    }
    __syncthreads();
    // (1) expand lists -> candidate codes: 8 threads per token, 4 list entries each (all loads in flight
    // at once); order within a token does not matter, the minimum is taken over (distance, index)
    {
      const int tok = threadIdx.x >> 3, sub = threadIdx.x & 7;
      const int t = t0 + tok;
      if (t < N) {
        const VqMeta mt = meta[t];
        const float thr = fmaxf(mt.m0, mt.m1) - vq_margin(mt.zz, *emax_ptr);
        const int q = sub >> 2, i0 = (sub & 3) * 4;
        const int n = q == 0 ? mt.n0 : mt.n1;
        if (n < 0 && (sub & 3) == 0) atomicAdd(&s_nc[tok], 2 * kCandMax);   // overflowed list -> full scan
        if (i0 < n) {
          const uint4* lp = reinterpret_cast<const uint4*>(list + ((size_t)t * 2 + q) * kListCap + i0);
          const uint4 e01 = lp[0], e23 = lp[1];
          const unsigned key[4] = {e01.x, e01.z, e23.x, e23.z};
          const unsigned msk[4] = {e01.y, e01.w, e23.y, e23.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i0 + i >= n) break;
            if (vq_key_upper(key[i]) < thr) continue;                 // chunk maximum (rounded towards +inf) below threshold
            unsigned mask = msk[i];
            const int c0 = (int)(key[i] & 0x7Fu) * kChunk;
            const int pos = atomicAdd(&s_nc[tok], __popc(mask));
            int w = pos;
            while (mask && w < kCandMax) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1;
              s_ck[tok * kCandMax + w++] = (unsigned short)(c0 + j);
            }
          }
        }
      }
    }
    __syncthreads();
    // pair ranges (warp 0: lane = token); tokens with one candidate are decided here
    if (wid == 0) {
      const int nc = valid ? s_nc[lane] : 1;
      const int ncand = nc > kCandMax ? -1 : nc;        // nc == 0 cannot happen (the maximum is always flagged)
      const int npairs = ncand > 1 ? ncand : 0;
      int off = npairs;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, off, o);
        if (lane >= o) off += v;
      }
      s_first[lane + 1] = off;
      if (lane == 0) s_first[0] = 0;
      off -= npairs;
      for (int i = 0; i < npairs; ++i) {
        s_pair_tok[off + i] = (unsigned short)lane;
        s_pair_k[off + i] = s_ck[lane * kCandMax + i];
      }
      int sb = 0;                                       // rows past the end: any valid code
      if (valid) sb = ncand == 1 ? (int)s_ck[lane * kCandMax] : (ncand <= 0 ? -2 : -1);
      s_best[lane] = sb;
      {   // statistics: one atomic per CTA and counter (same-address atomics from every token serialise in L2)
        const unsigned nr = __popc(__ballot_sync(0xffffffffu, valid && ncand > 1));
        const unsigned nf = __popc(__ballot_sync(0xffffffffu, valid && ncand < 1));
        if (lane == 0 && nr) atomicAdd(counters + kCtrRerank, nr);
        if (lane == 0 && nf) atomicAdd(counters + kCtrOverflow, nf);
      }
    }
    __syncthreads();
    // (2) one FP32 dot per (token, code) pair
    const int total = s_first[kFinishTokens];
    for (int p = wid; p < total; p += 8) {
      const int tok = s_pair_tok[p], k = s_pair_k[p];
      const float* er = E + (size_t)k * D;
      float dp = 0.f;
      for (int c = lane; c < D; c += 32) dp = fmaf(buf[c * 33 + tok], er[c], dp);
      const float dot = warp_sum(dp);
      if (lane == 0) s_pair_d[p] = fmaf(-2.f, dot, __fadd_rn(s_zz[tok], ee[k]));
    }
    __syncthreads();
    // (3) minimum per token, ties to the lowest index
    if (wid == 0 && s_best[lane] == -1) {
      float bd = FLT_MAX;
      int bk = 0x7fffffff;
      for (int p = s_first[lane]; p < s_first[lane + 1]; ++p) {
        const float d = s_pair_d[p];
        const int k = s_pair_k[p];
        if (d < bd || (d == bd && k < bk)) { bd = d; bk = k; }
      }
      s_best[lane] = bk;
    }
    __syncthreads();
    // (3b) full scans: the whole CTA on one token at a time (rare)
    for (int tok = 0; tok < kFinishTokens; ++tok) {
      if (s_best[tok] != -2) continue;             // uniform: read from shared memory by all threads
      const float zz = s_zz[tok];
      float bd = FLT_MAX;
      int bk = 0x7fffffff;
      for (int k = wid; k < K; k += 8) {
        const float* er = E + (size_t)k * D;
        float dp = 0.f;
        for (int c = lane; c < D; c += 32) dp = fmaf(buf[c * 33 + tok], er[c], dp);
        const float dot = warp_sum(dp);
        const float d = fmaf(-2.f, dot, __fadd_rn(zz, ee[k]));
        if (d < bd || (d == bd && k < bk)) { bd = d; bk = k; }
      }
      if (lane == 0) { s_wd[wid] = bd; s_wk[wid] = bk; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
          if (s_wd[w] < bd || (s_wd[w] == bd && s_wk[w] < bk)) { bd = s_wd[w]; bk = s_wk[w]; }
        s_best[tok] = bk;
      }
      __syncthreads();
    }
  }

  // (4) gather + straight-through value + loss partial
  float sq = 0.f;
  for (int tok = wid; tok < kFinishTokens; tok += 8) {
    const int t = t0 + tok;
    if (t >= N) break;
    const int best_k = min(max(s_best[tok], 0), K - 1);
    const float* er = E + (size_t)best_k * D;
    for (int c = lane; c < D; c += 32) {
      const float zv = buf[c * 33 + tok];
      const float diff = __fsub_rn(er[c], zv);
      buf[c * 33 + tok] = __fadd_rn(zv, diff);
      sq = fmaf(diff, diff, sq);
    }
    if (lane == 0) idx[t] = (int64_t)best_k;
  }
  __syncthreads();
  if (valid)
    for (int c = wid; c < D; c += 8) zq[base + (size_t)c * HW] = buf[c * 33 + lane];

  const double bsum = block_sum((double)sq, scratch);
  double total;
  if (publish_and_elect_last(bsum, partials, counters + kCtrLoss, gridDim.x, blockIdx.x, scratch, &total)) {
    if (threadIdx.x == 0) write_loss(total, (long long)N * D, beta, legacy, loss);
  }
}

// ------------------------------------------------------------------ finish, 128-bit version
// Same arithmetic as vq_finish_kernel, for e_dim % 4 == 0 and H*W % 4 == 0 (16-byte aligned tensors).  The
// 32-token tile lives in shared memory token-major ([32][D+4] floats), so that
//   * NCHW loads / stores move 4 consecutive tokens per thread (LDG.128 / STG.128, 128 contiguous bytes per
//     8 lanes) and are transposed 4x4 in registers on the way in and out,
//   * codebook rows and token rows are both read as float4 along channels (lane = 4 channels of each
//     128-channel block): one (token, code) dot costs 2 LDS.128 + 2 LDG.128 + 8 FFMA + the warp sum at e_dim 256.
// Row stride D+4 floats keeps every quarter-warp's 16-byte accesses on 8 distinct bank groups.
// Three block barriers per tile: (A) tile staged, meta + list entries (requested first, they do not depend on
// z) expanded into per-token candidate codes with the search's own |z|^2; (B) every warp takes 4 tokens from
// candidates to z_q: re-rank where more than one code was flagged (two codebook rows in flight), then the 4
// winning rows are fetched together for the gather / straight-through value / loss partial; (C) coalesced
// stores.  Tokens that need the whole codebook (overflowed list, FP16-unsafe) are scanned by the full CTA.
#ifdef DCVIC_TRACE
__device__ unsigned long long g_trace_fin[4096][12];
#define FT_MARK(i)                                         \
  do {                                                     \
    if (threadIdx.x == 0) {                                \
      const unsigned long long now = clock64();            \
      ft_acc[i] += now - ft_t;                             \
      ft_t = now;                                          \
    }                                                      \
  } while (0)
#else
#define FT_MARK(i)
#endif

// DT = e_dim when it is one of the specialised sizes (64, 128, 256), else 0 (run-time e_dim); T = tokens per CTA
// (16 or 32; T/4 warps, each owning 4 tokens in phase B)
template <int DT, int T>
__global__ void __launch_bounds__(T * 8, DT ? 1024 / (T * 8) : 1) vq_finish_v5_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                            const float* __restrict__ ee,
                                                            const float* __restrict__ emax_ptr,
                                                            const int* __restrict__ cand,
                                                            const VqMeta* __restrict__ meta,
                                                            const uint2* __restrict__ list, int N, int Drt, int HW, int K,
                                                            float beta, int legacy, float* __restrict__ zq,
                                                            int64_t* __restrict__ idx, float* __restrict__ loss,
                                                            double* __restrict__ partials,
                                                            unsigned* __restrict__ counters) {
  extern __shared__ __align__(16) float zt[];  // [T][D + 4]
  __shared__ int s_best[T];                  // decided code
  __shared__ float s_wd[T / 4];
  __shared__ int s_wk[T / 4];
  __shared__ int s_nc[T];                    // flagged codes per token (> kCandMax: scan the whole codebook)
  __shared__ unsigned short s_ck[T * kCandMax];
  __shared__ int s_nfull, s_nrerank;
  const int D = DT ? DT : Drt;
  constexpr int NV = DT ? (DT + 127) / 128 : 8;      // float4 per lane and row (run-time e_dim: up to 1024)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ld = D + 4;
  constexpr int W = T / 4;                    // warps
  constexpr int TQ = T / 4, CQ = 32 / TQ;     // token quads per tile, channel quads per warp-iteration
  const int t0 = blockIdx.x * T;
#ifdef DCVIC_TRACE
  unsigned long long ft_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, ft_t = clock64();
#endif
  if (threadIdx.x < T) s_nc[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_nfull = 0; s_nrerank = 0; }

  // (A) stage the token tile: lane = tq*CQ + cq, token quad tq (tokens 4tq..4tq+3), channel quad cq; warp w takes
  // channels [4*CQ*w, 4*CQ*(w+1)) of every 128-channel pass (a quarter warp = 8 channel quads or 2 token quads x 4
  // channel quads: 8 distinct bank groups for its 16-byte stores either way)
  const int tq = lane / CQ, cq = lane % CQ;
  const int tokq = t0 + 4 * tq;                          // first token of this thread's quad
  const bool qvalid = tokq < N;                          // N % 4 == 0: quads are valid as a whole
  const size_t qbase = qvalid ? ((size_t)(tokq / HW) * D * HW + (size_t)(tokq % HW)) : 0;
  constexpr int NPASS = DT ? (DT + 127) / 128 : 1;     // 128-channel passes requested together (all of them for the
                                                        // specialised sizes: one DRAM round trip per tile, not two)
  for (int cb = 0; cb < D; cb += 128 * NPASS) {
    float4 v[NPASS][4];
#pragma unroll
    for (int h = 0; h < NPASS; ++h) {
      const int c = cb + h * 128 + wid * (4 * CQ) + cq * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[h][k] = (qvalid && c < D) ? ldg_stream(reinterpret_cast<const float4*>(z + qbase + (size_t)(c + k) * HW))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < NPASS; ++h) {
      const int c = cb + h * 128 + wid * (4 * CQ) + cq * 4;
      if (c < D) {
        float* dst = zt + (4 * tq) * ld + c;                // 4x4 transpose: 4 tokens x 4 channels
        *reinterpret_cast<float4*>(dst) = make_float4(v[h][0].x, v[h][1].x, v[h][2].x, v[h][3].x);
        *reinterpret_cast<float4*>(dst + ld) = make_float4(v[h][0].y, v[h][1].y, v[h][2].y, v[h][3].y);
        *reinterpret_cast<float4*>(dst + 2 * ld) = make_float4(v[h][0].z, v[h][1].z, v[h][2].z, v[h][3].z);
        *reinterpret_cast<float4*>(dst + 3 * ld) = make_float4(v[h][0].w, v[h][1].w, v[h][2].w, v[h][3].w);
      }
    }
  }
  // Everything above only read z (an input of the whole call), so with programmatic dependent launch it overlaps
  // the tail of the search kernel; its outputs (meta, lists, cand) are read from here on.
  pdl_wait();
  // meta + this thread's 4 list entries (8 threads per token)
  const int ltok = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int lq = sub >> 2, li0 = (sub & 3) * 4;
  const bool lvalid = !cand && (t0 + ltok) < N;
  VqMeta mt = {};
  uint4 e01 = make_uint4(0u, 0u, 0u, 0u), e23 = e01;
  if (lvalid) {
    mt = meta[t0 + ltok];
    const uint4* lp = reinterpret_cast<const uint4*>(list + ((size_t)(t0 + ltok) * 2 + lq) * kListCap + li0);
    e01 = __ldg(lp);
    e23 = __ldg(lp + 1);
  }
  // expand list entries -> candidate codes (order within a token does not matter: the minimum is over
  // (distance, index)).  The threshold uses the |z|^2 the search itself used for its margin.
  if (lvalid) {
    const float thr = fmaxf(mt.m0, mt.m1) - vq_margin(mt.zz, *emax_ptr);
    const int n = lq == 0 ? mt.n0 : mt.n1;
    if (n < 0 && (sub & 3) == 0) atomicAdd(&s_nc[ltok], 2 * kCandMax);   // overflowed / unsafe -> full scan
    const unsigned key[4] = {e01.x, e01.z, e23.x, e23.z};
    const unsigned msk[4] = {e01.y, e01.w, e23.y, e23.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (li0 + i >= n) break;
      if (vq_key_upper(key[i]) < thr) continue;                 // chunk maximum (rounded towards +inf) below threshold
      unsigned mask = msk[i];
      const int c0 = (int)(key[i] & 0x7Fu) * kChunk;
      int w = atomicAdd(&s_nc[ltok], __popc(mask));
      while (mask && w < kCandMax) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        s_ck[ltok * kCandMax + w++] = (unsigned short)(c0 + j);
      }
    }
  }
  __syncthreads();
  FT_MARK(0);

  // FP32 dot of token row `tok` with a codebook row: lane takes channels 4*lane + 128*h, sequential FMAs inside
  // the lane, xor tree across lanes.  load_e / dot_e are split so that several rows can be in flight.
  auto load_e = [&](float4 (&b)[NV], const float* er) {
#pragma unroll
    for (int h = 0; h < NV; ++h) {
      const int c = lane * 4 + 128 * h;
      b[h] = c < D ? __ldg(reinterpret_cast<const float4*>(er + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto dot_e = [&](int tok, const float4 (&b)[NV]) {
    float dp = 0.f;
#pragma unroll
    for (int h = 0; h < NV; ++h) {
      const int c = lane * 4 + 128 * h;
      if (c < D) {
        const float4 a = *reinterpret_cast<const float4*>(zt + tok * ld + c);
        dp = fmaf(a.x, b[h].x, dp); dp = fmaf(a.y, b[h].y, dp); dp = fmaf(a.z, b[h].z, dp); dp = fmaf(a.w, b[h].w, dp);
      }
    }
    return warp_sum(dp);
  };
  auto zz_row = [&](int tok) {
    float zzp = 0.f;
#pragma unroll
    for (int h = 0; h < NV; ++h) {
      const int c = lane * 4 + 128 * h;
      if (c < D) {
        const float4 a = *reinterpret_cast<const float4*>(zt + tok * ld + c);
        zzp = __fadd_rn(zzp, __fmul_rn(a.x, a.x)); zzp = __fadd_rn(zzp, __fmul_rn(a.y, a.y));
        zzp = __fadd_rn(zzp, __fmul_rn(a.z, a.z)); zzp = __fadd_rn(zzp, __fmul_rn(a.w, a.w));
      }
    }
    return warp_sum(zzp);
  };

  // (B) warp w owns tokens w, w+8, w+16, w+24 from candidates to z_q
  int best[4];
  unsigned n_rerank = 0, n_full = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tok = wid + W * i;
    const int t = t0 + tok;
    int bk = 0;
    if (t < N) {
      if (cand) {
        bk = min(max(cand[t], 0), K - 1);
      } else {
        const int nc = s_nc[tok];
        if (nc == 1) {
          bk = s_ck[tok * kCandMax];
        } else if (nc > kCandMax || nc <= 0) {
          bk = -2;                                        // whole codebook, below
          ++n_full;
        } else {
          ++n_rerank;
          const float zz = zz_row(tok);
          float bd = FLT_MAX;
          bk = 0x7fffffff;
          for (int c = 0; c < nc; c += 2) {               // two candidate rows in flight
            const int k0 = s_ck[tok * kCandMax + c], k1 = s_ck[tok * kCandMax + min(c + 1, nc - 1)];
            float4 b0[NV], b1[NV];
            load_e(b0, E + (size_t)k0 * D);
            load_e(b1, E + (size_t)k1 * D);
            const float ee0 = __ldg(ee + k0), ee1 = __ldg(ee + k1);
            const float d0 = fmaf(-2.f, dot_e(tok, b0), __fadd_rn(zz, ee0));
            if (d0 < bd || (d0 == bd && k0 < bk)) { bd = d0; bk = k0; }
            if (c + 1 < nc) {
              const float d1 = fmaf(-2.f, dot_e(tok, b1), __fadd_rn(zz, ee1));
              if (d1 < bd || (d1 == bd && k1 < bk)) { bd = d1; bk = k1; }
            }
          }
        }
      }
    }
    best[i] = bk;
    if (lane == 0) s_best[tok] = bk;
  }
  if (lane == 0 && n_full) atomicAdd(&s_nfull, (int)n_full);
  if (lane == 0 && n_rerank) atomicAdd(&s_nrerank, (int)n_rerank);
  __syncthreads();
  FT_MARK(1);
  // full scans: the whole CTA on one token at a time (rare)
  if (s_nfull > 0) {
    for (int tok = 0; tok < T; ++tok) {
      if (s_best[tok] != -2) continue;             // uniform: read from shared memory by all threads
      const float zz = zz_row(tok);
      float bd = FLT_MAX;
      int bk = 0x7fffffff;
      for (int k = wid; k < K; k += W) {
        float4 b[NV];
        load_e(b, E + (size_t)k * D);
        const float d = fmaf(-2.f, dot_e(tok, b), __fadd_rn(zz, ee[k]));
        if (d < bd || (d == bd && k < bk)) { bd = d; bk = k; }
      }
      if (lane == 0) { s_wd[wid] = bd; s_wk[wid] = bk; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < W; ++w)
          if (s_wd[w] < bd || (s_wd[w] == bd && s_wk[w] < bk)) { bd = s_wd[w]; bk = s_wk[w]; }
        s_best[tok] = bk;
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) best[i] = s_best[wid + W * i];
  }
  if (threadIdx.x == 0) {   // statistics: one atomic per CTA and counter (same-address atomics serialise in L2)
    if (s_nrerank) atomicAdd(counters + kCtrRerank, (unsigned)s_nrerank);
    if (s_nfull) atomicAdd(counters + kCtrOverflow, (unsigned)s_nfull);
  }
  // gather + straight-through value + loss partial: the 4 winning rows of this warp are requested together
  float sq = 0.f;
  {
    constexpr int G = DT ? 4 : 1;               // rows fetched together: one L2 round trip for the warp's 4 tokens
#pragma unroll
    for (int i0 = 0; i0 < 4; i0 += G) {
      float4 eb[G][NV];
#pragma unroll
      for (int g = 0; g < G; ++g) load_e(eb[g], E + (size_t)min(max(best[i0 + g], 0), K - 1) * D);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int i = i0 + g;
        const int tok = wid + W * i;
        const int t = t0 + tok;
        if (t < N) {
#pragma unroll
          for (int h = 0; h < NV; ++h) {
            const int c = lane * 4 + 128 * h;
            if (c < D) {
              float4* zp = reinterpret_cast<float4*>(zt + tok * ld + c);
              const float4 a = *zp, e = eb[g][h];
              const float dx = __fsub_rn(e.x, a.x), dy = __fsub_rn(e.y, a.y), dz = __fsub_rn(e.z, a.z), dw = __fsub_rn(e.w, a.w);
              *zp = make_float4(__fadd_rn(a.x, dx), __fadd_rn(a.y, dy), __fadd_rn(a.z, dz), __fadd_rn(a.w, dw));
              sq = fmaf(dx, dx, sq); sq = fmaf(dy, dy, sq); sq = fmaf(dz, dz, sq); sq = fmaf(dw, dw, sq);
            }
          }
          if (lane == 0) idx[t] = (int64_t)min(max(best[i], 0), K - 1);
        }
      }
    }
  }
  // loss: one partial per warp, summed in a fixed order by vq_loss_finalize_kernel (deterministic; an in-kernel
  // last-CTA election cost 8.5 us of 54 on C2: four more block barriers and a device-scope fence per CTA)
  {
    const double wsum = warp_sum((double)sq);
    if (lane == 0) partials[(size_t)blockIdx.x * W + wid] = wsum;
  }
  __syncthreads();                             // closes phase B: every token row holds its z_q
  FT_MARK(2);
  // (C) write z_q back NCHW: the transpose of (A)
  if (qvalid)
    for (int c0 = wid * (4 * CQ); c0 < D; c0 += 128) {
      const int c = c0 + cq * 4;
      if (c < D) {
        const float* src = zt + (4 * tq) * ld + c;
        const float4 r0 = *reinterpret_cast<const float4*>(src), r1 = *reinterpret_cast<const float4*>(src + ld),
                     r2 = *reinterpret_cast<const float4*>(src + 2 * ld), r3 = *reinterpret_cast<const float4*>(src + 3 * ld);
        float* o = zq + qbase + (size_t)c * HW;
        stg_stream(reinterpret_cast<float4*>(o), make_float4(r0.x, r1.x, r2.x, r3.x));
        stg_stream(reinterpret_cast<float4*>(o + (size_t)HW), make_float4(r0.y, r1.y, r2.y, r3.y));
        stg_stream(reinterpret_cast<float4*>(o + 2 * (size_t)HW), make_float4(r0.z, r1.z, r2.z, r3.z));
        stg_stream(reinterpret_cast<float4*>(o + 3 * (size_t)HW), make_float4(r0.w, r1.w, r2.w, r3.w));
      }
    }
  FT_MARK(3);
#ifdef DCVIC_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 4096)
    for (int i = 0; i < 12; ++i) g_trace_fin[blockIdx.x][i] = ft_acc[i];
#endif

}

// Sum of the per-warp loss partials in index order -> the reference's loss scalar.  One CTA; launched with
// programmatic dependent launch right behind the finish kernel.
__global__ void __launch_bounds__(1024) vq_loss_finalize_kernel(const double* __restrict__ partials, int n,
                                                                 long long numel, float beta, int legacy,
                                                                 float* __restrict__ loss) {
  __shared__ double scratch[32];
  pdl_wait();
  // fixed assignment of partials to threads and a fixed tree: the same bits on every run.  Up to eight loads in
  // flight per thread (one dependent load per iteration made this one-CTA kernel several microseconds long).
  double acc = 0.0;
  for (int i0 = threadIdx.x; i0 < n; i0 += 8 * 1024) {
    double v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = (i0 + r * 1024) < n ? __ldcg(partials + i0 + r * 1024) : 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc += v[r];
  }
  const double tot = block_sum(acc, scratch);
  if (threadIdx.x == 0) write_loss(tot, numel, beta, legacy, loss);
}

int vq_launch_loss_finalize(const double* partials, int n, long long numel, float beta, int legacy, float* loss,
                            cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(1024);
  cfg.stream = s;
  cudaLaunchAttribute pdl[1];
  pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = pdl;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, vq_loss_finalize_kernel, partials, n, numel, beta, legacy, loss) != cudaSuccess)
    return DCVIC_ERR_CUDA;
  return dcvic_launch_status();
}

int vq_finish(const float* z, const float* E, const float* ee, const float* emax, const int* cand, const VqMeta* meta,
              const uint2* list, int B, int D, int HW, int K, float beta, int legacy, float* zq, int64_t* idx,
              float* loss, double* partials, unsigned* counters, cudaStream_t s) {
  const int N = B * HW;
  if (K > 65535 && !cand) return DCVIC_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(E) & 15) == 0 && vq_finish_tma_supported(z, zq, D, HW, K))
    return vq_finish_tma(z, E, ee, emax, cand, meta, list, B, D, HW, K, beta, legacy, zq, idx, loss, partials, counters,
                         s);
  const bool vec = (D % 4 == 0) && (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(zq) & 15) == 0) && ((reinterpret_cast<uintptr_t>(E) & 15) == 0);
  if (vec) {
    static const int tile_tokens = (getenv("DCVIC_FINISH_TILE") && atoi(getenv("DCVIC_FINISH_TILE")) == 16) ? 16 : 32;
    const size_t smem = (size_t)tile_tokens * (D + 4) * sizeof(float);
    if (smem > 200 * 1024) return DCVIC_ERR_UNSUPPORTED;
#define DCVIC_LAUNCH_FINISH(DT, T)                                                                                  \
  do {                                                                                                             \
    if (smem > 40 * 1024)                                                                                          \
      cudaFuncSetAttribute(vq_finish_v5_kernel<DT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    cudaLaunchConfig_t cfg{};                                                                                      \
    cfg.gridDim = dim3(ceil_div_i(N, T));                                                                          \
    cfg.blockDim = dim3(T * 8);                                                                                    \
    cfg.dynamicSmemBytes = smem;                                                                                   \
    cfg.stream = s;                                                                                                \
    cudaLaunchAttribute pdl[1];                                                                                    \
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                \
    pdl[0].val.programmaticStreamSerializationAllowed = 1;                                                         \
    cfg.attrs = pdl;                                                                                               \
    cfg.numAttrs = 1;                                                                                              \
    if (cudaLaunchKernelEx(&cfg, vq_finish_v5_kernel<DT, T>, z, E, ee, emax, cand, meta, list, N, D, HW, K, beta,   \
                           legacy, zq, idx, loss, partials, counters) != cudaSuccess)                              \
      return DCVIC_ERR_CUDA;                                                                                       \
    cfg.gridDim = dim3(1);                                                                                         \
    cfg.blockDim = dim3(1024);                                                                                     \
    cfg.dynamicSmemBytes = 0;                                                                                      \
    if (cudaLaunchKernelEx(&cfg, vq_loss_finalize_kernel, (const double*)partials, ceil_div_i(N, T) * (T / 4),      \
                           (long long)N * D, beta, legacy, loss) != cudaSuccess)                                   \
      return DCVIC_ERR_CUDA;                                                                                       \
  } while (0)
    if (tile_tokens == 32) {
      if (D == 256) DCVIC_LAUNCH_FINISH(256, 32);
      else if (D == 128) DCVIC_LAUNCH_FINISH(128, 32);
      else if (D == 64) DCVIC_LAUNCH_FINISH(64, 32);
      else DCVIC_LAUNCH_FINISH(0, 32);
    } else {
      if (D == 256) DCVIC_LAUNCH_FINISH(256, 16);
      else if (D == 128) DCVIC_LAUNCH_FINISH(128, 16);
      else if (D == 64) DCVIC_LAUNCH_FINISH(64, 16);
      else DCVIC_LAUNCH_FINISH(0, 16);
    }
#undef DCVIC_LAUNCH_FINISH
    return dcvic_launch_status();
  }
  const size_t smem = (size_t)D * 33 * sizeof(float);
  if (smem > 200 * 1024) return DCVIC_ERR_UNSUPPORTED;
  if (smem > 40 * 1024)
    cudaFuncSetAttribute(vq_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  vq_finish_kernel<<<ceil_div_i(N, kFinishTokens), 256, smem, s>>>(z, E, ee, emax, cand, meta, list, N, D, HW, K,
                                                                    beta, legacy, zq, idx, loss, partials, counters);
  return dcvic_launch_status();
}

#ifdef DCVIC_TRACE
}  // namespace dcvic
extern "C" int dcvic_debug_read_finish_trace(unsigned long long* host_out /* [4096][12] */) {
  return cudaMemcpyFromSymbol(host_out, dcvic::g_trace_fin, sizeof(dcvic::g_trace_fin)) == cudaSuccess ? 0 : -4;
}
namespace dcvic {
#endif

// ------------------------------------------------------------------ V1 extras
__global__ void __launch_bounds__(256) vq_onehot_rows_kernel(const int64_t* __restrict__ idx, int N, int K,
                                                              float* __restrict__ onehot) {
  // one warp per row; float4 stores when K % 4 == 0
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int hot = (int)idx[row];
  float* out = onehot + (size_t)row * K;
  if ((K & 3) == 0) {
    for (int c = lane * 4; c < K; c += 128) {
      float4 v = make_float4(c == hot, c + 1 == hot, c + 2 == hot, c + 3 == hot);
      stg_stream(reinterpret_cast<float4*>(out + c), v);
    }
  } else {
    for (int c = lane; c < K; c += 32) out[c] = (c == hot) ? 1.f : 0.f;
  }
}

__global__ void __launch_bounds__(256) vq_hist_kernel(const int64_t* __restrict__ idx, int N, int K,
                                                       unsigned* __restrict__ hist) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int k = (int)idx[i];
    if (k >= 0 && k < K) atomicAdd(hist + k, 1u);
  }
}

__global__ void __launch_bounds__(256) vq_perplexity_kernel(const unsigned* __restrict__ hist, int N, int K,
                                                             float* __restrict__ perplexity) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float p = (float)hist[k] / (float)N;  // torch.mean of a 0/1 fp32 column
    acc += (double)(p * logf(p + 1e-10f));
  }
  const double s = block_sum(acc, scratch);
  if (threadIdx.x == 0) *perplexity = expf(-(float)s);
}

int vq_v1_extras(const int64_t* idx, int N, int K, float* onehot, float* perplexity, unsigned* hist, unsigned* counters,
                 cudaStream_t s) {
  (void)counters;
  if (onehot) {
    vq_onehot_rows_kernel<<<ceil_div_i(N, 8), 256, 0, s>>>(idx, N, K, onehot);
    if (dcvic_launch_status() != DCVIC_OK) return DCVIC_ERR_CUDA;
  }
  if (perplexity) {
    if (cudaMemsetAsync(hist, 0, (size_t)K * sizeof(unsigned), s) != cudaSuccess) return DCVIC_ERR_CUDA;
    vq_hist_kernel<<<min(ceil_div_i(N, 256), 4 * kNumSMs), 256, 0, s>>>(idx, N, K, hist);
    vq_perplexity_kernel<<<1, 256, 0, s>>>(hist, N, K, perplexity);
  }
  return dcvic_launch_status();
}

// ------------------------------------------------------------------ backward
// Same 32-token transposed tile as the finish kernel.  dz is written coalesced along tokens;
// dE receives coalesced (lanes = channels) atomic adds.
__global__ void __launch_bounds__(256) vq_backward_kernel(const float* __restrict__ g_zq,
                                                           const float* __restrict__ g_loss,
                                                           const float* __restrict__ z, const float* __restrict__ E,
                                                           const int64_t* __restrict__ idx, int N, int D, int HW, int K,
                                                           float coef_z, float coef_e, float* __restrict__ dz,
                                                           float* __restrict__ dE) {
  extern __shared__ __align__(16) float buf[];  // [D][33] z, then dz
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t0 = blockIdx.x * kFinishTokens;
  const int tl = t0 + lane;
  const bool valid = tl < N;
  const size_t base = valid ? ((size_t)(tl / HW) * D * HW + (size_t)(tl % HW)) : 0;
  const float gl = g_loss ? *g_loss : 0.f;
  const float scale = 2.f / ((float)N * (float)D);
  const float az = gl * coef_z * scale, ae = gl * coef_e * scale;

  for (int c = wid; c < D; c += 8) buf[c * 33 + lane] = valid ? z[base + (size_t)c * HW] : 0.f;
  __syncthreads();
  for (int tok = wid; tok < kFinishTokens; tok += 8) {
    const int t = t0 + tok;
    if (t >= N) break;
    const int k = min(max((int)idx[t], 0), K - 1);
    const float* er = E + (size_t)k * D;
    for (int c = lane; c < D; c += 32) {
      const float diff = buf[c * 33 + tok] - er[c];  // z - e
      buf[c * 33 + tok] = az * diff;
      if (dE) atomicAdd(dE + (size_t)k * D + c, -ae * diff);
    }
  }
  __syncthreads();
  if (valid && dz)
    for (int c = wid; c < D; c += 8) {
      const size_t o = base + (size_t)c * HW;
      dz[o] = (g_zq ? g_zq[o] : 0.f) + buf[c * 33 + lane];
    }
}

// 128-bit version (H*W % 4 == 0, 16-byte aligned tensors): a thread moves 4 consecutive tokens of one channel per
// load / store (lane = 4*token quad + channel offset: the 4 scalar shared-memory accesses of a float4 fall on 32
// distinct banks across the warp), eight such loads in flight per thread.  Same arithmetic as vq_backward_kernel.
__global__ void __launch_bounds__(256) vq_backward_vec_kernel(const float* __restrict__ g_zq,
                                                               const float* __restrict__ g_loss,
                                                               const float* __restrict__ z, const float* __restrict__ E,
                                                               const int64_t* __restrict__ idx, int N, int D, int HW,
                                                               int K, float coef_z, float coef_e,
                                                               float* __restrict__ dz, float* __restrict__ dE) {
  extern __shared__ __align__(16) float buf[];  // [D][33] z, then dz
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tq = lane >> 2, cq = lane & 3;
  const int t0 = blockIdx.x * kFinishTokens;
  const int tokq = t0 + 4 * tq;
  const bool qvalid = tokq < N;                 // N % 4 == 0: quads are valid as a whole
  const size_t qbase = qvalid ? ((size_t)(tokq / HW) * D * HW + (size_t)(tokq % HW)) : 0;
  const float gl = g_loss ? *g_loss : 0.f;
  const float scale = 2.f / ((float)N * (float)D);
  const float az = gl * coef_z * scale, ae = gl * coef_e * scale;
  const int c_first = wid * 4 + cq;             // this thread's channels: c_first + 32 k

  for (int c0 = c_first; c0 < D; c0 += 256) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + 32 * k;
      v[k] = (qvalid && c < D) ? ldg_stream(reinterpret_cast<const float4*>(z + qbase + (size_t)c * HW))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + 32 * k;
      if (c < D) {
        float* b = buf + c * 33 + 4 * tq;
        b[0] = v[k].x; b[1] = v[k].y; b[2] = v[k].z; b[3] = v[k].w;
      }
    }
  }
  __syncthreads();
  for (int tok = wid; tok < kFinishTokens; tok += 8) {
    const int t = t0 + tok;
    if (t >= N) break;
    const int k = min(max((int)idx[t], 0), K - 1);
    const float* er = E + (size_t)k * D;
    for (int c = lane; c < D; c += 32) {
      const float diff = buf[c * 33 + tok] - er[c];  // z - e
      buf[c * 33 + tok] = az * diff;
      if (dE) atomicAdd(dE + (size_t)k * D + c, -ae * diff);
    }
  }
  __syncthreads();
  if (qvalid && dz)
    for (int c0 = c_first; c0 < D; c0 += 256) {
      float4 g[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = c0 + 32 * k;
        g[k] = (g_zq && c < D) ? ldg_stream(reinterpret_cast<const float4*>(g_zq + qbase + (size_t)c * HW))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = c0 + 32 * k;
        if (c < D) {
          const float* b = buf + c * 33 + 4 * tq;
          stg_stream(reinterpret_cast<float4*>(dz + qbase + (size_t)c * HW),
                     make_float4(g[k].x + b[0], g[k].y + b[1], g[k].z + b[2], g[k].w + b[3]));
        }
      }
    }
}

}  // namespace dcvic

using namespace dcvic;

extern "C" int dcvic_vq_backward(const float* g_zq, const float* g_loss, const float* z_nchw, const float* codebook,
                                 const int64_t* idx, int B, int D, int H, int W, int K, float beta, int legacy,
                                 float* dz, float* dE, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(z_nchw && codebook && idx);
  DCVIC_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0 && K > 0);
  DCVIC_CHECK_ARG(dz || dE);
  if ((long long)B * H * W > 0x7fffffffLL) return DCVIC_ERR_UNSUPPORTED;
  const size_t smem = (size_t)D * 33 * sizeof(float);
  if (smem > 200 * 1024) return DCVIC_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int HW = H * W, N = B * HW;
  if (dE && cudaMemsetAsync(dE, 0, (size_t)K * D * sizeof(float), s) != cudaSuccess) return DCVIC_ERR_CUDA;
  // legacy: loss = mean((sg(zq)-z)^2) + beta*mean((zq-sg(z))^2): z gets coefficient 1, E gets beta
  const float coef_z = legacy ? 1.f : beta, coef_e = legacy ? beta : 1.f;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = HW % 4 == 0 && al16(z_nchw) && al16(g_zq) && al16(dz);
  auto kern = vec ? vq_backward_vec_kernel : vq_backward_kernel;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<ceil_div_i(N, kFinishTokens), 256, smem, s>>>(g_zq, g_loss, z_nchw, codebook, idx, N, D, HW, K, coef_z, coef_e,
                                                       dz, dE);
  return dcvic_launch_status();
}

// ------------------------------------------------------------------ gather / one-hot feature
namespace dcvic {

__global__ void __launch_bounds__(256) vq_gather_nchw_kernel(const int64_t* __restrict__ idx,
                                                              const float* __restrict__ E, int N, int D, int HW, int K,
                                                              float* __restrict__ out, int* __restrict__ bad) {
  extern __shared__ __align__(16) float buf[];  // [D][33]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t0 = blockIdx.x * kFinishTokens;
  for (int tok = wid; tok < kFinishTokens; tok += 8) {
    const int t = t0 + tok;
    if (t >= N) break;
    const long long raw = idx[t];
    if ((raw < 0 || raw >= K) && lane == 0 && bad) atomicAdd(bad, 1);
    const int k = (int)min(max(raw, 0LL), (long long)K - 1);
    const float* er = E + (size_t)k * D;
    for (int c = lane; c < D; c += 32) buf[c * 33 + tok] = er[c];
  }
  __syncthreads();
  const int tl = t0 + lane;
  if (tl < N) {
    const size_t base = (size_t)(tl / HW) * D * HW + (size_t)(tl % HW);
    for (int c = wid; c < D; c += 8) out[base + (size_t)c * HW] = buf[c * 33 + lane];
  }
}

__global__ void __launch_bounds__(256) vq_gather_rows_kernel(const int64_t* __restrict__ idx,
                                                              const float* __restrict__ E, long long N, int D, int K,
                                                              float* __restrict__ out, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= N) return;
  const long long raw = idx[t];
  if ((raw < 0 || raw >= K) && lane == 0 && bad) atomicAdd(bad, 1);
  const int k = (int)min(max(raw, 0LL), (long long)K - 1);
  for (int c = lane; c < D; c += 32) out[(size_t)t * D + c] = E[(size_t)k * D + c];
}

// out[b][k][p] = (idx[b][p] == k): each thread owns 4 consecutive tokens of one (b,k) row.
__global__ void __launch_bounds__(256) vq_onehot_nchw_kernel(const int64_t* __restrict__ idx, int B, int HW, int K,
                                                              float* __restrict__ out) {
  const int b = blockIdx.z;
  const int k = blockIdx.y;
  const int64_t* ib = idx + (size_t)b * HW;
  float* ob = out + ((size_t)b * K + k) * HW;
  const bool vec = ((HW & 3) == 0);
  if (vec) {
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) * 4; p < HW; p += gridDim.x * blockDim.x * 4) {
      const longlong2 i01 = *reinterpret_cast<const longlong2*>(ib + p);
      const longlong2 i23 = *reinterpret_cast<const longlong2*>(ib + p + 2);
      float4 v = make_float4(i01.x == k, i01.y == k, i23.x == k, i23.y == k);
      stg_stream(reinterpret_cast<float4*>(ob + p), v);
    }
  } else {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x)
      ob[p] = (ib[p] == k) ? 1.f : 0.f;
  }
}

}  // namespace dcvic

extern "C" int dcvic_codebook_gather(const int64_t* idx, const float* codebook, int B, int HW, int D, int K,
                                     int to_nchw, float* out, int* bad_count, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(idx && codebook && out);
  DCVIC_CHECK_ARG(B > 0 && HW > 0 && D > 0 && K > 0);
  if ((long long)B * HW > 0x7fffffffLL) return DCVIC_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int N = B * HW;
  if (bad_count && cudaMemsetAsync(bad_count, 0, sizeof(int), s) != cudaSuccess) return DCVIC_ERR_CUDA;
  if (to_nchw) {
    const size_t smem = (size_t)D * 33 * sizeof(float);
    if (smem > 200 * 1024) return DCVIC_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(vq_gather_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vq_gather_nchw_kernel<<<ceil_div_i(N, kFinishTokens), 256, smem, s>>>(idx, codebook, N, D, HW, K, out, bad_count);
  } else {
    vq_gather_rows_kernel<<<ceil_div_i(N, 8), 256, 0, s>>>(idx, codebook, N, D, K, out, bad_count);
  }
  return dcvic_launch_status();
}

extern "C" int dcvic_onehot_nchw(const int64_t* idx, int B, int HW, int K, float* out, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(idx && out);
  DCVIC_CHECK_ARG(B > 0 && HW > 0 && K > 0);
  if (K > 65535 || B > 65535) return DCVIC_ERR_UNSUPPORTED;
  const bool vec = ((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (!vec && (HW & 3) == 0) return DCVIC_ERR_BAD_ARG;  // misaligned buffers with a vectorisable shape
  const int per = 256 * 4;
  dim3 grid(min(ceil_div_i(HW, per), 64), K, B);
  vq_onehot_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx, B, HW, K, out);
  return dcvic_launch_status();
}
