// EntropyBottleneck (factorized prior, filters=(3,3,3,3)) forward/backward + rate kernels.
// Reference semantics: compressai==1.2.4 EntropyBottleneck.{forward,_likelihood,_logits_cumulative}
// as called by src/models/subnet/entropy_model/entropy_bottleneck.py:13-28 (iwa-shi/DC_VIC);
// likelihood_to_bit: src/models/comp_model/hyperprior_vic_model.py:80-82.
// Works on NCHW directly (no permute copies): one CTA = one channel x a run of (b, hw) elements,
// the channel's 58 transformed parameters (softplus(matrix), bias, tanh(factor)) live in smem.
#include "common.cuh"

namespace dcvic {

constexpr int kEbParams = 58;
constexpr int EXPP = 59, EXPM = 62;   // e^{+m0_j}, e^{-m0_j} (first-layer matrix): slots 59-61, 62-64 of the shared table
constexpr int M0 = 0, M1 = 3, M2 = 12, M3 = 21, M4 = 30, B0 = 33, B1 = 36, B2 = 39, B3 = 42, B4 = 45, F0 = 46,
              F1 = 49, F2 = 52, F3 = 55, MED = 58;
constexpr int kEbThreads = 256;
constexpr int kEbPerThread = 4;

struct EbParamPtrs {
  const float* p[15];  // matrix0..4, bias0..4, factor0..3, quantiles
};
struct EbGradPtrs {
  float* p[14];
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// Per-element transcendentals of the hot kernels on the SFU: ex2.approx (2^-22 relative) and rcp.approx (1 ulp), i.e.
// ~2e-7 ABSOLUTE error on tanh and ~3e-7 relative on the sigmoid - two orders below the 1e-4 budget on likelihoods
// (tanh.approx itself, 5e-4 relative, is not: the likelihood is a difference of two nearby logits).  The library
// tanhf / expf (fast-math off) made these kernels instruction-bound: 24 tanh + 2 sigmoid per latent.
__device__ __forceinline__ float eb_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float eb_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// tanh(a) = 1 - 2 / (e^{2a} + 1), given E = e^{2a} (inf -> 1, 0 -> -1)
__device__ __forceinline__ float eb_tanh_from_exp(float E) { return fmaf(-2.f, eb_rcp(E + 1.f), 1.f); }
__device__ __forceinline__ float eb_tanh(float a) { return eb_tanh_from_exp(eb_ex2(a * 2.8853900817779268f)); }
__device__ __forceinline__ float eb_sigmoid(float x) { return eb_rcp(1.f + eb_ex2(x * -1.4426950408889634f)); }

// raw parameter (channel c, slot s) -> value; s indexes the packed 58-slot layout
__device__ __forceinline__ float eb_raw_param(const EbParamPtrs& P, int c, int s) {
  if (s < M1) return P.p[0][c * 3 + (s - M0)];
  if (s < M2) return P.p[1][c * 9 + (s - M1)];
  if (s < M3) return P.p[2][c * 9 + (s - M2)];
  if (s < M4) return P.p[3][c * 9 + (s - M3)];
  if (s < B0) return P.p[4][c * 3 + (s - M4)];
  if (s < B1) return P.p[5][c * 3 + (s - B0)];
  if (s < B2) return P.p[6][c * 3 + (s - B1)];
  if (s < B3) return P.p[7][c * 3 + (s - B2)];
  if (s < B4) return P.p[8][c * 3 + (s - B3)];
  if (s < F0) return P.p[9][c];
  if (s < F1) return P.p[10][c * 3 + (s - F0)];
  if (s < F2) return P.p[11][c * 3 + (s - F1)];
  if (s < F3) return P.p[12][c * 3 + (s - F2)];
  return P.p[13][c * 3 + (s - F3)];
}

__device__ __forceinline__ void eb_load_params(const EbParamPtrs& P, int c, float* sp) {
  for (int s = threadIdx.x; s <= MED; s += blockDim.x) {
    if (s == MED) {
      sp[MED] = P.p[14][c * 3 + 1];
    } else {
      const float raw = eb_raw_param(P, c, s);
      const float v = (s < B0) ? softplus_f(raw) : ((s >= F0) ? tanhf(raw) : raw);
      sp[s] = v;
      if (s < M1) { sp[EXPP + s] = expf(v); sp[EXPM + s] = expf(-v); }
    }
  }
}

// logits_cumulative at x - 1/2 and x + 1/2 at once.  First layer: a(x +- 1/2) = (m x + b) +- m / 2, so
// e^{2 a(x +- 1/2)} = e^{2 (m x + b)} e^{+-m}: one exponential serves both evaluations.
__device__ __forceinline__ void eb_logits_pair(const float* sp, float x, float& lower, float& upper) {
  float hl[3], hu[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float a = fmaf(sp[M0 + j], x, sp[B0 + j]);
    const float E = eb_ex2(a * 2.8853900817779268f);
    const float half = 0.5f * sp[M0 + j];
    const float al = a - half, au = a + half;
    hl[j] = fmaf(sp[F0 + j], eb_tanh_from_exp(E * sp[EXPM + j]), al);
    hu[j] = fmaf(sp[F0 + j], eb_tanh_from_exp(E * sp[EXPP + j]), au);
  }
#pragma unroll
  for (int layer = 0; layer < 3; ++layer) {
    const int m = M1 + 9 * layer, b = B1 + 3 * layer, f = F1 + 3 * layer;
    float gl[3], gu[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float m0 = sp[m + 3 * j], m1 = sp[m + 3 * j + 1], m2 = sp[m + 3 * j + 2], bb = sp[b + j], ff = sp[f + j];
      const float al = fmaf(m2, hl[2], fmaf(m1, hl[1], m0 * hl[0])) + bb;
      const float au = fmaf(m2, hu[2], fmaf(m1, hu[1], m0 * hu[0])) + bb;
      gl[j] = fmaf(ff, eb_tanh(al), al);
      gu[j] = fmaf(ff, eb_tanh(au), au);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) { hl[j] = gl[j]; hu[j] = gu[j]; }
  }
  lower = fmaf(sp[M4 + 2], hl[2], fmaf(sp[M4 + 1], hl[1], sp[M4] * hl[0])) + sp[B4];
  upper = fmaf(sp[M4 + 2], hu[2], fmaf(sp[M4 + 1], hu[1], sp[M4] * hu[0])) + sp[B4];
}

__device__ __forceinline__ float eb_lik_fast(float lower, float upper) {
  const float sum = lower + upper;
  const float sgn = (sum > 0.f) ? -1.f : ((sum < 0.f) ? 1.f : 0.f);  // -sign(lower+upper)
  return fabsf(eb_sigmoid(sgn * upper) - eb_sigmoid(sgn * lower));
}

// logits_cumulative for one scalar input
__device__ __forceinline__ float eb_logits(const float* sp, float x) {
  float h[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float a = fmaf(sp[M0 + j], x, sp[B0 + j]);
    h[j] = fmaf(sp[F0 + j], tanhf(a), a);
  }
#pragma unroll
  for (int layer = 0; layer < 3; ++layer) {
    const int m = M1 + 9 * layer, b = B1 + 3 * layer, f = F1 + 3 * layer;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = sp[m + 3 * j] * h[0];
      a = fmaf(sp[m + 3 * j + 1], h[1], a);
      a = fmaf(sp[m + 3 * j + 2], h[2], a);
      a += sp[b + j];
      g[j] = fmaf(sp[f + j], tanhf(a), a);
    }
    h[0] = g[0]; h[1] = g[1]; h[2] = g[2];
  }
  float o = sp[M4] * h[0];
  o = fmaf(sp[M4 + 1], h[1], o);
  o = fmaf(sp[M4 + 2], h[2], o);
  return o + sp[B4];
}

__device__ __forceinline__ float eb_lik_from_logits(float lower, float upper) {
  const float sum = lower + upper;
  const float sgn = (sum > 0.f) ? -1.f : ((sum < 0.f) ? 1.f : 0.f);  // -sign(lower+upper)
  return fabsf(sigmoid_f(sgn * upper) - sigmoid_f(sgn * lower));
}

template <int VEC>
__global__ void __launch_bounds__(kEbThreads) eb_forward_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ noise, EbParamPtrs P,
                                                                 int B, int C, int HW, float lik_bound, int x_hat_mode,
                                                                 float* __restrict__ x_hat, float* __restrict__ lik) {
  __shared__ float sp[72];
  const int c = blockIdx.y;
  eb_load_params(P, c, sp);
  __syncthreads();
  const float med = sp[MED];
  const long long per_ch = (long long)B * HW;
  // VEC == 4 (H*W % 4 == 0, 16-byte aligned tensors): one 128-bit load / store per thread and tensor, four latents
  // of one (image, channel) row; otherwise one latent per thread and step
  const long long e0 = ((long long)blockIdx.x * kEbThreads + threadIdx.x) * kEbPerThread;
  if (VEC == 4) {
    if (e0 >= per_ch) return;
    const long long b = e0 / HW, p = e0 % HW;
    const size_t o = ((size_t)b * C + c) * HW + p;
    const float4 v4 = ldg_stream(reinterpret_cast<const float4*>(x + o));
    float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (noise) n4 = ldg_stream(reinterpret_cast<const float4*>(noise + o));
    const float v[4] = {v4.x, v4.y, v4.z, v4.w}, nz[4] = {n4.x, n4.y, n4.z, n4.w};
    float xh[4], lk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float deq = __fadd_rn(rintf(__fsub_rn(v[i], med)), med);
      const float outputs = noise ? __fadd_rn(v[i], nz[i]) : deq;
      xh[i] = (x_hat_mode == 1) ? deq : outputs;
      float lower, upper;
      eb_logits_pair(sp, outputs, lower, upper);
      lk[i] = fmaxf(eb_lik_fast(lower, upper), lik_bound);
    }
    if (lik) stg_stream(reinterpret_cast<float4*>(lik + o), make_float4(lk[0], lk[1], lk[2], lk[3]));
    if (x_hat) stg_stream(reinterpret_cast<float4*>(x_hat + o), make_float4(xh[0], xh[1], xh[2], xh[3]));
  } else {
    const long long eb0 = (long long)blockIdx.x * (kEbThreads * kEbPerThread);
#pragma unroll
    for (int i = 0; i < kEbPerThread; ++i) {
      const long long e = eb0 + (long long)i * kEbThreads + threadIdx.x;
      if (e >= per_ch) break;
      const long long b = e / HW, p = e % HW;
      const size_t o = ((size_t)b * C + c) * HW + p;
      const float v = x[o];
      const float deq = __fadd_rn(rintf(__fsub_rn(v, med)), med);
      const float outputs = noise ? __fadd_rn(v, noise[o]) : deq;
      if (lik) {
        float lower, upper;
        eb_logits_pair(sp, outputs, lower, upper);
        lik[o] = fmaxf(eb_lik_fast(lower, upper), lik_bound);
      }
      if (x_hat) x_hat[o] = (x_hat_mode == 1) ? deq : outputs;
    }
  }
}

// Evaluation mode (no noise): the input of the likelihood is round(x - med) + med, i.e. per channel a function of an
// INTEGER.  The CTA evaluates the channel's likelihood once for the symbols -kEbTabR .. kEbTabR - 1 (one thread per
// entry, the same arithmetic as the direct kernel, so every value is the one the direct kernel would produce) and
// the latents become a table look-up: 12 B of HBM traffic and a handful of instructions per latent instead of 49
// SFU operations (the direct kernel is bound by the SFU pipe at 0.12 of HBM).  Symbols outside the table (|x - med|
// >= kEbTabR: far in the tails) are evaluated directly.  A CTA takes `chunk` consecutive float4 of its channel
// (kEbTabPerThread per thread) so that building the table is < 1 % of its work.
constexpr int kEbTabR = 64;
#ifndef DCVIC_EB_TAB_PER_THREAD
#define DCVIC_EB_TAB_PER_THREAD 16
#endif
constexpr int kEbTabPerThread = DCVIC_EB_TAB_PER_THREAD;
__global__ void __launch_bounds__(kEbThreads) eb_forward_table_kernel(const float* __restrict__ x, EbParamPtrs P, int B,
                                                                       int C, int HW, float lik_bound,
                                                                       float* __restrict__ x_hat,
                                                                       float* __restrict__ lik) {
  __shared__ float sp[72];
  __shared__ float tab[2 * kEbTabR];
  const int c = blockIdx.y;
  eb_load_params(P, c, sp);
  __syncthreads();
  const float med = sp[MED];
  if (threadIdx.x < 2 * kEbTabR) {
    const float outputs = __fadd_rn((float)((int)threadIdx.x - kEbTabR), med);
    float lower, upper;
    eb_logits_pair(sp, outputs, lower, upper);
    tab[threadIdx.x] = fmaxf(eb_lik_fast(lower, upper), lik_bound);
  }
  __syncthreads();
  const long long per_ch4 = (long long)B * HW / 4;               // float4 per channel (H*W % 4 == 0)
  const int hw4 = HW / 4;
  const long long q0 = (long long)blockIdx.x * (kEbThreads * kEbTabPerThread);
  auto offset_of = [&](long long q) {
    const long long b = q / hw4, p4 = q % hw4;
    return ((size_t)b * C + c) * HW + (size_t)p4 * 4;
  };
  auto finish = [&](size_t o, const float4& v4) {
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float xh[4], lk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = rintf(__fsub_rn(v[k], med));
      xh[k] = __fadd_rn(d, med);
      if (fabsf(d) < (float)kEbTabR) {                           // (NaN fails the test and takes the direct path)
        lk[k] = tab[(int)d + kEbTabR];
      } else {
        float lower, upper;
        eb_logits_pair(sp, xh[k], lower, upper);
        lk[k] = fmaxf(eb_lik_fast(lower, upper), lik_bound);
      }
    }
    if (lik) stg_stream(reinterpret_cast<float4*>(lik + o), make_float4(lk[0], lk[1], lk[2], lk[3]));
    if (x_hat) stg_stream(reinterpret_cast<float4*>(x_hat + o), make_float4(xh[0], xh[1], xh[2], xh[3]));
  };
  if (q0 + (long long)kEbThreads * kEbTabPerThread <= per_ch4) {
    // whole chunk in range (all but a channel's last CTA): four loads in flight per thread before any is used
#pragma unroll 1
    for (int i = 0; i < kEbTabPerThread; i += 4) {
      size_t o[4];
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        o[u] = offset_of(q0 + (long long)(i + u) * kEbThreads + threadIdx.x);
        v[u] = ldg_stream(reinterpret_cast<const float4*>(x + o[u]));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) finish(o[u], v[u]);
    }
  } else {
    for (int i = 0; i < kEbTabPerThread; ++i) {
      const long long q = q0 + (long long)i * kEbThreads + threadIdx.x;
      if (q >= per_ch4) break;
      const size_t o = offset_of(q);
      finish(o, ldg_stream(reinterpret_cast<const float4*>(x + o)));
    }
  }
}

// forward + backward through logits for one input, accumulating parameter grads (wrt the
// TRANSFORMED params) into acc[58]; returns d out / d x * gout.
__device__ __forceinline__ float eb_logits_backward(const float* sp, float x, float gout, float* acc) {
  float hin[4][3];   // inputs of layers 1..4
  float t[4][3];     // tanh(a) of layers 0..3
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float a = fmaf(sp[M0 + j], x, sp[B0 + j]);
    t[0][j] = eb_tanh(a);
    hin[0][j] = fmaf(sp[F0 + j], t[0][j], a);
  }
#pragma unroll
  for (int layer = 0; layer < 3; ++layer) {
    const int m = M1 + 9 * layer, b = B1 + 3 * layer, f = F1 + 3 * layer;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = sp[m + 3 * j] * hin[layer][0];
      a = fmaf(sp[m + 3 * j + 1], hin[layer][1], a);
      a = fmaf(sp[m + 3 * j + 2], hin[layer][2], a);
      a += sp[b + j];
      t[layer + 1][j] = eb_tanh(a);
      hin[layer + 1][j] = fmaf(sp[f + j], t[layer + 1][j], a);
    }
  }
  // layer 4
  float gh[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    acc[M4 + i] = fmaf(gout, hin[3][i], acc[M4 + i]);
    gh[i] = sp[M4 + i] * gout;
  }
  acc[B4] += gout;
#pragma unroll
  for (int layer = 2; layer >= 0; --layer) {
    const int m = M1 + 9 * layer, b = B1 + 3 * layer, f = F1 + 3 * layer;
    float ga[3], gin[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float tj = t[layer + 1][j];
      acc[f + j] = fmaf(gh[j], tj, acc[f + j]);
      ga[j] = gh[j] * fmaf(sp[f + j], 1.f - tj * tj, 1.f);
      acc[b + j] += ga[j];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        acc[m + 3 * j + i] = fmaf(ga[j], hin[layer][i], acc[m + 3 * j + i]);
        gin[i] = fmaf(sp[m + 3 * j + i], ga[j], gin[i]);
      }
    }
    gh[0] = gin[0]; gh[1] = gin[1]; gh[2] = gin[2];
  }
  float gx = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float tj = t[0][j];
    acc[F0 + j] = fmaf(gh[j], tj, acc[F0 + j]);
    const float ga = gh[j] * fmaf(sp[F0 + j], 1.f - tj * tj, 1.f);
    acc[B0 + j] += ga;
    acc[M0 + j] = fmaf(ga, x, acc[M0 + j]);
    gx = fmaf(sp[M0 + j], ga, gx);
  }
  return gx;
}

__global__ void __launch_bounds__(kEbThreads) eb_backward_kernel(const float* __restrict__ g_lik,
                                                                  const float* __restrict__ x,
                                                                  const float* __restrict__ noise, EbParamPtrs P,
                                                                  int B, int C, int HW, float lik_bound,
                                                                  float* __restrict__ d_x,
                                                                  float* __restrict__ packed /*[C][58]*/) {
  __shared__ float sp[72];
  __shared__ float red[kEbThreads / 32][kEbParams];
  const int c = blockIdx.y;
  eb_load_params(P, c, sp);
  __syncthreads();
  float acc[kEbParams];
#pragma unroll
  for (int s = 0; s < kEbParams; ++s) acc[s] = 0.f;
  const long long per_ch = (long long)B * HW;
  for (long long e = (long long)blockIdx.x * kEbThreads + threadIdx.x; e < per_ch;
       e += (long long)gridDim.x * kEbThreads) {
    const long long b = e / HW, p = e % HW;
    const size_t o = ((size_t)b * C + c) * HW + p;
    const float outputs = __fadd_rn(x[o], noise[o]);
    const float xl = outputs - 0.5f, xu = outputs + 0.5f;
    float lower, upper;
    eb_logits_pair(sp, outputs, lower, upper);
    const float sum = lower + upper;
    const float sgn = (sum > 0.f) ? -1.f : ((sum < 0.f) ? 1.f : 0.f);
    const float A = eb_sigmoid(sgn * upper), Bv = eb_sigmoid(sgn * lower);
    const float diff = A - Bv;
    const float L = fabsf(diff);
    float go = g_lik[o];
    if (!(L >= lik_bound || go < 0.f)) go = 0.f;
    const float sd = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
    const float gU = go * sd * A * (1.f - A) * sgn;
    const float gL = -go * sd * Bv * (1.f - Bv) * sgn;
    float gx = eb_logits_backward(sp, xu, gU, acc);
    gx += eb_logits_backward(sp, xl, gL, acc);
    if (d_x) d_x[o] = gx;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < kEbParams; ++s) {
    const float v = warp_sum(acc[s]);
    if (lane == 0) red[wid][s] = v;
  }
  __syncthreads();
  if (threadIdx.x < kEbParams) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kEbThreads / 32; ++w) v += red[w][threadIdx.x];
    atomicAdd(packed + (size_t)c * kEbParams + threadIdx.x, v);
  }
}

// packed grads (wrt transformed params) -> 14 raw-parameter gradient tensors
__global__ void __launch_bounds__(64) eb_unpack_grads_kernel(const float* __restrict__ packed, EbParamPtrs P,
                                                              EbGradPtrs G, int C) {
  const int c = blockIdx.x;
  const int s = threadIdx.x;
  if (c >= C || s >= kEbParams) return;
  const float raw = eb_raw_param(P, c, s);
  float g = packed[(size_t)c * kEbParams + s];
  if (s < B0) g *= sigmoid_f(raw);                                  // d softplus
  else if (s >= F0) { const float t = tanhf(raw); g *= (1.f - t * t); }  // d tanh
  if (s < M1) G.p[0][c * 3 + (s - M0)] = g;
  else if (s < M2) G.p[1][c * 9 + (s - M1)] = g;
  else if (s < M3) G.p[2][c * 9 + (s - M2)] = g;
  else if (s < M4) G.p[3][c * 9 + (s - M3)] = g;
  else if (s < B0) G.p[4][c * 3 + (s - M4)] = g;
  else if (s < B1) G.p[5][c * 3 + (s - B0)] = g;
  else if (s < B2) G.p[6][c * 3 + (s - B1)] = g;
  else if (s < B3) G.p[7][c * 3 + (s - B2)] = g;
  else if (s < B4) G.p[8][c * 3 + (s - B3)] = g;
  else if (s < F0) G.p[9][c] = g;
  else if (s < F1) G.p[10][c * 3 + (s - F0)] = g;
  else if (s < F2) G.p[11][c * 3 + (s - F1)] = g;
  else if (s < F3) G.p[12][c * 3 + (s - F2)] = g;
  else G.p[13][c * 3 + (s - F3)] = g;
}

// ------------------------------------------------------------------ rate
constexpr int kRateThreads = 256;
constexpr int kRateChunk = kRateThreads * 4 * 8;

// log2 on the SFU (2^-22 relative error, denormals handled): the full-precision logf made this read-only reduction
// compute-bound (20 instructions per element)
__device__ __forceinline__ float rate_lg2(float x) {
  float r;
  asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__global__ void __launch_bounds__(kRateThreads) rate_partial_kernel(const float* __restrict__ lik, long long n,
                                                                     double* __restrict__ part, int vec) {
  __shared__ double scratch[32];
  const long long b = blockIdx.y;
  const float* p = lik + b * n;
  const long long start = (long long)blockIdx.x * kRateChunk;
  float acc = 0.f;
  if (vec) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long e = start + ((long long)i * kRateThreads + threadIdx.x) * 4;
      if (e < n) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(p + e));
        acc += (rate_lg2(v.x) + rate_lg2(v.y)) + (rate_lg2(v.z) + rate_lg2(v.w));
      }
    }
  } else {
    for (int i = 0; i < 32; ++i) {
      const long long e = start + (long long)i * kRateThreads + threadIdx.x;
      if (e < n) acc += rate_lg2(p[e]);
    }
  }
  const double s = block_sum((double)acc, scratch);
  if (threadIdx.x == 0) part[b * gridDim.x + blockIdx.x] = s;
}

__global__ void __launch_bounds__(128) rate_finalize_kernel(const double* __restrict__ part, int per, long long B,
                                                             float* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  double s = 0.0;
  for (int j = lane; j < per; j += 32) s += part[b * per + j];
  s = warp_sum(s);
  // -(sum ln L) / ln 2 of the reference = -(sum log2 L)
  if (lane == 0) bits[b] = (float)(-s);
}

__global__ void __launch_bounds__(256) rate_backward_kernel(const float* __restrict__ lik,
                                                             const float* __restrict__ g_bits, long long n,
                                                             float* __restrict__ d_lik) {
  const long long b = blockIdx.y;
  const float g = -g_bits[b] * 1.4426950408889634f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    d_lik[b * n + e] = g / lik[b * n + e];
}

}  // namespace dcvic

using namespace dcvic;

extern "C" size_t dcvic_rate_workspace_bytes(int64_t B, int64_t n) {
  if (B <= 0 || n <= 0) return 0;
  return align_up((size_t)B * (size_t)ceil_div_i(n, kRateChunk) * sizeof(double), 256);
}

extern "C" int dcvic_rate_bits(const float* lik, int64_t B, int64_t n, float* bits, void* workspace, size_t ws_bytes,
                               dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(lik && bits && workspace);
  DCVIC_CHECK_ARG(B > 0 && n > 0 && B <= 65535);
  if (ws_bytes < dcvic_rate_workspace_bytes(B, n)) return DCVIC_ERR_WORKSPACE;
  const int per = ceil_div_i(n, kRateChunk);
  const int vec = ((reinterpret_cast<uintptr_t>(lik) & 15) == 0) && (n % 4 == 0);
  cudaStream_t s = (cudaStream_t)stream;
  rate_partial_kernel<<<dim3(per, (unsigned)B), kRateThreads, 0, s>>>(lik, n, reinterpret_cast<double*>(workspace),
                                                                       vec);
  rate_finalize_kernel<<<ceil_div_i(B, 4), 128, 0, s>>>(reinterpret_cast<double*>(workspace), per, B, bits);
  return dcvic_launch_status();
}

extern "C" int dcvic_rate_bits_backward(const float* lik, const float* g_bits, int64_t B, int64_t n, float* d_lik,
                                        dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(lik && g_bits && d_lik);
  DCVIC_CHECK_ARG(B > 0 && n > 0 && B <= 65535);
  rate_backward_kernel<<<dim3(min(ceil_div_i(n, 256), 1024), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
      lik, g_bits, n, d_lik);
  return dcvic_launch_status();
}

extern "C" size_t dcvic_eb_workspace_bytes(int B, int C, int HW) {
  if (B <= 0 || C <= 0 || HW <= 0) return 0;
  const size_t packed = align_up((size_t)C * kEbParams * sizeof(float), 256);
  return packed + dcvic_rate_workspace_bytes(B, (int64_t)C * HW);
}

static int eb_params_ok(const float* const* params, int n) {
  if (!params) return 0;
  for (int i = 0; i < n; ++i)
    if (!params[i]) return 0;
  return 1;
}

extern "C" int dcvic_eb_forward(const float* x, const float* noise, const float* const* params, int B, int C, int HW,
                                float lik_bound, int x_hat_mode, float* x_hat, float* lik, float* bits,
                                void* workspace, size_t ws_bytes, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(x && eb_params_ok(params, 15));
  DCVIC_CHECK_ARG(B > 0 && C > 0 && HW > 0 && C <= 65535);
  DCVIC_CHECK_ARG(x_hat || lik);
  DCVIC_CHECK_ARG(x_hat_mode == 0 || x_hat_mode == 1);
  DCVIC_CHECK_ARG(!bits || lik);
  EbParamPtrs P;
  for (int i = 0; i < 15; ++i) P.p[i] = params[i];
  const long long per_ch = (long long)B * HW;
  dim3 grid(ceil_div_i(per_ch, kEbThreads * kEbPerThread), C);
  auto misaligned = [](const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
  const bool vec = HW % 4 == 0 && !misaligned(x) && !misaligned(noise) && !misaligned(x_hat) && !misaligned(lik);
  // evaluation mode with enough latents per channel to pay for a table of 128 symbols per CTA: the look-up kernel
  if (vec && !noise && lik && per_ch >= 16 * 2 * kEbTabR) {
    dim3 tgrid(ceil_div_i(per_ch / 4, kEbThreads * kEbTabPerThread), C);
    eb_forward_table_kernel<<<tgrid, kEbThreads, 0, (cudaStream_t)stream>>>(x, P, B, C, HW, lik_bound, x_hat, lik);
  } else if (vec)
    eb_forward_kernel<4><<<grid, kEbThreads, 0, (cudaStream_t)stream>>>(x, noise, P, B, C, HW, lik_bound, x_hat_mode,
                                                                         x_hat, lik);
  else
    eb_forward_kernel<1><<<grid, kEbThreads, 0, (cudaStream_t)stream>>>(x, noise, P, B, C, HW, lik_bound, x_hat_mode,
                                                                         x_hat, lik);
  if (dcvic_launch_status() != DCVIC_OK) return DCVIC_ERR_CUDA;
  if (bits) {
    const size_t packed = align_up((size_t)C * kEbParams * sizeof(float), 256);
    if (!workspace || ws_bytes < dcvic_eb_workspace_bytes(B, C, HW)) return DCVIC_ERR_WORKSPACE;
    return dcvic_rate_bits(lik, B, (int64_t)C * HW, bits, reinterpret_cast<char*>(workspace) + packed,
                           ws_bytes - packed, stream);
  }
  return DCVIC_OK;
}

extern "C" int dcvic_eb_backward(const float* g_lik, const float* x, const float* noise, const float* const* params,
                                 int B, int C, int HW, float lik_bound, float* d_x, float* const* grads,
                                 void* workspace, size_t ws_bytes, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(g_lik && x && noise && eb_params_ok(params, 15));
  DCVIC_CHECK_ARG(B > 0 && C > 0 && HW > 0 && C <= 65535);
  DCVIC_CHECK_ARG(grads && workspace);
  for (int i = 0; i < 14; ++i) DCVIC_CHECK_ARG(grads[i] != nullptr);
  const size_t packed_bytes = (size_t)C * kEbParams * sizeof(float);
  if (ws_bytes < packed_bytes) return DCVIC_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  EbParamPtrs P;
  for (int i = 0; i < 15; ++i) P.p[i] = params[i];
  EbGradPtrs G;
  for (int i = 0; i < 14; ++i) G.p[i] = grads[i];
  float* packed = reinterpret_cast<float*>(workspace);
  if (cudaMemsetAsync(packed, 0, packed_bytes, s) != cudaSuccess) return DCVIC_ERR_CUDA;
  const long long per_ch = (long long)B * HW;
  dim3 grid(max(1, min(ceil_div_i(per_ch, kEbThreads * 2), 16)), C);
  eb_backward_kernel<<<grid, kEbThreads, 0, s>>>(g_lik, x, noise, P, B, C, HW, lik_bound, d_x, packed);
  eb_unpack_grads_kernel<<<C, 64, 0, s>>>(packed, P, G, C);
  return dcvic_launch_status();
}
