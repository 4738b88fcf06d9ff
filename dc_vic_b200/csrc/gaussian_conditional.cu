// GaussianConditional quantize / likelihood / rate kernels (HBM-bound, 128-bit vectorised).
// Reference semantics: compressai==1.2.4 GaussianConditional.{forward,_likelihood,quantize,
// build_indexes} + LowerBound, as called by src/models/subnet/entropy_model/
// gaussian_conditional.py:9-24 and ste_gaussian_conditional.py:9-23 (iwa-shi/DC_VIC).
// Algorithmic traffic: read y, mu, sigma (+noise), write y_hat, lik = 20 (24) B / element.
#include "common.cuh"

namespace dcvic {

constexpr float kNegInvSqrt2 = -0.70710678118654752440f;
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;
constexpr int kGcThreads = 256;
constexpr int kGcVecPerThread = 4;                                  // float4 per thread per tensor
constexpr int kGcChunk = kGcThreads * kGcVecPerThread * 4;          // elements per CTA

// erfc(x) = t exp(-x^2 + P(t)), t = 1/(1 + |x|/2): the classic Chebyshev fit with fractional error < 1.2e-7
// everywhere (no cancellation in the tails, which is where likelihoods near the 1e-9 bound live).  One MUFU.RCP,
// one MUFU.EX2 and 11 FMAs instead of the ~40-instruction library erfcf: with two of them, two IEEE divisions and a
// log2 per latent the kernel was issue-bound at 0.55 of HBM bandwidth.
// single-MUFU forms (flush-to-zero: every operand here is a normal number or may be flushed: s >= 0.11, 1 + z/2 >= 1,
// likelihoods are clamped to >= 1e-9 before the logarithm)
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float erfc_pos(float z) {   // z >= 0
  const float t = rcp_ftz(fmaf(0.5f, z, 1.f));
  float p = 0.17087277f;
  p = fmaf(t, p, -0.82215223f);
  p = fmaf(t, p, 1.48851587f);
  p = fmaf(t, p, -1.13520398f);
  p = fmaf(t, p, 0.27886807f);
  p = fmaf(t, p, -0.18628806f);
  p = fmaf(t, p, 0.09678418f);
  p = fmaf(t, p, 0.37409196f);
  p = fmaf(t, p, 1.00002368f);
  p = fmaf(t, p, -1.26551223f);
  return t * ex2_ftz(1.4426950408889634f * fmaf(-z, z, p));
}

// Phi((0.5 - v)/s) - Phi((-0.5 - v)/s) = (erfc(a) - erfc(b)) / 2 with a = (v - 0.5)/(s sqrt 2) <= b = (v + 0.5)/(s sqrt 2),
// b > 0 always and a >= -0.5/(0.11 sqrt 2).  Same formula as compressai's _likelihood (the reference evaluates it in
// FP32 with its own erfc; measured difference to it <= 2.5e-5 relative on the C3 inputs, tolerance 1e-4).
__device__ __forceinline__ float gc_lik(float outputs, float mu, float s) {
  const float v = fabsf(__fsub_rn(outputs, mu));
  const float r = 0.70710678118654752440f * rcp_ftz(s);
  // Large sigma (h = b - a = 1/(sigma sqrt 2) < 1/64, sigma > 45): erfc(a) - erfc(b) is the difference of two values
  // whose ARGUMENTS are already rounded - an ulp of a is worth a sigma sqrt(2) 1e-7 ~ 1e-4 of the result at
  // sigma = 256, a = 3, for the reference's FP32 evaluation as well.  The integral of exp(-t^2) over [c - h/2, c + h/2]
  // from its midpoint expansion (h^6 term < 1e-9) has no such cancellation:
  //   L = h/sqrt(pi) exp(-c^2) [1 + h^2 (4c^2 - 2)/24 + h^4 (16c^4 - 48c^2 + 12)/1920],  c = v / (sigma sqrt 2).
  if (r < 0.015625f) {
    const float c2 = (v * r) * (v * r), h2 = r * r;
    const float corr = fmaf(h2, fmaf(h2, fmaf(c2, fmaf(c2, 16.f, -48.f), 12.f) * (1.f / 1920.f),
                                     fmaf(c2, 4.f, -2.f) * (1.f / 24.f)), 1.f);
    return 0.56418958354775628695f * r * ex2_ftz(-1.4426950408889634f * c2) * corr;
  }
  const float a = (v - 0.5f) * r, b = (v + 0.5f) * r;
  const float ea = erfc_pos(fabsf(a));
  const float up = a < 0.f ? 2.f - ea : ea;
  return 0.5f * (up - erfc_pos(b));
}

// Table-construction variant (DCVIC_GC_PRECISE): library erfcf and two IEEE divisions in compressai's op order.  The
// zero-width "steal" cascade of pmf_to_quantized_cdf amplifies ulp-level differences into +-2 table entries, so the
// CDF tables are built with the same arithmetic a reference run on this device uses (torch.erfc on the GPU).
__device__ __forceinline__ float gc_lik_precise(float outputs, float mu, float s) {
  const float v = fabsf(__fsub_rn(outputs, mu));
  const float u = __fdiv_rn(__fsub_rn(0.5f, v), s);
  const float l = __fdiv_rn(__fsub_rn(-0.5f, v), s);
  const float up = 0.5f * erfcf(-0.70710678118654752440f * u);
  const float lo = 0.5f * erfcf(-0.70710678118654752440f * l);
  return __fsub_rn(up, lo);
}

struct GcArgs {
  const float* y;
  const float* mu;
  const float* sigma;
  const float* noise;
  long long n, y_bs, mu_bs, sg_bs;
  float scale_bound, lik_bound;
  int y_hat_mode;
  float* y_hat;
  float* lik;      // noisy (or the only) likelihood
  float* lik_q;    // dual only
  double* part;    // [B][gridDim.x] partial sums of log2(lik) (nullable)
  double* part_q;  // dual only
};

template <bool DUAL, bool PRECISE = false>
__device__ __forceinline__ void gc_element(const GcArgs& a, float y, float mu, float sg, float nz, bool train,
                                           float& y_hat, float& lik, float& lik_q) {
  const float s = fmaxf(sg, a.scale_bound);
  const float deq = __fadd_rn(rintf(__fsub_rn(y, mu)), mu);  // round(y - mu) + mu
  if (DUAL) {
    lik = fmaxf(gc_lik(__fadd_rn(y, nz), mu, s), a.lik_bound);
    lik_q = fmaxf(gc_lik(deq, mu, s), a.lik_bound);
    y_hat = deq;
  } else {
    const float outputs = train ? __fadd_rn(y, nz) : deq;
    lik = fmaxf(PRECISE ? gc_lik_precise(outputs, mu, s) : gc_lik(outputs, mu, s), a.lik_bound);
    y_hat = (a.y_hat_mode == 1) ? deq : outputs;
    lik_q = 0.f;
  }
}

template <bool DUAL, bool VEC, bool PRECISE = false>
__global__ void __launch_bounds__(kGcThreads) gc_forward_kernel(GcArgs a) {
  __shared__ double scratch[32];
  const long long b = blockIdx.y;
  const float* y = a.y + b * a.y_bs;
  const float* mu = a.mu ? a.mu + b * a.mu_bs : nullptr;
  const float* sg = a.sigma + b * a.sg_bs;
  const float* nz = a.noise ? a.noise + b * a.n : nullptr;
  float* yh = a.y_hat ? a.y_hat + b * a.n : nullptr;
  float* lk = a.lik ? a.lik + b * a.n : nullptr;
  float* lq = (DUAL && a.lik_q) ? a.lik_q + b * a.n : nullptr;
  const bool train = (nz != nullptr);
  const long long start = (long long)blockIdx.x * kGcChunk;
  float acc = 0.f, acc_q = 0.f;

  if (VEC) {
    float4 vy[kGcVecPerThread], vm[kGcVecPerThread], vs[kGcVecPerThread], vn[kGcVecPerThread];
#pragma unroll
    for (int i = 0; i < kGcVecPerThread; ++i) {
      const long long e = start + ((long long)i * kGcThreads + threadIdx.x) * 4;
      if (e < a.n) {
        vy[i] = ldg_stream(reinterpret_cast<const float4*>(y + e));
        vs[i] = ldg_stream(reinterpret_cast<const float4*>(sg + e));
        vm[i] = mu ? ldg_stream(reinterpret_cast<const float4*>(mu + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        vn[i] = nz ? ldg_stream(reinterpret_cast<const float4*>(nz + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < kGcVecPerThread; ++i) {
      const long long e = start + ((long long)i * kGcThreads + threadIdx.x) * 4;
      if (e < a.n) {
        float4 oy, ol, oq;
        gc_element<DUAL, PRECISE>(a, vy[i].x, vm[i].x, vs[i].x, vn[i].x, train, oy.x, ol.x, oq.x);
        gc_element<DUAL, PRECISE>(a, vy[i].y, vm[i].y, vs[i].y, vn[i].y, train, oy.y, ol.y, oq.y);
        gc_element<DUAL, PRECISE>(a, vy[i].z, vm[i].z, vs[i].z, vn[i].z, train, oy.z, ol.z, oq.z);
        gc_element<DUAL, PRECISE>(a, vy[i].w, vm[i].w, vs[i].w, vn[i].w, train, oy.w, ol.w, oq.w);
        if (yh) stg_stream(reinterpret_cast<float4*>(yh + e), oy);
        if (lk) stg_stream(reinterpret_cast<float4*>(lk + e), ol);
        if (DUAL && lq) stg_stream(reinterpret_cast<float4*>(lq + e), oq);
        if (a.part) acc += (lg2_ftz(ol.x) + lg2_ftz(ol.y)) + (lg2_ftz(ol.z) + lg2_ftz(ol.w));
        if (DUAL && a.part_q) acc_q += (lg2_ftz(oq.x) + lg2_ftz(oq.y)) + (lg2_ftz(oq.z) + lg2_ftz(oq.w));
      }
    }
  } else {
    for (int i = 0; i < kGcVecPerThread * 4; ++i) {
      const long long e = start + (long long)i * kGcThreads + threadIdx.x;
      if (e < a.n) {
        float oy, ol, oq;
        gc_element<DUAL, PRECISE>(a, y[e], mu ? mu[e] : 0.f, sg[e], nz ? nz[e] : 0.f, train, oy, ol, oq);
        if (yh) yh[e] = oy;
        if (lk) lk[e] = ol;
        if (DUAL && lq) lq[e] = oq;
        if (a.part) acc += lg2_ftz(ol);
        if (DUAL && a.part_q) acc_q += lg2_ftz(oq);
      }
    }
  }
  if (a.part) {
    const double s = block_sum((double)acc, scratch);
    if (threadIdx.x == 0) a.part[b * gridDim.x + blockIdx.x] = s;
  }
  if (DUAL && a.part_q) {
    const double s = block_sum((double)acc_q, scratch);
    if (threadIdx.x == 0) a.part_q[b * gridDim.x + blockIdx.x] = s;
  }
}

// bits[b] = -sum_j part[b][j]; one warp per sample, fixed order -> deterministic.
__global__ void __launch_bounds__(128) bits_finalize_kernel(const double* __restrict__ part, int per, long long B,
                                                             float* __restrict__ bits,
                                                             const double* __restrict__ part_q,
                                                             float* __restrict__ bits_q) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  if (part && bits) {
    double s = 0.0;
    for (int j = lane; j < per; j += 32) s += part[b * per + j];
    s = warp_sum(s);
    if (lane == 0) bits[b] = (float)(-s);
  }
  if (part_q && bits_q) {
    double s = 0.0;
    for (int j = lane; j < per; j += 32) s += part_q[b * per + j];
    s = warp_sum(s);
    if (lane == 0) bits_q[b] = (float)(-s);
  }
}

static inline bool gc_vec_ok(const void* p0, const void* p1, const void* p2, const void* p3, const void* p4,
                             const void* p5, const void* p6, long long n, long long s0, long long s1, long long s2) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al(p0) && al(p1) && al(p2) && al(p3) && al(p4) && al(p5) && al(p6) && (n % 4 == 0) && (s0 % 4 == 0) &&
         (s1 % 4 == 0) && (s2 % 4 == 0);
}

template <bool DUAL>
static int gc_launch(GcArgs a, long long B, float* bits, float* bits_q, void* workspace, size_t ws_bytes,
                     cudaStream_t s, bool precise = false) {
  const int per = ceil_div_i(a.n, kGcChunk);
  if (bits || bits_q) {
    const size_t need = (size_t)B * per * sizeof(double) * 2;
    if (!workspace || ws_bytes < need) return DCVIC_ERR_WORKSPACE;
    a.part = bits ? reinterpret_cast<double*>(workspace) : nullptr;
    a.part_q = bits_q ? reinterpret_cast<double*>(workspace) + (size_t)B * per : nullptr;
  }
  const bool vec = gc_vec_ok(a.y, a.mu, a.sigma, a.noise, a.y_hat, a.lik, a.lik_q, a.n, a.y_bs, a.mu_bs, a.sg_bs);
  dim3 grid(per, (unsigned)B);
  if (precise && !DUAL)
    gc_forward_kernel<false, false, true><<<grid, kGcThreads, 0, s>>>(a);
  else if (vec)
    gc_forward_kernel<DUAL, true><<<grid, kGcThreads, 0, s>>>(a);
  else
    gc_forward_kernel<DUAL, false><<<grid, kGcThreads, 0, s>>>(a);
  if (dcvic_launch_status() != DCVIC_OK) return DCVIC_ERR_CUDA;
  if (bits || bits_q) {
    bits_finalize_kernel<<<ceil_div_i(B, 4), 128, 0, s>>>(a.part, per, B, bits, a.part_q, bits_q);
  }
  return dcvic_launch_status();
}

// ------------------------------------------------------------------ backward (training mode)
// L = Phi(u) - Phi(l), u = (.5 - v)/s, l = (-.5 - v)/s, v = |y + noise - mu|, s = max(sigma, bound)
//   dL/dv = (phi(l) - phi(u))/s ; dL/dx = sign(x - mu) dL/dv ; dL/dmu = -dL/dx
//   dL/ds = (l phi(l) - u phi(u))/s ; LowerBound rules gate both bounds.
// One latent: the SFU forms of the forward kernel (Chebyshev erfc, ex2 for the pdf, rcp for 1/s); the likelihood is
// only needed for the LowerBound gate.
__device__ __forceinline__ void gc_backward_element(const GcArgs& a, float yv, float m, float raw, float nzv, bool train,
                                                    float go, float& dx, float& ds) {
  const float s = fmaxf(raw, a.scale_bound);
  // eval mode: outputs = round(y - mu) + mu, whose derivative wrt y and (net) mu is zero
  const float x = train ? __fadd_rn(yv, nzv) : __fadd_rn(rintf(__fsub_rn(yv, m)), m);
  const float diff = __fsub_rn(x, m);
  const float v = fabsf(diff);
  const float inv_s = rcp_ftz(s);
  const float u = (0.5f - v) * inv_s, l = (-0.5f - v) * inv_s;
  const float L = gc_lik(x, m, s);
  if (!(L >= a.lik_bound || go < 0.f)) go = 0.f;  // likelihood LowerBound
  const float pu = kInvSqrt2Pi * ex2_ftz(-0.72134752044448170368f * u * u);   // exp(-u^2 / 2)
  const float pl = kInvSqrt2Pi * ex2_ftz(-0.72134752044448170368f * l * l);
  const float dv = (pl - pu) * inv_s;
  const float sgn = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
  dx = train ? go * sgn * dv : 0.f;
  ds = go * (l * pl - u * pu) * inv_s;
  if (!(raw >= a.scale_bound || ds < 0.f)) ds = 0.f;  // scale LowerBound
}

template <bool VEC>
__global__ void __launch_bounds__(256) gc_backward_kernel(const float* __restrict__ g_lik, GcArgs a,
                                                           float* __restrict__ d_y, float* __restrict__ d_mu,
                                                           float* __restrict__ d_sigma) {
  const long long b = blockIdx.y;
  const float* y = a.y + b * a.y_bs;
  const float* mu = a.mu ? a.mu + b * a.mu_bs : nullptr;
  const float* sg = a.sigma + b * a.sg_bs;
  const float* nz = a.noise ? a.noise + b * a.n : nullptr;
  const float* g = g_lik + b * a.n;
  const bool train = nz != nullptr;
  if (VEC) {
    // 128-bit loads / stores: 28 B per latent (g, y, mu, sigma in; dy, dmu, dsigma out), HBM-bound
    for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < a.n;
         e += (long long)gridDim.x * blockDim.x * 4) {
      const float4 y4 = ldg_stream(reinterpret_cast<const float4*>(y + e));
      const float4 s4 = ldg_stream(reinterpret_cast<const float4*>(sg + e));
      const float4 g4 = ldg_stream(reinterpret_cast<const float4*>(g + e));
      const float4 m4 = mu ? ldg_stream(reinterpret_cast<const float4*>(mu + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 n4 = nz ? ldg_stream(reinterpret_cast<const float4*>(nz + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 dx, ds;
      gc_backward_element(a, y4.x, m4.x, s4.x, n4.x, train, g4.x, dx.x, ds.x);
      gc_backward_element(a, y4.y, m4.y, s4.y, n4.y, train, g4.y, dx.y, ds.y);
      gc_backward_element(a, y4.z, m4.z, s4.z, n4.z, train, g4.z, dx.z, ds.z);
      gc_backward_element(a, y4.w, m4.w, s4.w, n4.w, train, g4.w, dx.w, ds.w);
      const long long o = b * a.n + e;
      if (d_y) stg_stream(reinterpret_cast<float4*>(d_y + o), dx);
      if (d_mu) stg_stream(reinterpret_cast<float4*>(d_mu + o), make_float4(-dx.x, -dx.y, -dx.z, -dx.w));
      if (d_sigma) stg_stream(reinterpret_cast<float4*>(d_sigma + o), ds);
    }
  } else {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < a.n;
         e += (long long)gridDim.x * blockDim.x) {
      float dx, ds;
      gc_backward_element(a, y[e], mu ? mu[e] : 0.f, sg[e], nz ? nz[e] : 0.f, train, g[e], dx, ds);
      const long long o = b * a.n + e;
      if (d_y) d_y[o] = dx;
      if (d_mu) d_mu[o] = -dx;
      if (d_sigma) d_sigma[o] = ds;
    }
  }
}

__global__ void __launch_bounds__(256) gc_build_indexes_kernel(const float* __restrict__ sigma, long long n,
                                                                const float* __restrict__ table, int T, float bound,
                                                                int32_t* __restrict__ out) {
  extern __shared__ float s_table[];
  for (int i = threadIdx.x; i < T; i += blockDim.x) s_table[i] = table[i];
  __syncthreads();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float s = fmaxf(sigma[e], bound);
    // idx = (T-1) - #{ j < T-1 : s <= table[j] }  == #{ j < T-1 : table[j] < s }   (table ascending)
    int lo = 0, hi = T - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_table[mid] < s) lo = mid + 1; else hi = mid;
    }
    out[e] = lo;
  }
}

// Compress-side step of one CHARM slice (minnen20_charm_context_model.py:146-165 does these as four passes:
// entropy_model_y(y, [mu, sigma], is_train=False), build_indexes(sigma), quantize(y, "symbols", mu)): dequantized
// y_hat, its likelihood, the rANS symbol round(y - mu) and the scale-table index, from one read of y, mu, sigma.
// Batched like gc_forward: per-batch strides for tensors that are channel slices of larger ones.
template <bool VEC>
__global__ void __launch_bounds__(256) gc_codec_step_kernel(const float* __restrict__ y, const float* __restrict__ mu,
                                                             const float* __restrict__ sigma, long long n,
                                                             long long y_bs, long long mu_bs, long long sg_bs,
                                                             const float* __restrict__ table, int T,
                                                             float scale_bound, float lik_bound,
                                                             float* __restrict__ y_hat, float* __restrict__ lik,
                                                             int32_t* __restrict__ symbols,
                                                             int32_t* __restrict__ indexes) {
  extern __shared__ float s_table[];
  for (int i = threadIdx.x; i < T; i += blockDim.x) s_table[i] = table[i];
  __syncthreads();
  const long long b = blockIdx.y;
  y += b * y_bs;
  mu += b * mu_bs;
  sigma += b * sg_bs;
  const long long o = b * n;
  auto one = [&](float yv, float m, float sg, float& deq, float& lk, int& sym, int& ix) {
    const float s = fmaxf(sg, scale_bound);
    const float q = rintf(__fsub_rn(yv, m));              // round-half-even, as torch.round
    deq = __fadd_rn(q, m);
    lk = fmaxf(gc_lik(deq, m, s), lik_bound);
    sym = (int)q;
    int lo = 0, hi = T - 1;                               // #{ j < T-1 : table[j] < s }
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_table[mid] < s) lo = mid + 1; else hi = mid;
    }
    ix = lo;
  };
  if (VEC) {                                              // n % 4 == 0, all pointers and strides 16-byte aligned
    const long long n4 = n >> 2;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
      const float4 yv = ldg_stream(reinterpret_cast<const float4*>(y) + e);
      const float4 mv = ldg_stream(reinterpret_cast<const float4*>(mu) + e);
      const float4 sv = ldg_stream(reinterpret_cast<const float4*>(sigma) + e);
      float4 d, l;
      int4 sy, ix;
      one(yv.x, mv.x, sv.x, d.x, l.x, sy.x, ix.x);
      one(yv.y, mv.y, sv.y, d.y, l.y, sy.y, ix.y);
      one(yv.z, mv.z, sv.z, d.z, l.z, sy.z, ix.z);
      one(yv.w, mv.w, sv.w, d.w, l.w, sy.w, ix.w);
      const long long w = (o >> 2) + e;
      if (y_hat) stg_stream(reinterpret_cast<float4*>(y_hat) + w, d);
      if (lik) stg_stream(reinterpret_cast<float4*>(lik) + w, l);
      if (symbols) reinterpret_cast<int4*>(symbols)[w] = sy;
      if (indexes) reinterpret_cast<int4*>(indexes)[w] = ix;
    }
  } else {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
      float d, l;
      int sy, ix;
      one(y[e], mu[e], sigma[e], d, l, sy, ix);
      if (y_hat) y_hat[o + e] = d;
      if (lik) lik[o + e] = l;
      if (symbols) symbols[o + e] = sy;
      if (indexes) indexes[o + e] = ix;
    }
  }
}

__global__ void __launch_bounds__(256) ste_round_kernel(const float* __restrict__ x, long long n,
                                                         float* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = x[e];
    out[e] = __fadd_rn(__fsub_rn(rintf(v), v), v);
  }
}

}  // namespace dcvic

using namespace dcvic;

extern "C" size_t dcvic_gc_workspace_bytes(int64_t B, int64_t n) {
  if (B <= 0 || n <= 0) return 0;
  return align_up((size_t)B * (size_t)ceil_div_i(n, kGcChunk) * sizeof(double) * 2, 256);
}

static int gc_check(const float* y, const float* sigma, int64_t B, int64_t n, float scale_bound) {
  DCVIC_CHECK_ARG(y && sigma);
  DCVIC_CHECK_ARG(B > 0 && n > 0 && B <= 65535);
  DCVIC_CHECK_ARG(scale_bound > 0.f);
  return DCVIC_OK;
}

extern "C" int dcvic_gc_forward(const float* y, const float* mu, const float* sigma, const float* noise, int64_t B,
                                int64_t n, int64_t y_bstride, int64_t mu_bstride, int64_t sigma_bstride,
                                float scale_bound, float lik_bound, int y_hat_mode, float* y_hat, float* lik,
                                float* bits, void* workspace, size_t ws_bytes, dcvic_stream_t stream) {
  int rc = gc_check(y, sigma, B, n, scale_bound);
  if (rc) return rc;
  DCVIC_CHECK_ARG(y_hat || lik || bits);
  DCVIC_CHECK_ARG((y_hat_mode & ~(1 | DCVIC_GC_PRECISE)) == 0);
  GcArgs a{y, mu, sigma, noise, n, y_bstride, mu_bstride, sigma_bstride, scale_bound, lik_bound, y_hat_mode & 1,
           y_hat, lik, nullptr, nullptr, nullptr};
  return gc_launch<false>(a, B, bits, nullptr, workspace, ws_bytes, (cudaStream_t)stream,
                          (y_hat_mode & DCVIC_GC_PRECISE) != 0);
}

extern "C" int dcvic_gc_forward_dual(const float* y, const float* mu, const float* sigma, const float* noise,
                                     int64_t B, int64_t n, int64_t y_bstride, int64_t mu_bstride,
                                     int64_t sigma_bstride, float scale_bound, float lik_bound, float* y_hat,
                                     float* lik_noisy, float* lik_q, float* bits_noisy, float* bits_q, void* workspace,
                                     size_t ws_bytes, dcvic_stream_t stream) {
  int rc = gc_check(y, sigma, B, n, scale_bound);
  if (rc) return rc;
  DCVIC_CHECK_ARG(noise != nullptr);
  GcArgs a{y, mu, sigma, noise, n, y_bstride, mu_bstride, sigma_bstride, scale_bound, lik_bound, 1,
           y_hat, lik_noisy, lik_q, nullptr, nullptr};
  return gc_launch<true>(a, B, bits_noisy, bits_q, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dcvic_gc_backward(const float* g_lik, const float* y, const float* mu, const float* sigma,
                                 const float* noise, int64_t B, int64_t n, int64_t y_bstride, int64_t mu_bstride,
                                 int64_t sigma_bstride, float scale_bound, float lik_bound, float* d_y, float* d_mu,
                                 float* d_sigma, dcvic_stream_t stream) {
  int rc = gc_check(y, sigma, B, n, scale_bound);
  if (rc) return rc;
  DCVIC_CHECK_ARG(g_lik != nullptr);
  DCVIC_CHECK_ARG(d_y || d_mu || d_sigma);
  GcArgs a{y, mu, sigma, noise, n, y_bstride, mu_bstride, sigma_bstride, scale_bound, lik_bound, 0,
           nullptr, nullptr, nullptr, nullptr, nullptr};
  auto mis = [](const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
  const bool vec = n % 4 == 0 && y_bstride % 4 == 0 && mu_bstride % 4 == 0 && sigma_bstride % 4 == 0 && !mis(g_lik) &&
                   !mis(y) && !mis(mu) && !mis(sigma) && !mis(noise) && !mis(d_y) && !mis(d_mu) && !mis(d_sigma);
  if (vec) {
    dim3 grid(min(ceil_div_i(n, 1024), 4 * kNumSMs), (unsigned)B);
    gc_backward_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(g_lik, a, d_y, d_mu, d_sigma);
  } else {
    dim3 grid(min(ceil_div_i(n, 256), 2048), (unsigned)B);
    gc_backward_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(g_lik, a, d_y, d_mu, d_sigma);
  }
  return dcvic_launch_status();
}

extern "C" int dcvic_gc_build_indexes(const float* sigma, int64_t n, const float* table, int T, float scale_bound,
                                      int32_t* out, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(sigma && table && out);
  DCVIC_CHECK_ARG(n > 0 && T >= 1 && T <= 4096);
  gc_build_indexes_kernel<<<min(ceil_div_i(n, 256), 8 * kNumSMs), 256, T * sizeof(float), (cudaStream_t)stream>>>(
      sigma, n, table, T, scale_bound, out);
  return dcvic_launch_status();
}

extern "C" int dcvic_gc_codec_step(const float* y, const float* mu, const float* sigma, int64_t B, int64_t n,
                                   int64_t y_bstride, int64_t mu_bstride, int64_t sigma_bstride, const float* table,
                                   int T, float scale_bound, float lik_bound, float* y_hat, float* lik,
                                   int32_t* symbols, int32_t* indexes, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(y && mu && sigma && table);
  DCVIC_CHECK_ARG(B > 0 && n > 0 && T >= 1 && T <= 4096 && B <= 65535);
  DCVIC_CHECK_ARG(y_hat || lik || symbols || indexes);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = n % 4 == 0 && y_bstride % 4 == 0 && mu_bstride % 4 == 0 && sigma_bstride % 4 == 0 && al16(y) &&
                   al16(mu) && al16(sigma) && al16(y_hat) && al16(lik) && al16(symbols) && al16(indexes);
  const long long per = vec ? n / 4 : n;
  dim3 grid((unsigned)min(ceil_div_i(per, 256), 16 * kNumSMs), (unsigned)B);
  if (vec)
    gc_codec_step_kernel<true><<<grid, 256, T * sizeof(float), (cudaStream_t)stream>>>(
        y, mu, sigma, n, y_bstride, mu_bstride, sigma_bstride, table, T, scale_bound, lik_bound, y_hat, lik, symbols,
        indexes);
  else
    gc_codec_step_kernel<false><<<grid, 256, T * sizeof(float), (cudaStream_t)stream>>>(
        y, mu, sigma, n, y_bstride, mu_bstride, sigma_bstride, table, T, scale_bound, lik_bound, y_hat, lik, symbols,
        indexes);
  return dcvic_launch_status();
}

extern "C" int dcvic_ste_round(const float* x, int64_t n, float* out, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(x && out && n > 0);
  ste_round_kernel<<<min(ceil_div_i(n, 256), 8 * kNumSMs), 256, 0, (cudaStream_t)stream>>>(x, n, out);
  return dcvic_launch_status();
}
