// Internal interfaces between the VQ translation units.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>
#include <string.h>

namespace dcvic {

constexpr int kFinishTokens = 32;     // tokens per CTA in the finish (re-rank + gather + STE + loss) kernel
constexpr int kListCap = 16;          // (chunk key, flag mask) entries per token and accumulator buffer
constexpr int kChunk = 32;            // codes per flag mask (one tcgen05.ld.32x32b.x32 per row)
constexpr int kCb16Pad = 64;          // extra FP16 codebook columns: a 3-way split of -|e|^2/2 in the first three
constexpr int kCandMax = 24;          // FP32 re-rank candidates per token before falling back to a full scan

// What the tensor search hands to the finish kernel, per token t: one VqMeta record and two lists.
// A token tile of the search's last round may be scanned by two CTA pairs, each over half of the codebook ("split"):
// part 0 then uses entries 0-7 of each list and the first half of the record, part 1 entries 8-15 and the second.
struct __align__(16) VqMeta {
  float m0, m1;     // running maximum of the FP16 score over columns 0-127 / 128-255 of every accumulator (part 0)
  short n0, n1;     // list entries of epilogue warp quad 0 / 1, or -1: scan the whole codebook for this token
  float zz;         // |z|^2 as the search computed it (its margin and the finish's threshold use the same value)
  float mb0, mb1;   // part 1 (split tiles; -inf / 0 otherwise)
  short nb0, nb1;
  unsigned split;   // 1: lists are two halves of 8 entries
};
//   list[(2t + q) * kListCap + i] = { key, mask }:  key = (bits(chunk max) & ~0x7F) | chunk id,
//                           mask bit j set <=> score of code (chunk id * 32 + j) was within the
//                           margin of the running maximum when that chunk went by
struct VqWorkspace {
  size_t off_counters, off_ee, off_nhee, off_emax, off_partials, off_hist, off_cand, off_meta, off_list, off_cb16,
      off_eperm, total;
  int n_tokens;
};

inline VqWorkspace vq_workspace_layout(int B, int D, int HW, int K) {
  VqWorkspace w{};
  const size_t N = (size_t)B * HW;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
  w.off_counters = take(64 * sizeof(unsigned));
  w.off_ee = take((size_t)K * sizeof(float));
  w.off_nhee = take((size_t)K * sizeof(float));
  w.off_emax = take(4 * sizeof(float));
  // loss partials: one per finish warp (4 tokens each), 16 per persistent CTA of the TMA finish, or 24 per CTA of the
  // single-pass kernel (consumer warps + helping epilogue warps)
  w.off_partials = take((N / 4 + 8 + (size_t)kNumSMs * 32) * sizeof(double));
  w.off_hist = take((size_t)K * sizeof(unsigned));
  w.off_cand = take(N * sizeof(int));
  w.off_meta = take(N * sizeof(VqMeta));
  w.off_list = take(N * 2 * kListCap * sizeof(uint2));
  w.off_cb16 = take((size_t)K * (D + kCb16Pad) * sizeof(__half));
  w.off_eperm = take((size_t)K * D * sizeof(float));      // the codebook in the single-pass kernel's gather order
  w.total = o;
  w.n_tokens = (int)N;
  return w;
}

// counters[] slots
enum { kCtrLoss = 0, kCtrOverflow = 1, kCtrPerp = 2, kCtrRerank = 3, kCtrTotalCand = 4, kCtrPrep = 5 };

// The proven bound on |fp16 score - exact score| differences that the search and the finish must agree on.
// Both operands are rounded to nearest FP16 (relative 2^-11 each, so 2^-10 + 2^-22 per product, summed with
// Cauchy-Schwarz over channels) and the bound applies to the maximum and to the candidate (x2); 2 % slack for
// the tensor core's FP32 accumulation and the 7 mantissa bits dropped from the stored chunk maximum; an
// absolute term for operands in FP16's subnormal range (2^-25 each, e_dim <= 256); an absolute 4 * 2^-25 for the
// three-way FP16 split of -|e|^2/2, whose last piece is rounded on FP16's 2^-24 subnormal grid whenever |e|^2/2 < 1/4
// (up to 2^-25 per code, for the maximum and for the candidate, x2: it is what decides between codes for tokens
// with |z| << |e|); plus a few FP32 ulps of the reference distance itself (ties created by its rounding).
// Tokens with |z|^2 >= kVqFp16Zz2Max (an element could exceed FP16's range) and codebooks flagged by the
// prepare kernel (emax[1] != 0) are not searched on the tensor cores at all: they take the full FP32 scan.
constexpr float kVqFp16Zz2Max = 3.6e9f;       // (6e4)^2
__host__ __device__ __forceinline__ float vq_margin(float zz, float emax) {
  const float nz = sqrtf(zz);
  return 1.02f * 0.0019536f * nz * emax + 9.6e-7f * (nz + emax) + 1.9e-6f * (zz + emax * emax) + 1.2e-7f;
}

// The single-pass kernel's margin: the same bound from MEASURED rounding residuals instead of worst-case ones.
// With h = fp16(z), g_k = fp16(e_k), dz = |z - h|, de = max_k |e_k - g_k| (both computed exactly, so FP16 subnormals
// need no term of their own):  |z.e_k - h.g_k| <= dz (|e_k| + de) + |z| de  (Cauchy-Schwarz), the FP16 x FP16 products
// are exact in FP32 and their FP32 accumulation over D terms is off by at most D 2^-23 |z| |e_k|; the bound applies
// to the maximum and to the candidate (x2).  On top: 4 ulps of the reference's own FP32 distance (so that every code
// the reference's rounding could prefer is re-ranked; anything closer than that is a documented near-tie, 1e-6
// relative being 8 ulps) and 2 % slack.
__host__ __device__ __forceinline__ float vq_margin_measured(float zz, float dz2, float emax, float demax, int D) {
  const float nz = sqrtf(zz), dz = sqrtf(dz2) * 1.000001f;
  // (+ the constant's way into the accumulator: three FP16 pieces, exact to 2^-25 + 2^-33 |e|^2/2, carried through
  // e_dim / 16 + 1 FP32 accumulation steps at up to 2^-23 of |e|^2/2 each - x2, maximum and candidate)
  return 2.04f * (dz * (emax + demax) + nz * demax) + (float)D * 2.4e-7f * nz * emax + 4.8e-7f * (zz + emax * emax) +
         (float)(D / 16 + 1) * 1.2e-7f * emax * emax + 1.2e-7f;
}

// Upper bound of the chunk maximum stored in a list key.  The search keeps the top 25 bits of the FP32 chunk maximum
// (the low 7 carry the chunk id), i.e. truncates its magnitude: filling the 7 bits with ones rounds a positive value
// up, but a NEGATIVE one (scores z.e - |e|^2/2 are negative when |z| << |e|) further down - for those the truncated
// value itself is the upper bound.
__host__ __device__ __forceinline__ float vq_key_upper(unsigned key) {
  const unsigned bits = (key & 0x80000000u) ? (key & 0xFFFFFF80u) : (key | 0x7Fu);
#ifdef __CUDA_ARCH__
  return __uint_as_float(bits);
#else
  float f;
  memcpy(&f, &bits, sizeof f);
  return f;
#endif
}

// vq_simt.cu
// scratch: K floats (per-CTA maxima of the prepare kernel)
// eperm (optional): the FP32 codebook with the channels of every group of four rotated the way the single-pass
// kernel's consumers visit them (vq_eperm_dest)
int vq_prepare_codebook(const float* codebook, int K, int D, float* ee, float* scratch, float* emax,
                        __half* cb16, float* eperm, unsigned* counters, cudaStream_t s);
// Consumer lane l = (c >> 2) & 31 holds channels 4 (c >> 2) .. + 3 and visits them in the order (j + (l >> 1)) & 3
// (conflict-free 16-byte accesses to the swizzled z stage): channel c goes to slot j = (c - (l >> 1)) & 3 of its group.
__host__ __device__ __forceinline__ int vq_eperm_dest(int c) {
  const int b = c >> 2, rot = ((b & 31) >> 1) & 3;
  return 4 * b + (((c & 3) - rot) & 3);
}
int vq_narrow_forward(const float* z, const float* E, int B, int D, int HW, int K, float beta, int legacy, float* zq,
                      int64_t* idx, float* loss, double* partials, unsigned* counters, cudaStream_t s);
int vq_exact_search(const float* z, const float* E, const float* ee, int B, int D, int HW, int K, int* cand,
                    cudaStream_t s);
// cand != nullptr: one decided index per token (exact search).  Otherwise meta/list from the tensor search.
int vq_finish(const float* z, const float* E, const float* ee, const float* emax, const int* cand, const VqMeta* meta,
              const uint2* list, int B, int D, int HW, int K, float beta, int legacy, float* zq, int64_t* idx,
              float* loss, double* partials, unsigned* counters, cudaStream_t s);
int vq_launch_loss_finalize(const double* partials, int n, long long numel, float beta, int legacy, float* loss,
                            cudaStream_t s);
int vq_v1_extras(const int64_t* idx, int N, int K, float* onehot, float* perplexity, unsigned* hist, unsigned* counters,
                 cudaStream_t s);

// vq_finish_tma.cu: the finish for H*W % 32 == 0 and e_dim in {64, 128, 192, 256} (same arguments as vq_finish)
bool vq_finish_tma_supported(const float* z, const float* zq, int D, int HW, int K);
int vq_finish_tma(const float* z, const float* E, const float* ee, const float* emax, const int* cand,
                  const VqMeta* meta, const uint2* list, int B, int D, int HW, int K, float beta, int legacy,
                  float* zq, int64_t* idx, float* loss, double* partials, unsigned* counters, cudaStream_t s);

// vq_tcgen05.cu
bool vq_tensor_supported(int D, int K);
// after_prepare: the kernel launched just before on `s` was vq_prepare_codebook's (else the search waits for its
// predecessor before it reads z)
// allow_split: the finish kernel that will read meta / list understands split tiles (the TMA finish does)
int vq_tensor_search(const float* z, const __half* cb16, const float* emax, int B, int D, int HW, int K,
                     bool after_prepare, bool allow_split, VqMeta* meta, uint2* list, cudaStream_t s);

// vq_fused.cu: the single-pass wide forward (search + re-rank + gather + straight-through value + loss in one
// persistent kernel) for e_dim in {64, 128, 192, 256}, K % 128 == 0, K <= 2048, H*W % 32 == 0
bool vq_fused_supported(const float* z, const float* zq, const float* E, int D, int HW, int K);
int vq_fused_forward(const float* z, const float* Eperm, const float* ee, const float* emax, const __half* cb16, int B,
                     int D, int HW, int K, bool after_prepare, float beta, int legacy, float* zq, int64_t* idx,
                     float* loss, double* partials, unsigned* counters, cudaStream_t s);

// Shared tail: turn sum((e-z)^2) into the reference's loss scalar (taming quantize.py:291-296: legacy puts beta on
// the other term; the forward VALUE is mean * (1 + beta) either way, summed in the reference's order).
__device__ __forceinline__ void write_loss(double total, long long numel, float beta, int legacy, float* loss) {
  const float m = (float)(total / (double)numel);
  *loss = legacy ? __fadd_rn(m, __fmul_rn(beta, m)) : __fadd_rn(__fmul_rn(beta, m), m);
}

}  // namespace dcvic
