// Internal interfaces between the VQ translation units.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace dcvic {

constexpr int kCandCap = 16;          // candidate slots per token handed from the tensor search to the FP32 re-rank
constexpr int kFinishTokens = 32;     // tokens per CTA in the finish (re-rank + gather + STE + loss) kernel
constexpr int kTcK16Pad = 16;         // extra K columns of the BF16 codebook that fold -|e|^2/2 into the MMA

struct VqWorkspace {
  size_t off_counters, off_ee, off_emax, off_partials, off_hist, off_cand, off_count, off_cb16, total;
  int n_tokens, dpad16;
};

inline VqWorkspace vq_workspace_layout(int B, int D, int HW, int K) {
  VqWorkspace w{};
  const size_t N = (size_t)B * HW;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
  w.off_counters = take(64 * sizeof(unsigned));
  w.off_ee = take((size_t)K * sizeof(float));
  w.off_emax = take(4 * sizeof(float));
  w.off_partials = take((N / kFinishTokens + 2) * sizeof(double));
  w.off_hist = take((size_t)K * sizeof(unsigned));
  w.off_cand = take(N * kCandCap * sizeof(int));
  w.off_count = take(N * sizeof(int));
  w.dpad16 = D + kTcK16Pad;
  w.off_cb16 = take((size_t)K * w.dpad16 * sizeof(__nv_bfloat16));
  w.total = o;
  w.n_tokens = (int)N;
  return w;
}

// counters[] slots
enum { kCtrLoss = 0, kCtrOverflow = 1, kCtrPerp = 2, kCtrRerank = 3, kCtrTotalCand = 4 };

// vq_simt.cu
int vq_prepare_codebook(const float* codebook, int K, int D, float* ee, float* emax, __nv_bfloat16* cb16, int dpad16,
                        cudaStream_t s);
int vq_narrow_forward(const float* z, const float* E, int B, int D, int HW, int K, float beta, int legacy, float* zq,
                      int64_t* idx, float* loss, double* partials, unsigned* counters, cudaStream_t s);
int vq_exact_search(const float* z, const float* E, const float* ee, int B, int D, int HW, int K, int* cand,
                    cudaStream_t s);
int vq_finish(const float* z, const float* E, const float* ee, const int* cand, int cap, const int* count, int B, int D,
              int HW, int K, float beta, int legacy, float* zq, int64_t* idx, float* loss, double* partials,
              unsigned* counters, cudaStream_t s);
int vq_v1_extras(const int64_t* idx, int N, int K, float* onehot, float* perplexity, unsigned* hist, unsigned* counters,
                 cudaStream_t s);

// vq_tcgen05.cu
bool vq_tensor_supported(int D, int K);
int vq_tensor_search(const float* z, const __nv_bfloat16* cb16, int dpad16, const float* ee, const float* emax, int B,
                     int D, int HW, int K, int* cand, int* count, unsigned* counters, cudaStream_t s);

}  // namespace dcvic
