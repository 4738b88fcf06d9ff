// Host-side integer CDF construction for the entropy-coder tables (model-setup time, like the
// reference's C++ extension): compressai==1.2.4 `_CXX.pmf_to_quantized_cdf`, reached from
// EntropyModel._pmf_to_cdf <- EntropyBottleneck.update / GaussianConditional.update
// (callers in iwa-shi/DC_VIC: src/models/comp_model/hyperprior_dc_vic_model.py:65-68).
// Runs on the CPU in the reference too; this is not a fallback of a GPU path.
#include "common.cuh"
#include <cmath>
#include <vector>

extern "C" int dcvic_pmf_to_quantized_cdf(const float* pmf /*host*/, int n, int precision, int32_t* cdf /*host n+1*/) {
  if (!pmf || !cdf || n < 1 || precision < 1 || precision > 30) return DCVIC_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i)
    if (!std::isfinite(pmf[i]) || pmf[i] < 0.f) return DCVIC_ERR_BAD_ARG;
  const uint64_t one = 1ull << precision;
  std::vector<uint64_t> c((size_t)n + 1, 0);
  uint64_t total = 0;
  for (int i = 0; i < n; ++i) {
    const float scaled = pmf[i] * (float)one;         // float product, then round half away from zero
    c[i + 1] = (uint64_t)std::llround((double)scaled);
    total += c[i + 1];
  }
  if (total == 0) return DCVIC_ERR_BAD_ARG;
  uint64_t run = 0;
  for (int i = 0; i <= n; ++i) {
    run += (one * c[i]) / total;
    c[i] = run;
  }
  c[n] = one;
  for (int i = 0; i < n; ++i) {
    if (c[i] != c[i + 1]) continue;
    // zero-width symbol: take one count from the narrowest symbol that can spare it
    uint64_t best_w = ~0ull;
    int donor = -1;
    for (int j = 0; j < n; ++j) {
      const uint64_t w = c[j + 1] - c[j];
      if (w > 1 && w < best_w) { best_w = w; donor = j; }
    }
    if (donor < 0) return DCVIC_ERR_UNSUPPORTED;
    if (donor < i) for (int j = donor + 1; j <= i; ++j) c[j]--;
    else           for (int j = i + 1; j <= donor; ++j) c[j]++;
  }
  for (int i = 0; i <= n; ++i) cdf[i] = (int32_t)c[i];
  return DCVIC_OK;
}

// The same construction for a whole table on the device: one thread per row (rows are independent; a row is a short
// sequential integer recurrence), reading the PMF where the likelihood kernel left it.  cdf [rows, width + 2] int32.
namespace dcvic {
__global__ void __launch_bounds__(64) pmf_to_cdf_rows_kernel(const float* __restrict__ pmf, int width,
                                                              const float* __restrict__ tail,
                                                              const int32_t* __restrict__ lengths, int rows,
                                                              int precision, int32_t* __restrict__ cdf,
                                                              int32_t* __restrict__ status) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int len = lengths[r], n = len + 1;                  // + the tail-mass symbol
  const float* p = pmf + (size_t)r * width;
  int32_t* c = cdf + (size_t)r * (width + 2);
  const unsigned long long one = 1ull << precision;
  auto prob = [&](int i) { return i < len ? p[i] : tail[r]; };
  auto count = [&](int i) {                                 // float product, then round half away from zero
    const float scaled = prob(i) * (float)one;
    return (unsigned long long)llround((double)scaled);
  };
  unsigned long long total = 0;
  for (int i = 0; i < n; ++i) total += count(i);
  for (int i = n + 1; i < width + 2; ++i) c[i] = 0;
  if (total == 0 || len < 1 || len > width) { atomicExch(status, DCVIC_ERR_BAD_ARG); return; }
  unsigned long long run = 0;
  c[0] = 0;
  for (int i = 0; i < n; ++i) {
    run += (one * count(i)) / total;
    c[i + 1] = (int32_t)run;
  }
  c[n] = (int32_t)one;
  for (int i = 0; i < n; ++i) {
    if (c[i] != c[i + 1]) continue;
    int best_w = 0x7fffffff, donor = -1;
    for (int j = 0; j < n; ++j) {
      const int w = c[j + 1] - c[j];
      if (w > 1 && w < best_w) { best_w = w; donor = j; }
    }
    if (donor < 0) { atomicExch(status, DCVIC_ERR_UNSUPPORTED); return; }
    if (donor < i) for (int j = donor + 1; j <= i; ++j) c[j]--;
    else           for (int j = i + 1; j <= donor; ++j) c[j]++;
  }
}
}  // namespace dcvic

extern "C" int dcvic_pmf_to_quantized_cdf_rows(const float* pmf, int rows, int width, const float* tail_mass,
                                               const int32_t* lengths, int precision, int32_t* cdf, int32_t* status,
                                               dcvic_stream_t stream) {
  if (!pmf || !tail_mass || !lengths || !cdf || !status || rows < 1 || width < 1 || precision < 1 || precision > 30)
    return DCVIC_ERR_BAD_ARG;
  dcvic::pmf_to_cdf_rows_kernel<<<(rows + 63) / 64, 64, 0, (cudaStream_t)stream>>>(pmf, width, tail_mass, lengths, rows,
                                                                                    precision, cdf, status);
  return dcvic_launch_status();
}
