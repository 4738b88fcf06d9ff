// Range-ANS bitstream coder on the GPU: the coder DC-VIC writes its .bin files with is
// compressai.ans (CompressAI 1.2.4: rans_interface.cpp over ryg_rans' rans64.h - 64-bit state, 32-bit
// renormalisation words, 16-bit probabilities, 4-bit bypass digits for out-of-range values, symbols pushed in reverse).
// Call sites in the reference: EntropyModel.compress / decompress (hyperprior_dc_vic_model.py:308-328,378-387) and the
// CHARM decode loop (minnen20_charm_context_model.py:175-202).  A stream is an inherently sequential integer state
// machine, so parallelism is ACROSS streams (one warp per stream: images of a batch, the z and y strings) and, inside
// a stream, between the table look-ups (all lanes) and the state update (one lane).  Bit-exact against
// oracle/rans_oracle.c; no host round trip of the symbols (the reference marshals them through Python lists).
#include "common.cuh"

namespace dcvic {

constexpr uint32_t kRansPrecision = 16;
constexpr uint32_t kBypassBits = 4;
constexpr uint32_t kBypassMax = (1u << kBypassBits) - 1;
constexpr unsigned long long kRansL = 1ull << 31;

struct RansPut {
  unsigned long long x;
  uint32_t* ptr;       // next free word is ptr[-1]: the encoder writes backwards
  __device__ __forceinline__ void put(uint32_t start, uint32_t freq) {          // Rans64EncPut, scale_bits = 16
    const unsigned long long x_max = ((kRansL >> kRansPrecision) << 32) * freq;
    if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
    x = ((x / freq) << kRansPrecision) + (x % freq) + start;
  }
  __device__ __forceinline__ void put_bits(uint32_t val) {                      // Rans64EncPutBits, 4 bits
    const unsigned long long x_max = ((kRansL >> 16) << 32) * (unsigned long long)(1u << (16 - kBypassBits));
    if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
    x = (x << kBypassBits) | val;
  }
};

// One warp per stream.  Symbols are visited from the last to the first, 32 at a time: every lane looks its symbol up
// (value, CDF interval, bypass code), lane 0 then feeds them to the state machine in reverse order.
__global__ void __launch_bounds__(32) rans_encode_kernel(const int32_t* __restrict__ symbols,
                                                          const int32_t* __restrict__ indexes,
                                                          const long long* __restrict__ starts,
                                                          const int32_t* __restrict__ cdf, int width,
                                                          const int32_t* __restrict__ cdf_sizes,
                                                          const int32_t* __restrict__ offsets, uint32_t* __restrict__ out,
                                                          const long long* __restrict__ out_starts,
                                                          int32_t* __restrict__ nwords) {
  __shared__ uint32_t s_start[32], s_range[32], s_raw[32];
  __shared__ int s_nb[32];
  const int s = blockIdx.x, lane = threadIdx.x;
  const long long lo = starts[s], hi = starts[s + 1];
  uint32_t* const end = out + out_starts[s + 1];
  RansPut st{kRansL, end};
  for (long long base = ((hi - lo - 1) / 32) * 32 + lo; base >= lo; base -= 32) {
    const long long i = base + lane;
    int nb = -1;                        // -1: no symbol in this lane; 0: plain; > 0: bypass digits
    uint32_t start = 0, range = 1, raw = 0;
    if (i < hi) {
      const int32_t ci = indexes[i];
      const int32_t* row = cdf + (size_t)ci * width;
      const int32_t max_value = cdf_sizes[ci] - 2;
      int32_t value = symbols[i] - offsets[ci];
      nb = 0;
      if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
      else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
      start = (uint32_t)row[value] & 0xFFFFu;
      range = ((uint32_t)(row[value + 1] - row[value])) & 0xFFFFu;
      if (value == max_value) {
        int n = 0;
        while (n < 8 && (raw >> (n * kBypassBits)) != 0) ++n;
        nb = n + 1;                     // + 1: "is a bypass symbol" even when the raw value is 0
      }
    }
    s_start[lane] = start; s_range[lane] = range; s_raw[lane] = raw; s_nb[lane] = nb;
    __syncwarp();
    if (lane == 0) {
      for (int j = 31; j >= 0; --j) {
        const int b = s_nb[j];
        if (b < 0) continue;
        if (b > 0) {
          const int n_bypass = b - 1;
          const uint32_t rv = s_raw[j];
          for (int d = n_bypass - 1; d >= 0; --d) st.put_bits((rv >> (d * kBypassBits)) & kBypassMax);
          // the digit count: forward order is (15)* then the remainder, so the remainder goes in first here
          st.put_bits((uint32_t)(n_bypass % (int)kBypassMax));
          for (int k = 0; k < n_bypass / (int)kBypassMax; ++k) st.put_bits(kBypassMax);
        }
        st.put(s_start[j], s_range[j]);
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    // Rans64EncFlush
    st.ptr -= 2;
    st.ptr[0] = (uint32_t)(st.x >> 0);
    st.ptr[1] = (uint32_t)(st.x >> 32);
    nwords[s] = (int32_t)(end - st.ptr);
  }
}

__device__ __forceinline__ uint32_t rans_get_bits(unsigned long long& x, const uint32_t* words, long long& pos) {
  const uint32_t val = (uint32_t)(x & kBypassMax);
  x >>= kBypassBits;
  if (x < kRansL) { x = (x << 32) | words[pos]; ++pos; }
  return val;
}

// One thread per call: the decoder of ONE stream advances by n symbols; state[0] = rANS state, state[1] = read
// position (words), state[2] = initialised.  (The CHARM loop decodes slice by slice, each slice's probabilities
// depending on the previous slices' values, so the calls are sequential by construction.)
__global__ void __launch_bounds__(32) rans_decode_kernel(const uint32_t* __restrict__ words, long long nw,
                                                          long long* __restrict__ state,
                                                          const int32_t* __restrict__ indexes, long long n,
                                                          const int32_t* __restrict__ cdf, int width,
                                                          const int32_t* __restrict__ cdf_sizes,
                                                          const int32_t* __restrict__ offsets, int32_t* __restrict__ out) {
  if (threadIdx.x != 0) return;
  unsigned long long x = (unsigned long long)state[0];
  long long pos = state[1];
  if (!state[2]) {                      // Rans64DecInit
    x = (unsigned long long)words[0] | ((unsigned long long)words[1] << 32);
    pos = 2;
    state[2] = 1;
  }
  for (long long i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    const int32_t* row = cdf + (size_t)ci * width;
    const int32_t size = cdf_sizes[ci], max_value = size - 2;
    const uint32_t cum = (uint32_t)(x & 0xFFFFu);
    // s = (first v with cdf[v] > cum) - 1: binary search over the ascending row
    int lo = 0, hi = size - 1;          // invariant: row[lo] <= cum, row[hi] > cum  (row[size - 1] = 65536)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((uint32_t)__ldg(row + mid) <= cum) lo = mid; else hi = mid;
    }
    const uint32_t start = (uint32_t)__ldg(row + lo), freq = (uint32_t)__ldg(row + lo + 1) - start;
    x = (unsigned long long)freq * (x >> kRansPrecision) + (x & 0xFFFFu) - start;          // Rans64DecAdvance
    if (x < kRansL && pos < nw) { x = (x << 32) | words[pos]; ++pos; }
    int32_t value = lo;
    if (value == max_value) {
      int32_t val = (int32_t)rans_get_bits(x, words, pos);
      int32_t n_bypass = val;
      while (val == (int32_t)kBypassMax) { val = (int32_t)rans_get_bits(x, words, pos); n_bypass += val; }
      int32_t raw = 0;
      for (int32_t j = 0; j < n_bypass; ++j) raw |= (int32_t)rans_get_bits(x, words, pos) << (j * kBypassBits);
      value = raw >> 1;
      if (raw & 1) value = -value - 1; else value += max_value;
    }
    out[i] = value + offsets[ci];
  }
  state[0] = (long long)x;
  state[1] = pos;
}

}  // namespace dcvic

using namespace dcvic;

extern "C" int dcvic_rans_encode(const int32_t* symbols, const int32_t* indexes, const int64_t* starts, int n_streams,
                                 const int32_t* cdf, int rows, int width, const int32_t* cdf_sizes,
                                 const int32_t* offsets, uint32_t* out_words, const int64_t* out_starts,
                                 int32_t* n_words, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(symbols && indexes && starts && cdf && cdf_sizes && offsets && out_words && out_starts && n_words);
  DCVIC_CHECK_ARG(n_streams > 0 && rows > 0 && width >= 2);
  rans_encode_kernel<<<n_streams, 32, 0, (cudaStream_t)stream>>>(
      symbols, indexes, reinterpret_cast<const long long*>(starts), cdf, width, cdf_sizes, offsets, out_words,
      reinterpret_cast<const long long*>(out_starts), n_words);
  return dcvic_launch_status();
}

extern "C" int dcvic_rans_decode(const uint32_t* words, int64_t n_words, int64_t* state, const int32_t* indexes,
                                 int64_t n, const int32_t* cdf, int rows, int width, const int32_t* cdf_sizes,
                                 const int32_t* offsets, int32_t* out_symbols, dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(words && state && indexes && cdf && cdf_sizes && offsets && out_symbols);
  DCVIC_CHECK_ARG(n_words >= 2 && n >= 0 && rows > 0 && width >= 2);
  if (n == 0) return DCVIC_OK;
  rans_decode_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(words, n_words, reinterpret_cast<long long*>(state), indexes, n,
                                                         cdf, width, cdf_sizes, offsets, out_symbols);
  return dcvic_launch_status();
}
