/* TEST INFRASTRUCTURE (oracle): CPU restatement of the range-ANS coder DC-VIC writes its bitstreams with.
 *
 * The coder is NOT in /root/reference: it is the pybind11 extension `compressai.ans` of the un-vendored dependency
 * compressai==1.2.4 (pyproject.toml:13, poetry.lock:312-313), whose sources are
 *   compressai/cpp_exts/rans/rans_interface.cpp  (BufferedRansEncoder / RansEncoder / RansDecoder, bypass coding)
 *   third_party/ryg_rans/rans64.h                (Fabian Giesen's public-domain rANS, 64-bit state, 32-bit words)
 * restated here from their published algorithm; parity is anchored on the reference's call sites:
 *   EntropyModel.compress / decompress  -> src/models/comp_model/hyperprior_dc_vic_model.py:308-328,378-387
 *   RansDecoder.set_stream / decode_stream -> src/models/subnet/context_model/minnen20_charm_context_model.py:175-202
 * PARITY UNPINNED at the CompressAI boundary (no wheel offline, the reference holds no bitstream fixture); pinned by
 * construction: encode -> decode round trips, ryg_rans invariants (state in [2^31, 2^63)), byte-for-byte equality of
 * the GPU coder with this file.
 *
 * Conventions (rans_interface.cpp): precision 16 bits; symbol value v = symbol - offset[idx]; values outside
 * [0, max_value) with max_value = cdf_size[idx] - 2 are coded as the sentinel max_value followed by a bypass code
 * (4-bit digits): raw = v < 0 ? -2v - 1 : 2 (v - max_value); n_bypass digits count, unary in base-15 chunks, then the
 * raw digits low to high.  The encoder pushes symbols in REVERSE, the stream is the 32-bit words in decode order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION 16
#define BYPASS_PRECISION 4
#define MAX_BYPASS_VAL ((1 << BYPASS_PRECISION) - 1)
#define RANS64_L (1ull << 31)

typedef struct { uint16_t start, range; uint8_t bypass; } sym_t;

static inline void enc_put(uint64_t* r, uint32_t** pptr, uint32_t start, uint32_t freq, uint32_t scale_bits) {
  uint64_t x = *r;
  const uint64_t x_max = ((RANS64_L >> scale_bits) << 32) * freq;
  if (x >= x_max) { *pptr -= 1; **pptr = (uint32_t)x; x >>= 32; }
  *r = ((x / freq) << scale_bits) + (x % freq) + start;
}
static inline void enc_put_bits(uint64_t* r, uint32_t** pptr, uint32_t val, uint32_t nbits) {
  uint64_t x = *r;
  const uint32_t freq = 1u << (16 - nbits);
  const uint64_t x_max = ((RANS64_L >> 16) << 32) * freq;
  if (x >= x_max) { *pptr -= 1; **pptr = (uint32_t)x; x >>= 32; }
  *r = (x << nbits) | val;
}

/* symbols[n], indexes[n]; cdf [rows][width] row-major; returns the number of 32-bit words written to out (the stream,
 * in decode order), or -1 if out_cap words are not enough.  RansEncoder.encode_with_indexes. */
long rans_oracle_encode(const int32_t* symbols, const int32_t* indexes, long n, const int32_t* cdf, int width,
                        const int32_t* cdf_sizes, const int32_t* offsets, uint32_t* out, long out_cap) {
  /* 1st pass (forward): the list of coder symbols, as BufferedRansEncoder::encode_with_indexes builds it */
  long cap = 2 * n + 16, ns = 0;
  sym_t* syms = (sym_t*)malloc(sizeof(sym_t) * (size_t)cap);
  for (long i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    const int32_t* row = cdf + (size_t)ci * width;
    const int32_t max_value = cdf_sizes[ci] - 2;
    int32_t value = symbols[i] - offsets[ci];
    uint32_t raw_val = 0;
    if (value < 0) { raw_val = (uint32_t)(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { raw_val = (uint32_t)(2 * (value - max_value)); value = max_value; }
    if (ns + 24 > cap) { cap *= 2; syms = (sym_t*)realloc(syms, sizeof(sym_t) * (size_t)cap); }
    syms[ns++] = (sym_t){(uint16_t)row[value], (uint16_t)(row[value + 1] - row[value]), 0};
    if (value == max_value) {
      int32_t n_bypass = 0;
      while ((raw_val >> (n_bypass * BYPASS_PRECISION)) != 0) ++n_bypass;
      int32_t val = n_bypass;
      while (val >= MAX_BYPASS_VAL) { syms[ns++] = (sym_t){MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, 1}; val -= MAX_BYPASS_VAL; }
      syms[ns++] = (sym_t){(uint16_t)val, (uint16_t)(val + 1), 1};
      for (int32_t j = 0; j < n_bypass; ++j) {
        const int32_t v = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL;
        syms[ns++] = (sym_t){(uint16_t)v, (uint16_t)(v + 1), 1};
      }
    }
  }
  /* 2nd pass (backward): BufferedRansEncoder::flush */
  const long buf_words = ns + 4;
  uint32_t* buf = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)buf_words);
  uint32_t* ptr = buf + buf_words;
  uint64_t rans = RANS64_L;
  for (long i = ns - 1; i >= 0; --i) {
    if (!syms[i].bypass) enc_put(&rans, &ptr, syms[i].start, syms[i].range, PRECISION);
    else enc_put_bits(&rans, &ptr, syms[i].start, BYPASS_PRECISION);
  }
  ptr -= 2;
  ptr[0] = (uint32_t)(rans >> 0);
  ptr[1] = (uint32_t)(rans >> 32);
  const long nwords = (buf + buf_words) - ptr;
  long ret = -1;
  if (nwords <= out_cap) { memcpy(out, ptr, sizeof(uint32_t) * (size_t)nwords); ret = nwords; }
  free(buf);
  free(syms);
  return ret;
}

/* Decoder with persistent state across calls (RansDecoder::set_stream + decode_stream):
 * st[0] = rANS state, st[1] = read position in words, st[2] = initialised flag. */
static inline uint32_t dec_get_bits(uint64_t* r, const uint32_t* words, uint64_t* pos, uint32_t n_bits) {
  uint64_t x = *r;
  const uint32_t val = (uint32_t)(x & ((1u << n_bits) - 1));
  x >>= n_bits;
  if (x < RANS64_L) { x = (x << 32) | words[*pos]; *pos += 1; }
  *r = x;
  return val;
}

void rans_oracle_decode(const uint32_t* words, long nwords, uint64_t* st, const int32_t* indexes, long n,
                        const int32_t* cdf, int width, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out) {
  (void)nwords;
  uint64_t x = st[0], pos = st[1];
  if (!st[2]) { x = (uint64_t)words[0] | ((uint64_t)words[1] << 32); pos = 2; st[2] = 1; }   /* Rans64DecInit */
  for (long i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    const int32_t* row = cdf + (size_t)ci * width;
    const int32_t max_value = cdf_sizes[ci] - 2;
    const uint32_t cum = (uint32_t)(x & ((1u << PRECISION) - 1));
    int32_t s = 0;                                  /* std::find_if(cdf[v] > cum) - 1 */
    while (s + 1 < cdf_sizes[ci] && (uint32_t)row[s + 1] <= cum) ++s;
    const uint32_t start = (uint32_t)row[s], freq = (uint32_t)(row[s + 1] - row[s]);
    x = (uint64_t)freq * (x >> PRECISION) + (x & ((1u << PRECISION) - 1)) - start;     /* Rans64DecAdvance */
    if (x < RANS64_L) { x = (x << 32) | words[pos]; pos += 1; }
    int32_t value = s;
    if (value == max_value) {
      int32_t val = (int32_t)dec_get_bits(&x, words, &pos, BYPASS_PRECISION);
      int32_t n_bypass = val;
      while (val == MAX_BYPASS_VAL) { val = (int32_t)dec_get_bits(&x, words, &pos, BYPASS_PRECISION); n_bypass += val; }
      int32_t raw_val = 0;
      for (int32_t j = 0; j < n_bypass; ++j) {
        val = (int32_t)dec_get_bits(&x, words, &pos, BYPASS_PRECISION);
        raw_val |= val << (j * BYPASS_PRECISION);
      }
      value = raw_val >> 1;
      if (raw_val & 1) value = -value - 1; else value += max_value;
    }
    out[i] = value + offsets[ci];
  }
  st[0] = x;
  st[1] = pos;
}
