// Single-pass wide VQ forward on sm_100a: tcgen05 candidate search, FP32 re-rank, codebook gather, straight-through
// value, loss partials and indices in ONE persistent kernel.  z is fetched from HBM once (the second read, for the
// straight-through value, follows ~one tile time later and is an L2 hit), z_q is written once, and neither the
// distance matrix nor any candidate list ever leaves the SM.
//
// Reference semantics: taming/modules/vqvae/quantize.py:271-312 (d = |z|^2 + |e|^2 - 2 z.e, argmin with the lowest
// index on ties, z_q = z + (e - z), loss from mean((e - z)^2)).  Same arithmetic contract as the two-kernel path
// (vq_tcgen05.cu + vq_finish_tma.cu): FP16 x FP16 -> FP32 scores s = z.e - |e|^2/2 on the tensor cores flag every
// code within a PROVEN margin (vq_margin) of the row maximum; the flagged codes are re-ranked in FP32 in the
// reference's operation order.
//
// One CTA pair (cta_group::2, UMMA 256 x 128 x 16) per 256-token tile, 128 tokens (4 groups of 32) per CTA, 32 warps:
//   warp 0      TMA: this CTA's half of every N-tile's FP16 codebook slab (two 3-D boxes of the chunk-major codebook)
//   warp 1      TMA: z chunks [64 ch x 32 tokens] FP32, straight from NCHW, into an 8-stage conversion ring
//   warp 2      TMA: finish ring - loads z groups [e_dim x 32 tokens] (L2 hits), stores z_q from the same stages
//   warp 3      MMA issuer (leader CTA; warp-uniform loop, one elected lane).  The A operand lives in TENSOR MEMORY
//               (tcgen05.mma with [a_tmem]): shared memory only carries the codebook ring, the two z rings and the
//               candidate bookkeeping, and the MMA's operand reads take a quarter of what the SS form takes.
//   warps 4-7   converters: lane = token.  FP32 chunk from the ring -> FP16 pairs -> tcgen05.st into the A operand
//               (double-buffered: 2 x 128 columns, so a tile is converted while its predecessor multiplies), |z|^2
//               and the rounding residual per token
//   warps 8-23  epilogue: four column quarters x four lane quarters of every 128-column accumulator (double-buffered:
//               2 x 128 columns): ONE 32-code chunk per warp and N-tile.  tcgen05.ld 32 scores per row - the
//               accumulator goes back to the MMA as soon as a warp's scores are in registers -, -|e|^2/2 added exactly
//               from a table in shared memory (tensor memory is full: no room for the constant operand of an extra
//               K-step), running maximum, one flag mask per chunk; flagged chunks go to a 4-entry list per (token,
//               quarter) in shared memory (entries that a later, larger maximum rules out are dropped on the fly).
//               At the end of a tile the four quarters of a row exchange their maxima and append their surviving
//               codes to the token's candidate array (atomic slot counter).  Sixteen warps at 64 registers instead
//               of eight at 88 with two chunks each: the epilogue's latency per N-tile was what paced the kernel.
//   warps 24-31 consumers (the finish): (group, token quad) units.  First candidates' codebook rows are requested
//               before the group's z has arrived; tokens with more than one candidate are re-ranked in FP32
//               ((|z|^2 + |e|^2) - 2 z.e, lowest index on ties); z + (e - z) overwrites z in the stage; loss partials.
//
// Compile-time switches (python -m dc_vic_b200.build --variant <name> -D...; a variant library is only ever loaded
// through DCVIC_B200_LIB).  Measurement aids that produce WRONG results, used for the what-if table in DESIGN 4.1:
//   DCVIC_FZ_EXP=3|4 (no re-rank | consumers idle), DCVIC_FZ_EPIFREE (no flag arithmetic), DCVIC_FZ_X bits 1|2|4|8
//   (no z_q stores | no second read of z | conversion loads from the L2 | refill a finish stage without waiting for
//   its read-out), DCVIC_FZ_HALFB (half the codebook bytes), DCVIC_FZ_NOLD (no accumulator read-out).
// Correct variants that measured slower and are off: DCVIC_FZ_WHOLE_TILES, DCVIC_FZ_HELPERS, DCVIC_FZ_SUB64, DCVIC_FZ_CLUSTER=4
//   (+ DCVIC_FZ_MC_DIRECT=1), DCVIC_FZ_HINTS=0, DCVIC_FZ_NO_PREFETCH; ring depths DCVIC_FZ_NB / NZ / NF and the
//   register split FZ_REGS_CONV / FZ_REGS_CONS are tunables.  -DDCVIC_FZ_DEBUG (+ _TIMING_ONLY / _MARKS_ONLY /
//   DCVIC_FZ_NO_CMARKS) builds the instrumented kernels of tools/debug_fused.py and tools/fused_span.py.
#include <cuda.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "vq_common.cuh"
#include "sm100_ptx.cuh"

namespace dcvic {
namespace fz {

using namespace tc;

#if defined(DCVIC_FZ_SUB64) && DCVIC_FZ_SUB64
#define DCVIC_FZ_SUB64_EARLY 1
#else
#define DCVIC_FZ_SUB64_EARLY 0
#endif
constexpr int BM = 128;   // tokens per CTA tile
constexpr int BN = 128;   // codes per N-tile (UMMA N); 64 per CTA of the pair
constexpr int BK = 64;    // channels per chunk (64 fp16 = one SWIZZLE_128B row)
constexpr int GT = 32;    // tokens per group (one TMA box column block, one finish stage)
constexpr int NG = BM / GT;
#ifndef DCVIC_FZ_NB
#define DCVIC_FZ_NB 4
#endif
#ifndef DCVIC_FZ_NZ
#define DCVIC_FZ_NZ 8
#endif
#ifndef DCVIC_FZ_NF
#define DCVIC_FZ_NF 2
#endif
constexpr int NB = DCVIC_FZ_NB;     // codebook ring stages: half an N-tile slab each (see b_chunks_a)
constexpr int NZ = DCVIC_FZ_NZ;     // conversion ring stages (8 KB each), a multiple of 4: every stage always serves the
                                    // same converter warp, which therefore meets its barrier phases in order
constexpr int NF = DCVIC_FZ_NF;     // finish ring stages (e_dim * 128 B each)
constexpr int B_CHUNK = (BN / 2) * BK * 2;   // 8 KB: this CTA's 64 codes x 64 channels
constexpr int BP_BYTES = (BN / 2) * 16;      // 1 KB: 64 codes x 8 FP16 (the three pieces of -|e|^2/2 and five zeros)
constexpr int AP_BYTES = 256;                // two 8-row core matrices: [1, 1, 1, 0 x 5] per row, and zeros
constexpr int Z_STAGE = BK * GT * 4;         // 8 KB
constexpr int NT_MAX = 8;          // N-tiles per tile: one flag-mask slot per (token, column quarter, N-tile)
#if DCVIC_FZ_SUB64_EARLY   // (the four extra barriers of that experiment need 32 bytes)
constexpr int CK_MAX = 15;
#else
constexpr int CK_MAX = 16;         // candidate codes per token after compaction; more -> whole-codebook scan
#endif
constexpr int MAX_K = NT_MAX * BN;  // -|e|^2/2 table and flag-mask slots in shared memory (larger codebooks take the
                                    // two-kernel path)
// Cluster size: 2 = one CTA pair per cluster.  4 (experiment, -DDCVIC_FZ_CLUSTER=4) = two pairs that walk the codebook
// in lockstep: pair 0's CTAs load every codebook slab once and TMA-multicast it to the CTA of the same parity in pair 1
// (half the L2 -> SM codebook traffic: 156 of the 450 MB a launch moves through the L2).
#ifndef DCVIC_FZ_CLUSTER
#define DCVIC_FZ_CLUSTER 2
#endif
// Accumulator pipeline: 0 = two 128-column buffers, one per N-tile; 1 (experiment) = FOUR 64-column buffers: every
// 128-code slab of the ring is multiplied as two 64-code MMAs (rows 0-31 / 32-63 of either CTA's half), so the issuer
// can run three sub-tiles ahead of the slowest epilogue warp instead of one N-tile
#ifndef DCVIC_FZ_SUB64
#define DCVIC_FZ_SUB64 0
#endif
constexpr bool kSub64 = DCVIC_FZ_SUB64 != 0;
constexpr int NACC = kSub64 ? 4 : 2;                    // accumulator buffers
constexpr int CL = DCVIC_FZ_CLUSTER;
// (cluster of 4) how a multicast load reports to the MMA issuer: 0 = it completes on the barrier of the CTA it lands
// in and a non-leader forwards that to its leader (plain .multicast::cluster); 1 = the .cta_group::2 form, whose
// barrier operand names the pair's leader - one hop less
#ifndef DCVIC_FZ_MC_DIRECT
#define DCVIC_FZ_MC_DIRECT 0
#endif
constexpr bool kMcDirect = DCVIC_FZ_MC_DIRECT != 0;
static_assert(CL == 2 || CL == 4, "cluster of one or two CTA pairs");
constexpr int NCONS = 8;
constexpr int NEPI = 16;           // epilogue warps: 4 column quarters x 4 TMEM lane quarters
constexpr int NFIN = NCONS + NEPI; // warps that may run the finish (loss partial slots per CTA)
constexpr int kFullFlag = 1 << 16; // added to a token's candidate count: whole-codebook scan
static_assert(NZ % NG == 0, "a conversion-ring stage must always belong to the same converter warp");
constexpr int W_TMAB = 0, W_ZLOAD = 1, W_FIN = 2, W_MMA = 3, W_CONV0 = 4, W_EPI0 = 8, W_CONS0 = W_EPI0 + NEPI;
constexpr int NWARPS = W_CONS0 + NCONS;
constexpr int NTHREADS = NWARPS * 32;
// registers: 32 warps launch with 64 each = the whole file, and setmaxnreg can only move registers WITHIN the launch
// allocation (a total above it leaves the last warpgroup waiting for ever): TMA / MMA warps drop to 24, converters
// take 56 (at 40 their loop spilled and recomputed its swizzled addresses), the epilogue warps keep their 64,
// consumers take 88: 128 * (24 + 56 + 4 * 64 + 2 * 88) = 65,536
#define FZ_REGS_AUX 24
#ifndef FZ_REGS_CONV
#define FZ_REGS_CONV 56
#endif
#ifndef FZ_REGS_CONS
#define FZ_REGS_CONS 88
#endif
static_assert(128 * (FZ_REGS_AUX + FZ_REGS_CONV + 4 * 64 + 2 * FZ_REGS_CONS) <= 65536, "register file");
static_assert(NTHREADS == 1024, "register budget above assumes 32 warps");


struct Smem {
  // dynamic shared memory map (base aligned to 1024 B); the finish stages are sized for e_dim = 256
  static constexpr int OFF_Z = 0;
  static constexpr int OFF_F = OFF_Z + NZ * Z_STAGE;
  __host__ __device__ static constexpr int off_b(int D) { return OFF_F + NF * D * 128; }
  // an N-tile's codebook slab = KC channel chunks, brought by two 3-D boxes: chunks [0, nA) and [nA, KC); a ring
  // stage holds the larger one
  __host__ __device__ static constexpr int b_chunks_a(int D) { return (D / BK + 1) / 2; }
  // ... followed by BP_BYTES for this CTA's 64 codes of the N-tile's -|e|^2/2 operand (see the MMA issuer)
  __host__ __device__ static constexpr int b_slab(int D) { return b_chunks_a(D) * B_CHUNK; }
  __host__ __device__ static constexpr int b_stage(int D) { return b_slab(D) + BP_BYTES; }
  __host__ __device__ static constexpr int off_ap(int D) { return off_b(D) + NB * b_stage(D); }            // the ones operand
  __host__ __device__ static constexpr int off_list(int D) { return off_ap(D) + AP_BYTES; }                // [4][NT_MAX][BM] u32
  __host__ __device__ static constexpr int off_ck(int D) { return off_list(D) + BM * 4 * NT_MAX * 4; }      // [2][BM][CK_MAX] u16
  __host__ __device__ static constexpr int off_nc(int D) { return off_ck(D) + 2 * BM * CK_MAX * 2; }        // [2][BM] int
  __host__ __device__ static constexpr int off_zz(int D) { return off_nc(D) + 2 * BM * 4; }                 // [2][BM] float
  __host__ __device__ static constexpr int off_dz(int D) { return off_zz(D) + 2 * BM * 4; }                 // [2][BM] float
  __host__ __device__ static constexpr int off_m(int D) { return off_dz(D) + 2 * BM * 4; }                  // [BM][4] float
  __host__ __device__ static constexpr int off_bar(int D) { return off_m(D) + BM * 4 * 4; }
  // barrier slots (8 bytes each)
  static constexpr int BAR_B_FULL = 0;                      // [NB] leader only
  static constexpr int BAR_B_EMPTY = BAR_B_FULL + NB;       // [NB]
  static constexpr int BAR_Z_FULL = BAR_B_EMPTY + NB;       // [NZ]
  static constexpr int BAR_Z_EMPTY = BAR_Z_FULL + NZ;       // [NZ]
  static constexpr int BAR_A_FULL = BAR_Z_EMPTY + NZ;       // [2][4] leader only: chunk kc of A buffer b is written
  static constexpr int BAR_A_EMPTY = BAR_A_FULL + 8;        // [2] the tile's MMAs have read A buffer b
  static constexpr int BAR_T_FULL = BAR_A_EMPTY + 2;        // [NACC]
  static constexpr int BAR_T_EMPTY = BAR_T_FULL + NACC;     // [NACC] leader only
  static constexpr int BAR_ZZ = BAR_T_EMPTY + NACC;         // [2]
  static constexpr int BAR_C_FULL = BAR_ZZ + 2;             // [2]
  static constexpr int BAR_C_EMPTY = BAR_C_FULL + 2;        // [2]
  static constexpr int BAR_F_FULL = BAR_C_EMPTY + 2;        // [NF]
  static constexpr int BAR_F_DONE = BAR_F_FULL + NF;        // [NF]
  static constexpr int BAR_COUNT = BAR_F_DONE + NF;
  __host__ __device__ static constexpr int off_tmem(int D) { return off_bar(D) + BAR_COUNT * 8; }   // [0] TMEM base, [1] tiles converted x 4
  __host__ __device__ static constexpr int bytes(int D) { return off_tmem(D) + 16; }
};
static_assert(Smem::bytes(256) + 1024 <= 232448, "shared memory budget");

__device__ __forceinline__ void tma_load_2d_cta(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
// L2 eviction-priority hints (createpolicy): z is read twice by this kernel (conversion, then the finish about one
// tile time later) and z_q is never read again.  Bits of DCVIC_FZ_HINTS: 1 conversion loads evict_last, 2 z_q stores
// evict_first, 4 finish loads evict_first.  All three: 1.5-2 us per launch (0 = none, for A/B runs).
#ifndef DCVIC_FZ_HINTS
#define DCVIC_FZ_HINTS 7
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_cta_hint(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar,
                                                     uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
      "[%4], %5;" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, int x, int y, uint32_t src, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(map),
               "r"(x), "r"(y), "r"(src), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x),
               "r"(y), "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Bookkeeping in shared memory goes through shared-space instructions on 32-bit addresses: pointers derived from the
// aligned dynamic-smem base are GENERIC to the compiler (LD.E / ST.E / ATOM.E.GPU on every access otherwise).
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Progress marks for tools/debug_fused.py (-DDCVIC_FZ_DEBUG builds only): every warp's lane 0 writes
// (code << 16 | value) to a mapped host buffer, which the host can read while the kernel is still running (or hung).
#ifdef DCVIC_FZ_DEBUG
__device__ volatile int* g_fz_dbg = nullptr;
constexpr int FZ_REC = 48;          // ints per warp: [0] progress, [1..6] cycle slots, [7] total, [8..47] time marks
#define FZ_SLOT(k) g_fz_dbg[(blockIdx.x * 32 + (threadIdx.x >> 5)) * FZ_REC + (k)]
#if defined(DCVIC_FZ_TIMING_ONLY) || defined(DCVIC_FZ_MARKS_ONLY)
#define FZ_DBG(code, val)           // (a posted host write per step perturbs the timing)
#else
#define FZ_DBG(code, val)                                                                      \
  do {                                                                                         \
    if (g_fz_dbg && (threadIdx.x & 31) == 0) FZ_SLOT(0) = ((code) << 20) | ((val) & 0xFFFFF);  \
  } while (0)
#endif
// time marks: the SM clock (low 32 bits) when a warp passes a point; tools/debug_fused.py draws the tile timeline
#ifdef DCVIC_FZ_NO_CMARKS   // (tools/fused_span.py: wall-clock marks only)
#define FZ_MARK(k)
#else
#define FZ_MARK(k)                                                                             \
  do {                                                                                         \
    if (g_fz_dbg && (threadIdx.x & 31) == 0 && (k) < 40) FZ_SLOT(8 + (k)) = (int)clock64();    \
  } while (0)
#endif
// wall-clock marks (ns, low 32 bits of %globaltimer): comparable ACROSS SMs, unlike the SM clock
#define FZ_GMARK(k)                                                                            \
  do {                                                                                         \
    if (g_fz_dbg && (threadIdx.x & 31) == 0) {                                                 \
      unsigned long long t_;                                                                   \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                   \
      FZ_SLOT(8 + (k)) = (int)t_;                                                              \
    }                                                                                          \
  } while (0)
#ifdef DCVIC_FZ_MARKS_ONLY
#define FZ_TDECL
#define FZ_T()
#define FZ_ACC(k)
#define FZ_PUT()
#else
// cycle accounting per warp: FZ_T() marks "now"; FZ_ACC(k) adds the cycles since the last mark to slot k (1..6);
// FZ_PUT() writes the slots (and the warp's total in slot 7)
#define FZ_TDECL long long fz_t = clock64(), fz_t0 = fz_t, fz_acc[7] = {0, 0, 0, 0, 0, 0, 0}
#define FZ_T() fz_t = clock64()
#define FZ_ACC(k)                       \
  do {                                  \
    const long long now_ = clock64();   \
    fz_acc[k] += now_ - fz_t;           \
    fz_t = now_;                        \
  } while (0)
#define FZ_PUT()                                                            \
  do {                                                                      \
    if (g_fz_dbg && (threadIdx.x & 31) == 0) {                              \
      fz_acc[0] = clock64() - fz_t0;                                        \
      for (int k_ = 1; k_ < 7; ++k_) FZ_SLOT(k_) = (int)(fz_acc[k_] >> 3);  \
      FZ_SLOT(7) = (int)(fz_acc[0] >> 3);                                   \
    }                                                                       \
  } while (0)
#endif
#else
#define FZ_DBG(code, val)
#define FZ_MARK(k)
#define FZ_GMARK(k)
#define FZ_TDECL
#define FZ_T()
#define FZ_ACC(k)
#define FZ_PUT()
#endif

// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, N = 128, M = 256 (cta_group::2)
constexpr uint32_t kIdesc =
    (1u << 4) | ((uint32_t)((kSub64 ? BN / 2 : BN) >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
// Shared-memory descriptor of an un-swizzled K-major operand: core matrices of 8 rows x 16 bytes (128 contiguous
// bytes), `lbo` bytes between the two core matrices of a K = 16 step, `sbo` bytes between 8-row groups
// (validated, strides of 0 included, by tools/probe_ts_mma.cu).
__device__ __forceinline__ uint64_t umma_desc_none(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}

template <int D>
__global__ void __launch_bounds__(NTHREADS, 1)
vq_fused_kernel(const __grid_constant__ CUtensorMap tm_cb, const __grid_constant__ CUtensorMap tm_cb2,
                const __grid_constant__ CUtensorMap tm_bp, const __grid_constant__ CUtensorMap tm_zc,
                const __grid_constant__ CUtensorMap tm_zf, const __grid_constant__ CUtensorMap tm_zq,
                const float* __restrict__ Ep, const float* __restrict__ ee, const float* __restrict__ emax_ptr, int N,
                int HW, int K, int num_gp, int wait_first,
                float beta, int legacy, int64_t* __restrict__ idx, float* __restrict__ loss,
                double* __restrict__ partials, unsigned* __restrict__ counters) {
  constexpr int KC = D / BK;                 // channel chunks per tile
  constexpr int F_STAGE = D * 128;           // finish stage: [D channels][32 tokens] FP32
  constexpr int NA = Smem::b_chunks_a(D);    // chunks in the first box of an N-tile's slab, KC - NA in the second
  constexpr int B_STAGE = Smem::b_stage(D);
  constexpr int B_SLAB = Smem::b_slab(D);
  constexpr int OFF_B = Smem::off_b(D);
  constexpr int NH = (D + 127) / 128;        // 128-channel blocks (consumer lanes hold 4 channels of each)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + Smem::off_bar(D);
  auto bar = [&](int slot) { return bar0 + slot * 8; };
  // shared-space addresses of the bookkeeping arrays (see lds_u32 above)
  const uint32_t a_ap = sbase + Smem::off_ap(D);       // ones operand of the -|e|^2/2 K-step
  const uint32_t a_mask = sbase + Smem::off_list(D);   // [4 quarters][NT_MAX][BM] u32 flag masks
  const uint32_t a_ck = sbase + Smem::off_ck(D);       // [2][BM][CK_MAX] u16 candidate codes
  const uint32_t a_nc = sbase + Smem::off_nc(D);       // [2][BM] candidate counters
  const uint32_t a_zz = sbase + Smem::off_zz(D);       // [2][BM] |z|^2
  const uint32_t a_dz = sbase + Smem::off_dz(D);       // [2][BM] |z - fp16(z)|^2
  const uint32_t a_m = sbase + Smem::off_m(D);         // [BM][4] quarter maxima
  const uint32_t a_tmem = sbase + Smem::off_tmem(D);   // [0] TMEM base, [1] tiles converted x 4, [2] next consumer unit, [3] finish groups requested

  __shared__ int s_last;
  __shared__ double s_scratch[32];                   // [0, NFIN): loss partial per finishing warp; reused by the final sum
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 32) s_scratch[threadIdx.x] = 0.0;
  FZ_GMARK(33);                                      // kernel entry
  const uint32_t crank = cluster_ctarank();          // 0 .. CL - 1
  const uint32_t rank = crank & 1u;                  // inside the CTA pair: 0 = leader (issues the MMAs)
  const uint32_t pbase = crank & ~1u;                // cluster rank of this pair's leader
  const bool leader = rank == 0;
  const int cl = blockIdx.x / CL, ncl = gridDim.x / CL, pin = (int)(crank >> 1);   // cluster, pair inside it
  const int NT = K / BN;
  // Work split: the tokens go in "group pairs" of 64 (a group of 32 for either CTA), and CTA pair p takes the
  // contiguous run of group pairs [p * num_gp / npairs, (p + 1) * num_gp / npairs): tiles of 4 group pairs, of which
  // only the LAST may be partial.  Every pair gets the same work to within one group (whole 256-token tiles dealt
  // round-robin left 34 of 74 pairs with a fourth tile on the headline shape and the other 40 idle meanwhile); in a
  // partial tile the warps of the missing groups only keep the barriers going.
  // (Cluster of two pairs: the cluster's run is halved between its pairs, and BOTH walk as many tiles as the larger
  // half needs - they share the codebook ring's phases; a pair's surplus tile has no groups.)
  const int c_lo = (int)((long long)cl * num_gp / ncl), c_n = (int)((long long)(cl + 1) * num_gp / ncl) - c_lo;
  constexpr int PP = CL / 2;
  const int gp_lo = c_lo + pin * c_n / PP, gp_hi = c_lo + (pin + 1) * c_n / PP;
  const int my_groups = gp_hi - gp_lo;               // groups of 32 tokens this CTA handles
  const int my_tiles = ((c_n + PP - 1) / PP + NG - 1) / NG;
  auto tile_groups = [&](int it) { return max(0, min(NG, my_groups - it * NG)); };
  // first token of this CTA's group g of its it-th tile
  // (< N + 64: 32-bit unsigned arithmetic throughout - the TMA / MMA warps live on 24 registers)
  auto group_token0 = [&](int it, int g) { return ((uint32_t)(gp_lo + it * NG + g) * 2u + rank) * (uint32_t)GT; };
  auto leader_bar = [&](int slot) { return map_to_cta(bar(slot), pbase); };

  if (threadIdx.x == 0) {
    sts_u32(a_tmem + 4, 0u);
    sts_u32(a_tmem + 8, 0u);
    sts_u32(a_tmem + 12, 0u);
    // (cluster of two pairs: every CTA's B_FULL counts its own half's bytes, a leader's also its peer's forwarded
    // arrival; B_EMPTY counts the MMA commits of both pairs)
    for (int s = 0; s < NB; ++s) {
      mbar_init(bar(Smem::BAR_B_FULL + s), CL == 4 && leader && !kMcDirect ? 2 : 1);
      mbar_init(bar(Smem::BAR_B_EMPTY + s), CL / 2);
    }
    for (int s = 0; s < NZ; ++s) { mbar_init(bar(Smem::BAR_Z_FULL + s), 1); mbar_init(bar(Smem::BAR_Z_EMPTY + s), 1); }
    for (int c = 0; c < 8; ++c) mbar_init(bar(Smem::BAR_A_FULL + c), 2 * NG);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(Smem::BAR_A_EMPTY + b), 1);
      mbar_init(bar(Smem::BAR_T_FULL + b), 1);
      mbar_init(bar(Smem::BAR_T_EMPTY + b), 2 * NEPI);
      if (kSub64) {
        mbar_init(bar(Smem::BAR_T_FULL + 2 + b), 1);
        mbar_init(bar(Smem::BAR_T_EMPTY + 2 + b), 2 * NEPI);
      }
      mbar_init(bar(Smem::BAR_ZZ + b), NG);
      mbar_init(bar(Smem::BAR_C_FULL + b), NEPI);
      mbar_init(bar(Smem::BAR_C_EMPTY + b), NG * 8);          // one arrival per (group, quad) unit
    }
    for (int s = 0; s < NF; ++s) { mbar_init(bar(Smem::BAR_F_FULL + s), 1); mbar_init(bar(Smem::BAR_F_DONE + s), 8); }
    fence_barrier_init();
  }
  // ones operand: core matrix 0 = 8 rows of FP16 [1, 1, 1, 0, 0, 0, 0, 0], core matrix 1 = zeros
  if (threadIdx.x < AP_BYTES / 4) {
    const int w = threadIdx.x & 3;                   // 32-bit word of a 16-byte row
    sts_u32(a_ap + threadIdx.x * 4, threadIdx.x >= 32 ? 0u : w == 0 ? 0x3C003C00u : w == 1 ? 0x00003C00u : 0u);
  }
  fence_proxy_async();                               // the MMA reads it through the async proxy
  if (threadIdx.x < 2 * BM) sts_u32(a_nc + threadIdx.x * 4, 0u);     // candidate counters (atomic appends; the consumers re-zero them)
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + Smem::off_tmem(D)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  FZ_MARK(0);
  FZ_GMARK(34);                                      // set-up done
  const uint32_t tmem_base = lds_u32(a_tmem);
  // tensor memory: accumulators 2 x 128 columns | A operand 2 x 128 columns
  const uint32_t tmem_acc = tmem_base, tmem_a = tmem_base + 2 * BN;
  // PDL: see vq_tcgen05.cu.  wait_first: the predecessor in the stream may be the producer of z.
#ifndef DCVIC_FZ_NO_PREFETCH
  // While the predecessor in the stream finishes (it may be writing z), ask the L2 for this CTA's first tile of z: a
  // prefetch only moves lines into the L2, which stays coherent with whatever the predecessor still writes, and the
  // conversion ring's first loads then hit the L2 instead of paying an HBM round trip with the tensor cores idle.
  if (wait_first && warp == W_ZLOAD && lane == 0 && my_tiles > 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_zf) : "memory");
    const int ng = tile_groups(0);
    for (int g = 0; g < ng; ++g) {
      const uint32_t tg = group_token0(0, g);
      asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tm_zf),
                   "r"((int)(tg % (uint32_t)HW)), "r"((int)(tg / (uint32_t)HW) * D)
                   : "memory");
    }
  }
#endif
  if (wait_first) pdl_wait();
  pdl_launch_dependents();
  FZ_GMARK(35);                                      // predecessor complete

  // ------------------------------------------------------------------------------------------------------------------
  // The finish of (group, token quad) units drawn from a shared counter: FP32 re-rank of the tokens with more than one
  // candidate, z + (e - z) in place in the finish stage, loss partial, indices.  The consumer warps run it with TOK = 4
  // (the quad in one pass, 88 registers); the epilogue warps, once their own work is done, with TOK = 2 (two passes of
  // two tokens each fit the 64 registers they have).
  // Lane l holds 4 consecutive channels per 128-channel block (c = 4l + 128h) of the pass's tokens: one LDG.128 per
  // codebook row and block, one LDS / STS per channel.  Lane l visits its 4 channels in the order (j + (l >> 1)) & 3 so
  // that a quarter warp touches 8 different 16-byte pieces of the swizzled stage; the codebook copy it gathers from
  // (Ep, written by the prepare kernel) has every group of four channels in exactly that order, so a row's float4
  // pairs with the stage's values component by component.
  auto consume = [&](auto tok_c, const int slot) {
    constexpr int TOK = decltype(tok_c)::value;
    constexpr int NPASS = 4 / TOK;
    const int rot = (lane >> 1) & 3;
    const bool hv[2] = {4 * lane < D, 4 * lane + 128 < D};
    double dsq = 0.0;
    unsigned n_rr = 0, n_fs = 0;
    auto load_row = [&](float4 (&r)[NH], int k) {
      const float* rowp = Ep + (size_t)k * D + 4 * lane;
#pragma unroll
      for (int h = 0; h < NH; ++h)
        r[h] = hv[h] ? __ldg(reinterpret_cast<const float4*>(rowp + 128 * h)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto comp = [](const float4& v, int jj) { return jj == 0 ? v.x : jj == 1 ? v.y : jj == 2 ? v.z : v.w; };
    // the pass's TOK tokens of one channel: the quad's 16-byte piece, or one half of it
    auto ldz = [&](uint32_t a, float (&v)[TOK]) {
      if constexpr (TOK == 4) {
        const float4 t = lds128(a);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a));
      }
    };
    auto stz = [&](uint32_t a, const float (&v)[TOK]) {
      if constexpr (TOK == 4) {
        sts128(a, make_float4(v[0], v[1], v[2], v[3]));
      } else {
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v[0]), "f"(v[1]) : "memory");
      }
    };
    // this lane's four channel rows of a finish stage (stage base and the quad's 16-byte piece added per unit)
    uint32_t zrow[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) zrow[jj] = (4 * lane + ((jj + rot) & 3)) * 128;
    FZ_TDECL;
#ifdef DCVIC_FZ_DEBUG
    int fz_last_it = -1;
#endif
    // (group, token quad) units are handed out in order through a shared counter: a warp that drew long re-ranks does
    // not hold up its CTA (a fixed quad per warp cost 25 us on inputs where every fifth token is re-ranked)
    const int total_units = my_groups * 8;
    for (;;) {
      int u = 0;
      if (lane == 0) u = (int)atoms_add(a_tmem + 8, 1u);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= total_units) break;
      const int j = u >> 3, cw = u & 7;            // group (in this CTA's sequence), token quad inside it
      const int it = j / NG, g = j % NG, par = it & 1, st = j % NF;
      FZ_DBG(15, j);
      FZ_T();
      mbar_wait(bar(Smem::BAR_C_FULL + par), (it >> 1) & 1);
      FZ_ACC(1);
#ifdef DCVIC_FZ_DEBUG
      if (it != fz_last_it) { FZ_MARK(1 + it * 4); fz_last_it = it; }
#endif
      uint32_t zoff[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) zoff[jj] = zrow[jj] + (((uint32_t)cw ^ ((zrow[jj] >> 7) & 7u)) << 4);
      const uint32_t tq = group_token0(it, g) + 4 * cw;
#if defined(DCVIC_FZ_EXP) && DCVIC_FZ_EXP >= 4
      const bool live = tq > 0x7fffffffu;
#else
      const bool live = tq < (uint32_t)N;          // (a quad is valid or invalid as a whole: N % 4 == 0)
#endif
      const uint32_t zb = sbase + Smem::OFF_F + st * F_STAGE;
      int kfin[4] = {0, 0, 0, 0};
#pragma unroll
      for (int sub = 0; sub < NPASS; ++sub) {
        const int rq = g * GT + 4 * cw + sub * TOK;     // first row (token of the CTA tile) of this pass
        const uint32_t zsub = (uint32_t)(sub * TOK * 4);
        int nc[TOK], bk[TOK];
#pragma unroll
        for (int i = 0; i < TOK; ++i) {
          nc[i] = (int)lds_u32(a_nc + (par * BM + rq + i) * 4);
          bk[i] = (int)lds_u16(a_ck + (par * BM + rq + i) * (CK_MAX * 2));
          if (nc[i] <= 0 || nc[i] > CK_MAX) {        // flagged for a whole-codebook scan, too many candidates, or none
            if (live && nc[i] == 0 && lane == 0) atomicAdd(counters + 8, 1u);
            nc[i] = -1;
            bk[i] = 0;
          }
        }
#if defined(DCVIC_FZ_EXP) && DCVIC_FZ_EXP >= 3
#pragma unroll
        for (int i = 0; i < TOK; ++i) { nc[i] = 1; bk[i] = (rq + i) & 1023; }
#endif
        __syncwarp();
        if (lane < TOK) sts_u32(a_nc + (par * BM + rq + lane) * 4, 0u);  // for the tile after next (ordered by the C_EMPTY arrival below)
        float4 er[TOK][NH];
        // Second candidates: the row of the pass's first re-ranked token is requested together with the first
        // candidates' rows (and the |e|^2 values with them); inside the re-rank loop every token requests its
        // successor's before it reduces its own - one L2 round trip per pass instead of one per re-ranked token.
        const uint32_t ck0 = a_ck + (par * BM + rq) * (CK_MAX * 2);
        int jr = -1;
#pragma unroll
        for (int i = TOK - 1; i >= 0; --i)
          if (nc[i] > 1) jr = i;
        float4 e2[NH];
        int k2 = 0;
        float ee1[TOK], ee2 = 0.f;
#pragma unroll
        for (int i = 0; i < TOK; ++i) ee1[i] = 0.f;
#pragma unroll
        for (int h = 0; h < NH; ++h) e2[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
#pragma unroll
          for (int i = 0; i < TOK; ++i) load_row(er[i], bk[i]);
          if (jr >= 0) {
            k2 = (int)lds_u16(ck0 + jr * (CK_MAX * 2) + 2);
            load_row(e2, k2);
            ee2 = __ldg(ee + k2);
#pragma unroll
            for (int i = 0; i < TOK; ++i) ee1[i] = __ldg(ee + bk[i]);
          }
        }
        if (sub == 0) {
          FZ_DBG(16, j);
          FZ_ACC(2);
          // (With the 16 helper warps drawing units as well, one of 24 warps can be TWO rounds of the finish ring ahead
          // of a straggler, where the barrier's parity repeats - it hung: first make sure this group's load has been
          // requested at all, a monotonic count, then wait for its phase.  Eight warps cannot get that far ahead.)
#ifdef DCVIC_FZ_HELPERS
          while (lds_u32(a_tmem + 12) <= (uint32_t)j) __nanosleep(32);
#endif
          mbar_wait(bar(Smem::BAR_F_FULL + st), (j / NF) & 1);
          FZ_ACC(3);
          FZ_DBG(17, j);
        }
        if (live) {
          bool any_rr = false;
#pragma unroll
          for (int i = 0; i < TOK; ++i) any_rr |= nc[i] != 1;
          if (any_rr) {                                  // warp-uniform
#pragma unroll
            for (int i = 0; i < TOK; ++i) {
              if (nc[i] == 1) continue;
              const uint32_t ck = ck0 + i * (CK_MAX * 2);
              // (e2, k2, ee2 hold this token's second candidate.)  The next re-ranked token's is requested now:
              int jn = -1;
              float4 e2n[NH];
              int k2n = 0;
              float ee2n = 0.f;
#pragma unroll
              for (int h = 0; h < NH; ++h) e2n[h] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (nc[i] > 1) {
#pragma unroll
                for (int t = TOK - 1; t > i; --t)
                  if (nc[t] > 1) jn = t;
                if (jn >= 0) {
                  k2n = (int)lds_u16(ck0 + jn * (CK_MAX * 2) + 2);
                  load_row(e2n, k2n);
                  ee2n = __ldg(ee + k2n);
                }
              }
              float4 zg[NH];                           // this token's z in visiting order
              float zz = 0.f;
#pragma unroll
              for (int h = 0; h < NH; ++h) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (hv[h]) {
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) {
                    float t[TOK];
                    ldz(zb + zoff[jj] + zsub + h * 16384, t);
                    v[jj] = t[i];
                  }
                }
                zg[h] = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) zz = __fadd_rn(zz, __fmul_rn(v[jj], v[jj]));
              }
              auto part = [&](const float4 (&r)[NH]) {     // this lane's share of z . e
                float dp = 0.f;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                  dp = fmaf(zg[h].x, r[h].x, dp); dp = fmaf(zg[h].y, r[h].y, dp);
                  dp = fmaf(zg[h].z, r[h].z, dp); dp = fmaf(zg[h].w, r[h].w, dp);
                }
                return dp;
              };
              auto dot = [&](const float4 (&r)[NH]) { return warp_sum(part(r)); };
              float bd = FLT_MAX;
              int kb = 0x7fffffff;
              if (nc[i] > 1) {
                ++n_rr;
                // |z|^2 and the first two candidates' products go through the butterfly together (three dependent
                // 5-step shuffle chains one after the other were most of a re-rank's latency)
                float pa = part(er[i]), pb = part(e2);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  zz += __shfl_xor_sync(0xffffffffu, zz, o);
                  pa += __shfl_xor_sync(0xffffffffu, pa, o);
                  pb += __shfl_xor_sync(0xffffffffu, pb, o);
                }
                bd = fmaf(-2.f, pa, __fadd_rn(zz, ee1[i]));
                kb = bk[i];
                {
                  const float d0 = fmaf(-2.f, pb, __fadd_rn(zz, ee2));
                  if (d0 < bd || (d0 == bd && k2 < kb)) {
                    bd = d0;
                    kb = k2;
#pragma unroll
                    for (int h = 0; h < NH; ++h) er[i][h] = e2[h];
                  }
                }
#pragma unroll 1
                for (int ci = 2; ci < nc[i]; ++ci) {
                  float4 e0[NH];
                  const int k0 = (int)lds_u16(ck + ci * 2);
                  load_row(e0, k0);
                  const float d0 = fmaf(-2.f, dot(e0), __fadd_rn(zz, __ldg(ee + k0)));
                  if (d0 < bd || (d0 == bd && k0 < kb)) {
                    bd = d0;
                    kb = k0;
#pragma unroll
                    for (int h = 0; h < NH; ++h) er[i][h] = e0[h];
                  }
                }
                if (jn >= 0) {
                  k2 = k2n;
                  ee2 = ee2n;
#pragma unroll
                  for (int h = 0; h < NH; ++h) e2[h] = e2n[h];
                }
              } else {
                // whole-codebook scan (FP16-unsafe input or more candidates than fit; rare): two rows in flight
                ++n_fs;
                zz = warp_sum(zz);
#pragma unroll 1
                for (int k = 0; k < K; k += 2) {
                  float4 e0[NH], e1[NH];
                  const int k1 = min(k + 1, K - 1);
                  load_row(e0, k);
                  load_row(e1, k1);
                  const float d0 = fmaf(-2.f, dot(e0), __fadd_rn(zz, __ldg(ee + k)));
                  const float d1 = fmaf(-2.f, dot(e1), __fadd_rn(zz, __ldg(ee + k1)));
                  if (d0 < bd || (d0 == bd && k < kb)) { bd = d0; kb = k; }
                  if (d1 < bd || (d1 == bd && k1 < kb)) { bd = d1; kb = k1; }
                }
                load_row(er[i], kb);
              }
              bk[i] = kb;
            }
          }
          if (sub == NPASS - 1) FZ_ACC(4);
          // ---- z_q = z + (e - z) in place, loss partial
          float sq = 0.f;
          float za[NH][4][TOK];                         // all loads first: one round trip to shared memory per pass
#pragma unroll
          for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              if (hv[h]) {
                ldz(zb + zoff[jj] + zsub + h * 16384, za[h][jj]);
              } else {
#pragma unroll
                for (int i = 0; i < TOK; ++i) za[h][jj][i] = 0.f;
              }
            }
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            if (!hv[h]) continue;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float o[TOK];
#pragma unroll
              for (int i = 0; i < TOK; ++i) {
                const float a = za[h][jj][i];
                const float d = __fsub_rn(comp(er[i][h], jj), a);
                o[i] = __fadd_rn(a, d);
                sq = fmaf(d, d, sq);
              }
              stz(zb + zoff[jj] + zsub + h * 16384, o);
            }
          }
          dsq += (double)sq;
        }
#pragma unroll
        for (int i = 0; i < TOK; ++i) kfin[sub * TOK + i] = bk[i];
      }
      fence_proxy_async();                          // generic-proxy writes of the stage -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(Smem::BAR_F_DONE + st));
        mbar_arrive(bar(Smem::BAR_C_EMPTY + par));
      }
      // (after the fence: the fence would otherwise wait for this global store as well)
      if (live && lane < 4)
        idx[tq + lane] = (int64_t)(lane == 0 ? kfin[0] : lane == 1 ? kfin[1] : lane == 2 ? kfin[2] : kfin[3]);
      FZ_ACC(5);
      if ((u & 31) >= 24) FZ_MARK(3 + it * 4);     // (the tile's last eight units; a mark per unit perturbs)
    }
    FZ_PUT();
    // loss: one partial per finishing warp, summed in index order by the last CTA (deterministic for a given
    // assignment of units to warps)
    {
      const double wsum = warp_sum(dsq);
      if (lane == 0) s_scratch[slot] = wsum;           // (summed per CTA at the exit, in slot order)
    }
    if (lane == 0) {
      if (n_rr) atomicAdd(counters + kCtrRerank, n_rr);
      if (n_fs) atomicAdd(counters + kCtrOverflow, n_fs);
    }
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FZ_REGS_AUX));
    if (warp == W_TMAB) {
      // ===================== codebook ring: this CTA's 64 codes of every [128 codes x 64 ch] chunk =====================
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_cb) : "memory");
        pdl_wait();                                 // the FP16 codebook is written by the prepare kernel
        int stage = 0;
        uint32_t phase = 0;
        FZ_TDECL;
        // two 3-D boxes per N-tile: [64 ch] x [this CTA's 64 codes] x [NA | KC - NA chunks] of the chunk-major FP16
        // codebook (a request costs the issuing thread ~235 cycles whatever its size - tools/probe_tma.cu -, so 8 KB
        // boxes could not feed an MMA that eats one every 218 cycles)
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_cb2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bp) : "memory");
        for (int it = 0; it < my_tiles; ++it)
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int part = 0; part < (KC > NA ? 2 : 1); ++part) {
              FZ_DBG(1, (it * NT + nt) * 2 + part);
              FZ_T();
              mbar_wait(bar(Smem::BAR_B_EMPTY + stage), phase ^ 1);
              FZ_ACC(1);
              const int nch = part == 0 ? NA : KC - NA;
#ifdef DCVIC_FZ_HALFB   // experiment: half the codebook traffic (the second box is not loaded; wrong results)
              if (part == 1) {
                if (leader) mbar_arrive(bar(Smem::BAR_B_FULL + stage));
                if (++stage == NB) { stage = 0; phase ^= 1; }
                continue;
              }
#endif
              if constexpr (CL == 4) {
                // Two pairs in lockstep: every CTA arms its OWN barrier for its half's bytes; pair 0's CTAs issue the
                // loads, multicast to themselves and to the CTA of the same parity in pair 1 (data and complete_tx land
                // at the same offsets in both).  A non-leader's warp 3 forwards its barrier's completion to its leader.
                if constexpr (kMcDirect) {
                  if (leader)
                    mbar_arrive_expect_tx(bar(Smem::BAR_B_FULL + stage), 2 * (nch * B_CHUNK + (part == 0 ? BP_BYTES : 0)));
                  if (pin == 0) {
                    const uint16_t mask = (uint16_t)((1u << crank) | (1u << (crank + 2)));
                    const uint32_t lb = leader_bar(Smem::BAR_B_FULL + stage);
                    if (part == 0)
                      asm volatile(
                          "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
                          ".multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(sbase + OFF_B + stage * B_STAGE + B_SLAB),
                          "l"(&tm_bp), "r"(0), "r"(nt * BN + (int)rank * (BN / 2)), "r"(lb), "h"(mask)
                          : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
                        ".multicast::cluster [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(sbase + OFF_B + stage * B_STAGE),
                        "l"(part == 0 ? &tm_cb : &tm_cb2), "r"(0), "r"(nt * BN + (int)rank * (BN / 2)),
                        "r"(part == 0 ? 0 : NA), "r"(lb), "h"(mask)
                        : "memory");
                  }
                  if (++stage == NB) { stage = 0; phase ^= 1; }
                  continue;
                }
                const uint32_t fb = bar(Smem::BAR_B_FULL + stage);
                mbar_arrive_expect_tx(fb, nch * B_CHUNK + (part == 0 ? BP_BYTES : 0));
                if (pin == 0) {
                  const uint16_t mask = (uint16_t)((1u << crank) | (1u << (crank + 2)));
                  if (part == 0)
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
                        "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(sbase + OFF_B + stage * B_STAGE + B_SLAB),
                        "l"(&tm_bp), "r"(0), "r"(nt * BN + (int)rank * (BN / 2)), "r"(fb), "h"(mask)
                        : "memory");
                  asm volatile(
                      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
                      "[%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(sbase + OFF_B + stage * B_STAGE),
                      "l"(part == 0 ? &tm_cb : &tm_cb2), "r"(0), "r"(nt * BN + (int)rank * (BN / 2)),
                      "r"(part == 0 ? 0 : NA), "r"(fb), "h"(mask)
                      : "memory");
                }
                if (++stage == NB) { stage = 0; phase ^= 1; }
                continue;
              }
              if (leader)
                mbar_arrive_expect_tx(bar(Smem::BAR_B_FULL + stage), 2 * (nch * B_CHUNK + (part == 0 ? BP_BYTES : 0)));
              if (part == 0)       // this CTA's 64 codes of the N-tile's -|e|^2/2 operand: 16-byte rows, un-swizzled
                asm volatile(
                    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
                    "{%2, %3}], [%4];" ::"r"(sbase + OFF_B + stage * B_STAGE + B_SLAB),
                    "l"(&tm_bp), "r"(0), "r"(nt * BN + (int)rank * (BN / 2)), "r"(leader_bar(Smem::BAR_B_FULL + stage))
                    : "memory");
              asm volatile(
                  "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
                  "{%2, %3, %4}], [%5];" ::"r"(sbase + OFF_B + stage * B_STAGE),
                  "l"(part == 0 ? &tm_cb : &tm_cb2), "r"(0), "r"(nt * BN + (int)rank * (BN / 2)), "r"(part == 0 ? 0 : NA),
                  "r"(leader_bar(Smem::BAR_B_FULL + stage))
                  : "memory");
              if (++stage == NB) { stage = 0; phase ^= 1; }
            }
        FZ_PUT();
      }
    } else if (warp == W_ZLOAD) {
      // ===================== conversion ring: z chunks [64 ch x 32 tokens], order (tile, chunk, group) ================
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_zc) : "memory");
        FZ_TDECL;
        for (int it = 0; it < my_tiles; ++it) {
          const int ng = tile_groups(it);
#pragma unroll 1
          for (int kc = 0; kc < KC; ++kc)
#pragma unroll 1
            for (int g = 0; g < ng; ++g) {
              const int s = (it * KC + kc) * NG + g;    // (the converters' numbering; a partial tile - the last - skips slots)
              const int st = s % NZ;
              FZ_DBG(2, s);
              FZ_T();
              mbar_wait(bar(Smem::BAR_Z_EMPTY + st), ((s / NZ) & 1) ^ 1);
              FZ_ACC(1);
#if defined(DCVIC_FZ_X) && (DCVIC_FZ_X & 4)   // experiment: conversion loads from L2 (one group, over and over)
              const uint32_t tg = rank * (uint32_t)GT;
#else
              const uint32_t tg = group_token0(it, g);   // (a group beyond N: out-of-bounds box, zero fill)
#endif
              mbar_arrive_expect_tx(bar(Smem::BAR_Z_FULL + st), Z_STAGE);
#if DCVIC_FZ_HINTS & 1
              tma_load_2d_cta_hint(sbase + Smem::OFF_Z + st * Z_STAGE, &tm_zc, (int)(tg % (uint32_t)HW),
                                   (int)(tg / (uint32_t)HW) * D + kc * BK, bar(Smem::BAR_Z_FULL + st),
                                   l2_policy_evict_last());
#else
              tma_load_2d_cta(sbase + Smem::OFF_Z + st * Z_STAGE, &tm_zc, (int)(tg % (uint32_t)HW),
                              (int)(tg / (uint32_t)HW) * D + kc * BK,
                              bar(Smem::BAR_Z_FULL + st));
#endif
            }
          FZ_MARK(3 + it * 4);
        }
        FZ_PUT();
      }
    } else if (warp == W_FIN) {
      // ===================== finish ring: load z groups (after their tile has been converted: L2 hits), store z_q ====
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_zf) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_zq) : "memory");
        const int total = my_groups;
        FZ_TDECL;
        auto coords = [&](int j, int& x, int& y) {
          const uint32_t tg = group_token0(j / NG, j % NG);
          x = (int)(tg % (uint32_t)HW);
          y = (int)(tg / (uint32_t)HW) * D;
        };
        auto load = [&](int j) {
          const int it = j / NG, st = j % NF;
          // tile `it` has been converted, i.e. its z is in L2 (a monotonic counter, not the ZZ barrier: this thread
          // may trail the converters by more than one phase of it; a heuristic for the cache, not a data dependence)
          FZ_DBG(3, j);
          FZ_T();
          while (lds_u32(a_tmem + 4) < (uint32_t)(NG * (it + 1))) __nanosleep(64);
          FZ_ACC(1);
          int x, y;
          coords(j, x, y);
#if defined(DCVIC_FZ_X) && (DCVIC_FZ_X & 2)   // experiment: no second read of z
          mbar_arrive(bar(Smem::BAR_F_FULL + st));
#else
          mbar_arrive_expect_tx(bar(Smem::BAR_F_FULL + st), F_STAGE);
#if DCVIC_FZ_HINTS & 4
          tma_load_2d_cta_hint(sbase + Smem::OFF_F + st * F_STAGE, &tm_zf, x, y, bar(Smem::BAR_F_FULL + st),
                               l2_policy_evict_first());
#else
          tma_load_2d_cta(sbase + Smem::OFF_F + st * F_STAGE, &tm_zf, x, y, bar(Smem::BAR_F_FULL + st));
#endif
#endif
          sts_u32(a_tmem + 12, (uint32_t)(j + 1));   // groups requested so far (see the finish's wait for its stage)
          if (j % NG == 0) FZ_MARK(1 + it * 4);
        };
        for (int j = 0; j < total && j < NF; ++j) load(j);
        for (int j = 0; j < total; ++j) {
          const int st = j % NF;
          FZ_DBG(4, j);
          FZ_T();
          mbar_wait(bar(Smem::BAR_F_DONE + st), (j / NF) & 1);
          FZ_ACC(2);
          int x, y;
          coords(j, x, y);
#if !defined(DCVIC_FZ_X) || !(DCVIC_FZ_X & 1)  // (experiment bit 1: no z_q stores)
#if DCVIC_FZ_HINTS & 2
          tma_store_2d_hint(&tm_zq, x, y, sbase + Smem::OFF_F + st * F_STAGE, l2_policy_evict_first());
#else
          tma_store_2d(&tm_zq, x, y, sbase + Smem::OFF_F + st * F_STAGE);
#endif
#endif
          bulk_commit();
          if (j % NG == 0) FZ_MARK(2 + (j / NG) * 4);
          if (j % NG == NG - 1) FZ_MARK(3 + (j / NG) * 4);
          if (j + NF < total) {
            FZ_T();
#if !defined(DCVIC_FZ_X) || !(DCVIC_FZ_X & 8)  // (experiment bit 8: refill without waiting for the read-out - racy)
            bulk_wait_read_all();                    // the stage has been read out: refill it
#endif
            FZ_ACC(3);
            load(j + NF);
          }
        }
        FZ_T();
        bulk_wait_all();
        FZ_ACC(4);
        FZ_PUT();
      }
    } else if (leader) {
      // ===================== MMA issuer: whole warp walks the loop, one elected lane issues =====================
      const bool issuer = elect_one();
      const uint16_t pair_mask = (uint16_t)(3u << pbase);   // this pair's two CTAs in the cluster
      const uint32_t rt_one = my_tiles > 0 ? 1u : 0u;   // 1, but not a compile-time constant (see vq_tcgen05.cu)
      const uint32_t rt_zero = my_tiles < 0 ? 1u : 0u;  // 0, likewise
      // The loop below is the kernel's pacemaker and shares its scheduler with seven busy warps: it is kept to the
      // fewest instructions per MMA (descriptors advance by constants, channel chunks unrolled, one divergent region
      // per codebook box).
      const uint64_t bd_ring = umma_desc_sw128(sbase + OFF_B);     // stage 0; a stage further: + B_STAGE / 16
      const uint64_t bp_ring = umma_desc_none(sbase + OFF_B + B_SLAB, 0, 128);
      const uint64_t ad_ones = umma_desc_none(a_ap, 128, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;
      FZ_TDECL;
      for (int it = 0; it < my_tiles; ++it) {
        const int abuf = it & 1;
        const uint32_t a0 = tmem_a + abuf * BM;           // 128 columns per A buffer
        for (int nt = 0; nt < NT; ++nt, ++g) {
          if constexpr (kSub64) {
            // Two 64-code MMAs per slab (rows 0-31, then 32-63, of either CTA's half) into four 64-column buffers:
            // accumulator column j of sub-tile `sub` is code nt * 128 + (j / 32) * 64 + sub * 32 + j % 32.
            constexpr int NPARTS = KC > NA ? 2 : 1;
            constexpr int c0s[2] = {0, NA}, c1s[2] = {NA, KC};
            int st[2];
            uint32_t ph[2];
#pragma unroll
            for (int part = 0; part < NPARTS; ++part) {
              st[part] = stage;
              ph[part] = phase;
              if (++stage == NB) { stage = 0; phase ^= 1; }
            }
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const uint32_t sidx = g * 2 + sub, buf = sidx & 3;
              FZ_DBG(5, sidx);
              FZ_T();
              if (sidx >= 4) mbar_wait(bar(Smem::BAR_T_EMPTY + buf), ((sidx >> 2) - 1) & 1);
              FZ_ACC(1);
              const uint32_t d = tmem_acc + buf * (BN / 2);
#pragma unroll
              for (int part = 0; part < NPARTS; ++part) {
                const int c0 = c0s[part], c1 = c1s[part];
                if (sub == 0) {
                  FZ_T();
                  mbar_wait(bar(Smem::BAR_B_FULL + st[part]), ph[part]);
                  FZ_ACC(3);
                  if (nt == 0) {
                    FZ_T();
#pragma unroll
                    for (int c = c0; c < c1; ++c) mbar_wait(bar(Smem::BAR_A_FULL + abuf * 4 + c), (it >> 1) & 1);
                    FZ_ACC(2);
                  }
                }
                tc_fence_after();
                if (issuer) {
                  const uint64_t bds = bd_ring + (uint64_t)(uint32_t)(st[part] * (B_STAGE >> 4) + sub * (4096 >> 4));
                  if (part == 0)
                    umma_ss(d, ad_ones, bp_ring + (uint64_t)(uint32_t)(st[0] * (B_STAGE >> 4) + sub * (512 >> 4)), rt_zero);
#pragma unroll
                  for (int c = c0; c < c1; ++c) {
                    const uint64_t bd = bds + (uint64_t)((c - c0) * (B_CHUNK >> 4));
                    const uint32_t a = a0 + c * (BK / 2);
                    umma_ts(d, a, bd, rt_one);
                    umma_ts(d, a + 8, bd + 2, rt_one);
                    umma_ts(d, a + 16, bd + 4, rt_one);
                    umma_ts(d, a + 24, bd + 6, rt_one);
                  }
                  if (sub == 1) umma_commit<2>(bar(Smem::BAR_B_EMPTY + st[part]), CL == 4 ? (uint16_t)0xF : pair_mask);
                  if (part == NPARTS - 1) umma_commit<2>(bar(Smem::BAR_T_FULL + buf), pair_mask);
                }
              }
            }
            continue;
          }
          const uint32_t buf = g & 1;
          FZ_DBG(5, g);
          FZ_T();
          if (g >= 2) mbar_wait(bar(Smem::BAR_T_EMPTY + buf), ((g >> 1) - 1) & 1);   // both CTAs hold its scores in registers
          FZ_ACC(1);
          const uint32_t d = tmem_acc + buf * BN;
#pragma unroll
          for (int part = 0; part < (KC > NA ? 2 : 1); ++part) {
            constexpr int c0s[2] = {0, NA}, c1s[2] = {NA, KC};
            const int c0 = c0s[part], c1 = c1s[part];
            FZ_T();
            mbar_wait(bar(Smem::BAR_B_FULL + stage), phase);
            FZ_ACC(3);
            if (nt == 0) {                                 // the tile's A operand arrives chunk by chunk
              FZ_T();
#pragma unroll
              for (int c = c0; c < c1; ++c) {
                mbar_wait(bar(Smem::BAR_A_FULL + abuf * 4 + c), (it >> 1) & 1);
                if (c == 0) FZ_MARK(1 + it * 4);
                if (c == KC - 1) FZ_MARK(2 + it * 4);
              }
              FZ_ACC(2);
            }
            tc_fence_after();
            if (issuer) {
              const uint64_t bds = bd_ring + (uint64_t)(uint32_t)(stage * (B_STAGE >> 4));
              // The N-tile's first MMA overwrites the buffer with -|e_k|^2/2 for every row: a K = 16 step in the SS
              // form, ones operand [1, 1, 1, 0 ...] x the three FP16 pieces of the constant (un-swizzled K-major
              // operands; a stride of 0 makes every 8-row group read the same core matrix, and the codes' second core
              // matrix - whatever it holds - meets the ones operand's zeros).  The epilogue pays nothing for it.
              if (part == 0) umma_ss(d, ad_ones, bp_ring + (uint64_t)(uint32_t)(stage * (B_STAGE >> 4)), rt_zero);
#pragma unroll
              for (int c = c0; c < c1; ++c) {
                const uint64_t bd = bds + (uint64_t)((c - c0) * (B_CHUNK >> 4));
                const uint32_t a = a0 + c * (BK / 2);
                umma_ts(d, a, bd, rt_one);
                umma_ts(d, a + 8, bd + 2, rt_one);
                umma_ts(d, a + 16, bd + 4, rt_one);
                umma_ts(d, a + 24, bd + 6, rt_one);
              }
              umma_commit<2>(bar(Smem::BAR_B_EMPTY + stage), CL == 4 ? (uint16_t)0xF : pair_mask);   // (every loader)
              if (part == (KC > NA ? 1 : 0)) umma_commit<2>(bar(Smem::BAR_T_FULL + buf), pair_mask);
            }
            if (++stage == NB) { stage = 0; phase ^= 1; }
          }
        }
        if (issuer) umma_commit<2>(bar(Smem::BAR_A_EMPTY + abuf), pair_mask);
        FZ_MARK(3 + it * 4);
        __syncwarp();
      }
      FZ_PUT();
    } else if (CL == 4 && !kMcDirect) {
      // ===================== non-leader, warp 3 (cluster of two pairs): forward "my half has landed" =====================
      // The multicast loads complete on the barrier of the CTA they land in; the MMA issuer waits on ITS barrier, which
      // counts its own half (bytes) and this arrival.
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const int n = my_tiles * NT * (KC > NA ? 2 : 1);
        for (int i = 0; i < n; ++i) {
          mbar_wait(bar(Smem::BAR_B_FULL + stage), phase);
          mbar_arrive_cluster_release(leader_bar(Smem::BAR_B_FULL + stage));
          if (++stage == NB) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < W_EPI0) {
    // ===================== converters: FP32 ring stage -> FP16 A operand in tensor memory =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FZ_REGS_CONV));
    const int g = warp - W_CONV0;                    // token group == TMEM lane quarter (warp % 4)
    const int row = g * GT + lane;
    const uint32_t tl = tmem_a + ((uint32_t)(g * 32) << 16);
    const uint32_t lane_off = ((uint32_t)(lane & 3)) << 2;
    const uint32_t lq = (uint32_t)(lane >> 2);
    FZ_TDECL;
    for (int it = 0; it < my_tiles; ++it) {
      const int abuf = it & 1;
      FZ_DBG(8, it);
      FZ_T();
      mbar_wait(bar(Smem::BAR_A_EMPTY + abuf), ((it >> 1) & 1) ^ 1);   // the MMAs of tile it-2 have read this buffer
      FZ_ACC(1);
      FZ_MARK(1 + it * 4);
      tc_fence_after();
      if (g >= tile_groups(it)) {                    // partial (last) tile: nothing to convert, the barriers still count us
        if (lane == 0) {
          for (int kc = 0; kc < KC; ++kc) mbar_arrive_cluster(leader_bar(Smem::BAR_A_FULL + abuf * 4 + kc));
          mbar_arrive(bar(Smem::BAR_ZZ + abuf));
          atoms_add(a_tmem + 4, 1u);
        }
        continue;
      }
      float zz = 0.f, dz2 = 0.f;
#pragma unroll 1
      for (int kc = 0; kc < KC; ++kc) {
        const int s = (it * KC + kc) * NG + g;
        const int st = s % NZ;
        FZ_DBG(9, s);
        FZ_T();
        mbar_wait(bar(Smem::BAR_Z_FULL + st), (s / NZ) & 1);
        FZ_ACC(2);
        if (kc == 0) FZ_MARK(2 + it * 4);
        const uint32_t zb = sbase + Smem::OFF_Z + st * Z_STAGE + lane_off;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                 // 32 channels -> 16 columns per store
          uint32_t r[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const int ch = h * 32 + c;
            const float v0 = lds32(zb + ch * 128 + ((lq ^ (uint32_t)(ch & 7)) << 4));
            const float v1 = lds32(zb + (ch + 1) * 128 + ((lq ^ (uint32_t)((ch + 1) & 7)) << 4));
            zz = fmaf(v0, v0, zz);
            zz = fmaf(v1, v1, zz);
            const uint32_t pk = pack_f16x2(v0, v1);
            r[c >> 1] = pk;
            // the rounding residual of this pair (exact in FP32): the margin uses its measured norm
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&pk));
            const float d0 = v0 - hf.x, d1 = v1 - hf.y;
            dz2 = fmaf(d0, d0, dz2);
            dz2 = fmaf(d1, d1, dz2);
          }
          TMEM_ST16(tl + abuf * BM + kc * (BK / 2) + h * 16, r);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Smem::BAR_Z_EMPTY + st));     // every lane's loads have been consumed
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_bar(Smem::BAR_A_FULL + abuf * 4 + kc));
      }
      sts_u32(a_zz + (abuf * BM + row) * 4, __float_as_uint(zz));
      sts_u32(a_dz + (abuf * BM + row) * 4, __float_as_uint(dz2));
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(Smem::BAR_ZZ + abuf));
        atoms_add(a_tmem + 4, 1u);
      }
      FZ_MARK(3 + it * 4);
    }
    FZ_PUT();
  } else if (warp < W_CONS0) {
    // ===================== epilogue: flag masks per 32 codes, running maximum per row, candidate lists ==========
    // (these warps keep the 64 registers of the launch)
    const int q = (warp - W_EPI0) >> 2;              // column quarter of every accumulator: one 32-code chunk per N-tile
    const int part = warp & 3;                       // TMEM lane quarter
    const int row = part * 32 + lane;
    const uint32_t tlane = tmem_acc + ((uint32_t)(part * 32) << 16) + q * kChunk;
    FZ_DBG(20, 0);
    pdl_wait();                                      // emax, -|e|^2/2 (prepare kernel)
    FZ_DBG(21, 0);
    const float emax = emax_ptr[0];
    const bool cb_unsafe = __float_as_uint(emax_ptr[1]) != 0u;
    const float demax = emax_ptr[2];
    // The flag mask of (row, quarter q, N-tile nt) lives at a_mask + ((q * NT_MAX + nt) * BM + row) * 4: ONE writer and
    // one reader, this thread; the 32 rows of a warp are 32 consecutive words.  `live` says which slots count.
    const uint32_t my_mask = a_mask + ((q * NT_MAX) * BM + row) * 4;
    uint32_t g = 0;
    FZ_TDECL;
    for (int it = 0; it < my_tiles; ++it) {
      const int abuf = it & 1;
      const bool have = part < tile_groups(it);       // (partial tile: this lane quarter's group may be missing)
      const bool valid = have && group_token0(it, part) + lane < (uint32_t)N;
      FZ_DBG(10, it);
      FZ_T();
      mbar_wait(bar(Smem::BAR_ZZ + abuf), (it >> 1) & 1);
      FZ_ACC(1);
      const float zz = __uint_as_float(lds_u32(a_zz + (abuf * BM + row) * 4));
      const float margin =
          vq_margin_measured(zz, __uint_as_float(lds_u32(a_dz + (abuf * BM + row) * 4)), emax, demax, D);
      float m = -INFINITY;                           // running maximum over this quarter's codes
      uint32_t live = 0u;                            // N-tiles whose flag mask may hold a candidate
      if constexpr (kSub64) {
        // 64-column buffers: this warp reads 16 columns of every sub-tile (q >> 1: which CTA's half of the slab, q & 1:
        // which 16 of that half's 32 codes); masks are 16 bits at a_mask + ((q * 16 + slot) * BM + row) * 2
        for (int sl = 0; sl < 2 * NT; ++sl) {
          const uint32_t sidx = g * 2 + (uint32_t)sl, buf = sidx & 3;
          FZ_T();
          mbar_wait(bar(Smem::BAR_T_FULL + buf), (sidx >> 2) & 1);
          FZ_ACC(2);
          if (!have) {
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_bar(Smem::BAR_T_EMPTY + buf));
            continue;
          }
          tc_fence_after();
          uint32_t rb[16];
          TMEM_LD16(rb, tmem_acc + ((uint32_t)(part * 32) << 16) + buf * (BN / 2) + q * 16);
          TMEM_WAIT_LD16(rb);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_bar(Smem::BAR_T_EMPTY + buf));
          FZ_ACC(3);
          float cm;
          const float m_old = m;
          const uint32_t mask = chunk_flags16(rb, margin, m, cm);
          if (mask != 0u) {
            if (cm > m_old + margin) live = 0u;
            sts_u16(a_mask + ((q * 16 + sl) * BM + row) * 2, mask);
            live |= 1u << sl;
          }
          FZ_ACC(5);
        }
        g += NT;
      } else {   // (braces: with `else for (...) {...}` nvcc 12.9 dropped the statement AFTER the loop in the SUB64 build)
      for (int nt = 0; nt < NT; ++nt, ++g) {
        const uint32_t buf = g & 1;
        FZ_DBG(11, g);
        FZ_T();
        mbar_wait(bar(Smem::BAR_T_FULL + buf), (g >> 1) & 1);
        FZ_ACC(2);
        if (nt == 0) FZ_MARK(1 + it * 4);
        if (!have) {                                   // nothing to read: hand the accumulator straight back
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_bar(Smem::BAR_T_EMPTY + buf));
          continue;
        }
        tc_fence_after();
        uint32_t ra[32];
#ifdef DCVIC_FZ_NOLD      // experiment: the pipeline without the accumulator read-out (TMEM read bandwidth)
#pragma unroll
        for (int i_ = 0; i_ < 32; ++i_) ra[i_] = (uint32_t)(g + i_);
#else
        TMEM_LD32(ra, tlane + buf * BN);
        TMEM_WAIT_LD32(ra);
#endif
        // every score of this warp's slice is in registers: the accumulator can be overwritten
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_bar(Smem::BAR_T_EMPTY + buf));
        FZ_ACC(3);
        FZ_ACC(4);
        float cm;
        const float m_old = m;
#ifdef DCVIC_FZ_EPIFREE   // experiment: what the kernel costs without the flag arithmetic (wrong results)
        cm = __uint_as_float(ra[0] ^ ra[31]);
        m = fmaxf(m, cm);
        const uint32_t mask = nt == 0 ? 1u : 0u;
#else
        const uint32_t mask = chunk_flags(ra, margin, m, cm);
#endif
        // A flagged chunk's mask is kept.  When this chunk's maximum beats the previous running maximum by more than
        // the margin, every earlier mask is out of reach (all its scores are <= m_old) and is dropped.
        if (mask != 0u) {
          if (cm > m_old + margin) live = 0u;
          sts_u32(my_mask + nt * (BM * 4), mask);
          live |= 1u << nt;
        }
        FZ_ACC(5);
      }
      }
      // ---- end of tile: the four quarters of a row exchange their maxima; each appends the codes of its live masks
      // to the row's candidate array (slots handed out by an atomic counter: the order does not matter, the re-rank
      // breaks ties by code index).  Masks were taken against the running threshold of their time, which is below the
      // final one: a superset of the codes within the margin of the row maximum.
      sts_u32(a_m + (row * 4 + q) * 4, __float_as_uint(m));
      FZ_MARK(2 + it * 4);
      FZ_DBG(12, it);
      named_bar_sync(1 + part, 128);
      const float4 mq = lds128(a_m + row * 16);
      // (a_m is written again after the NEXT tile's last N-tile, which the MMA issues only once every epilogue warp
      // has drained N-tile NT - 3 of that tile, i.e. has left this section - if the tile has that many)
      if (NT < 3) named_bar_sync(1 + part, 128);
      const float thr = fmaxf(fmaxf(mq.x, mq.y), fmaxf(mq.z, mq.w)) - margin;
      const int par = it & 1;
      FZ_DBG(13, it);
      FZ_ACC(6);
      if (it >= 2) mbar_wait(bar(Smem::BAR_C_EMPTY + par), ((it >> 1) - 1) & 1);   // consumers are done with tile it-2
      FZ_ACC(4);                                     // (slot 4: waiting for the consumers)
      if (valid) {
        const uint32_t ncp = a_nc + (par * BM + row) * 4;
        if (cb_unsafe || !(zz < kVqFp16Zz2Max)) {
          if (q == 0) {
            atoms_add(ncp, (uint32_t)kFullFlag);
            atomicAdd(counters + 9, 1u);              // diagnostics: why a token is scanned in full
          }
        } else if (live != 0u && !(m < thr)) {         // (a quarter whose maximum is out of reach has nothing to add)
          // (64-column buffers: slot sl = 2 nt + sub holds 16 codes from nt * 128 + (q >> 1) * 64 + sub * 32 + (q & 1) * 16)
          auto slot_mask = [&](int sl) {
            return kSub64 ? lds_u16(a_mask + ((q * 16 + sl) * BM + row) * 2) : lds_u32(my_mask + sl * (BM * 4));
          };
          auto slot_code0 = [&](int sl) {
            return kSub64 ? (sl >> 1) * BN + (q >> 1) * (BN / 2) + (sl & 1) * 32 + (q & 1) * 16 : sl * BN + q * kChunk;
          };
          int cnt = 0;
          for (uint32_t lv = live; lv; lv &= lv - 1) cnt += __popc(slot_mask(__ffs(lv) - 1));
          int w = (int)atoms_add(ncp, (uint32_t)cnt);
          if (w <= CK_MAX && w + cnt > CK_MAX) atomicAdd(counters + 7, 1u);
          const uint32_t ck = a_ck + (par * BM + row) * (CK_MAX * 2);
          for (uint32_t lv = live; lv; lv &= lv - 1) {
            const int nt = __ffs(lv) - 1;
            const int c0 = slot_code0(nt);
            for (uint32_t mk = slot_mask(nt); mk; mk &= mk - 1) {
              if (w < CK_MAX) sts_u16(ck + w * 2, (uint32_t)(c0 + __ffs(mk) - 1));
              ++w;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Smem::BAR_C_FULL + par));
      FZ_MARK(3 + it * 4);
      FZ_DBG(14, it);
      FZ_ACC(6);
    }
    FZ_PUT();
    // Experiment (-DDCVIC_FZ_HELPERS; off: measured 58.9 us against 54.8): their own work done, the sixteen epilogue
    // warps help with the finish, two tokens per pass in their 64 registers.  The consumers run behind the MMA by
    // construction and the tail after the last MMA is 7-9 us of eight warps working through dependent L2 / shared-
    // memory round trips with the other 24 idle.  The tail did shrink (11.4 -> 9.9 us), but the second instance of
    // the finish takes the kernel from 4,544 to 6,368 instructions (72 -> 102 KB) and the MMA phase grew by 3 us.
#ifdef DCVIC_FZ_HELPERS
    consume(std::integral_constant<int, 2>{}, NCONS + (warp - W_EPI0));
#endif
  } else {
    // ===================== consumers: the finish, four tokens per pass =====================
    FZ_DBG(22, 0);
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FZ_REGS_CONS));
    FZ_DBG(23, 0);
    pdl_wait();                                      // |e|^2 and Ep come from the prepare kernel
    consume(std::integral_constant<int, 4>{}, warp - W_CONS0);
  }

  FZ_DBG(30, 0);
  FZ_MARK(39);
  FZ_GMARK(36);                                      // role loop left
  tc_fence_before();
  __syncthreads();
  // Loss: every CTA sums its finishing warps' partials (shared memory, slot order) into one value; the CTA that
  // finishes last sums those in index order (deterministic for a given assignment of units to warps) - one
  // device-scope fence and one atomic per CTA instead of a 1-CTA kernel behind this one (3.5 us + a launch gap).
  if (threadIdx.x == 0) {
    double cta = 0.0;                                // this CTA's finishing warps, in slot order
    for (int w = 0; w < NFIN; ++w) cta += s_scratch[w];
    partials[blockIdx.x] = cta;
    __threadfence();
    s_last = atomicAdd(counters + kCtrLoss, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  cluster_sync();            // no CTA leaves while its peer may still touch its shared memory / barriers
  if (s_last) {
    __threadfence();
    const int n = (int)gridDim.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += NTHREADS) acc += __ldcg(partials + i);
    const double tot = block_sum(acc, s_scratch);
    if (threadIdx.x == 0) {
      write_loss(tot, (long long)N * D, beta, legacy, loss);
      counters[kCtrLoss] = 0u;
    }
  }
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
  FZ_GMARK(37);                                      // kernel exit
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn cached = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return (EncodeTiledFn) nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(fn);
  }();
  return cached;
}

static bool map_2d(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, uint64_t inner, uint64_t outer,
                   uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn encode = encode_fn();
  if (!encode) return false;
  (void)esize;
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstride[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  return encode(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int D>
static int launch(const CUtensorMap& tcb, const CUtensorMap& tcb2, const CUtensorMap& tbp, const CUtensorMap& tzc, const CUtensorMap& tzf, const CUtensorMap& tzq,
                  const float* E, const float* ee, const float* emax, int N, int HW, int K,
                  int wait_first, float beta, int legacy, int64_t* idx, float* loss, double* partials,
                  unsigned* counters, cudaStream_t s) {
  const int smem = Smem::bytes(D) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(vq_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return DCVIC_ERR_CUDA;
    attr_set = true;
  }
  const int num_gp = (N + 2 * GT - 1) / (2 * GT);          // group pairs: 64 tokens, 32 for either CTA of a pair
  int max_pairs = kNumSMs / 2;
  if (CL == 4) {
    // clusters of four must fit a GPC: fewer than 148 / 4 of them may be co-resident, and a second wave would double
    // the kernel's time
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(kNumSMs / CL * CL);
      q.blockDim = dim3(NTHREADS);
      q.dynamicSmemBytes = smem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, vq_fused_kernel<D>, &q) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = kNumSMs / CL / 2;
      }
      max_clusters = n < kNumSMs / CL ? n : kNumSMs / CL;
      if (getenv("DCVIC_FZ_VERBOSE")) fprintf(stderr, "vq_fused: %d clusters of %d co-resident (occupancy query: %d)\n", max_clusters, CL, n);
    }
    max_pairs = max_clusters * (CL / 2);
  }
  int npairs = num_gp < max_pairs ? num_gp : max_pairs;             // (small inputs: one group pair per CTA pair)
  if (CL == 4) npairs = (npairs + 1) / 2 * 2;                       // whole clusters
#ifdef DCVIC_FZ_WHOLE_TILES   // (measured: 57.6 us against 55.1 - the finish of a FULL last tile is a longer tail)
  // A partial tile costs a whole pass over the codebook (MMA time and 0.5 MB of L2 traffic per CTA pair), and the
  // kernel runs at the L2's throughput: with two or more tiles per pair, use the FEWEST pairs that keep the number of
  // rounds (64 pairs x 4 whole tiles on the headline shape instead of 74 x (3 + a partial one): 256 codebook passes
  // instead of 296).
  {
    const int gp_per_pair = (num_gp + max_pairs - 1) / max_pairs;
    const int rounds = (gp_per_pair + NG - 1) / NG;
    if (rounds >= 2) npairs = (num_gp + rounds * NG - 1) / (rounds * NG);
  }
#endif
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (cudaLaunchKernelEx(&cfg, vq_fused_kernel<D>, tcb, tcb2, tbp, tzc, tzf, tzq, E, ee, emax, N, HW, K, num_gp,
                         wait_first, beta, legacy, idx, loss, partials, counters) != cudaSuccess)
    return DCVIC_ERR_CUDA;
  return dcvic_launch_status();
}

}  // namespace fz

#ifdef DCVIC_FZ_DEBUG
}  // namespace dcvic
extern "C" int dcvic_debug_set_fz_progress(int* mapped) {
  return cudaMemcpyToSymbol(dcvic::fz::g_fz_dbg, &mapped, sizeof(mapped)) == cudaSuccess ? 0 : -4;
}
namespace dcvic {
#endif

bool vq_fused_supported(const float* z, const float* zq, const float* E, int D, int HW, int K) {
  static const bool disabled = getenv("DCVIC_VQ_FUSED") && atoi(getenv("DCVIC_VQ_FUSED")) == 0;
  if (disabled) return false;
  if (!(D == 64 || D == 128 || D == 192 || D == 256)) return false;
  if (K % fz::BN != 0 || K < fz::BN || K > fz::MAX_K) return false;
  if (HW % fz::GT != 0) return false;
  if ((reinterpret_cast<uintptr_t>(z) & 15) || (reinterpret_cast<uintptr_t>(zq) & 15) ||
      (reinterpret_cast<uintptr_t>(E) & 15))
    return false;
  static const bool sm100 = [] {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return major == 10;
  }();
  return sm100 && fz::encode_fn() != nullptr;
}

int vq_fused_forward(const float* z, const float* E, const float* ee, const float* emax,
                     const __half* cb16, int B, int D, int HW, int K, bool after_prepare, float beta, int legacy,
                     float* zq, int64_t* idx, float* loss, double* partials, unsigned* counters, cudaStream_t s) {
  using namespace fz;
  CUtensorMap tcb, tcb2, tbp, tzc, tzf, tzq;
  const int N = B * HW;
  {
    EncodeTiledFn encode = encode_fn();
    if (!encode) return DCVIC_ERR_DEVICE;
    const int nchunk = D / BK + 1;      // (the chunk-major codebook also carries the pad chunk of the two-kernel path)
    const cuuint64_t gdim[3] = {(cuuint64_t)BK, (cuuint64_t)K, (cuuint64_t)nchunk};
    const cuuint64_t gstride[2] = {(cuuint64_t)BK * 2, (cuuint64_t)K * BK * 2};
    const int na = Smem::b_chunks_a(D);
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int part = 0; part < 2; ++part) {
      const int nz = part == 0 ? na : (D / BK - na > 0 ? D / BK - na : 1);
      const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(BN / 2), (cuuint32_t)nz};
      if (encode(part == 0 ? &tcb : &tcb2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(cb16), gdim, gstride,
                 box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return DCVIC_ERR_CUDA;
    }
  }
  {
    // -|e|^2/2 operand: the first 8 FP16 of every code's row of the pad chunk (chunk e_dim / 64 of the chunk-major
    // codebook: the prepare kernel's three-way split and zeros), 64 codes per box, 16-byte rows, no swizzle
    EncodeTiledFn encode = encode_fn();
    const cuuint64_t gdim[2] = {8, (cuuint64_t)K};
    const cuuint64_t gstride[1] = {(cuuint64_t)BK * 2};
    const cuuint32_t box[2] = {8, (cuuint32_t)(BN / 2)};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&tbp, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(cb16) + (size_t)(D / BK) * K * BK, gdim,
               gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DCVIC_ERR_CUDA;
  }
  if (!map_2d(&tzc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, z, (uint64_t)HW, (uint64_t)B * D, (uint64_t)HW * 4, GT, BK) ||
      !map_2d(&tzf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, z, (uint64_t)HW, (uint64_t)B * D, (uint64_t)HW * 4, GT, D) ||
      !map_2d(&tzq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, zq, (uint64_t)HW, (uint64_t)B * D, (uint64_t)HW * 4, GT, D))
    return DCVIC_ERR_CUDA;
  const int wait_first = after_prepare ? 0 : 1;
  int rc;
#define DCVIC_FZ(DD) \
  launch<DD>(tcb, tcb2, tbp, tzc, tzf, tzq, E, ee, emax, N, HW, K, wait_first, beta, legacy, idx, loss, partials, counters, s)
  switch (D) {
    case 64: rc = DCVIC_FZ(64); break;
    case 128: rc = DCVIC_FZ(128); break;
    case 192: rc = DCVIC_FZ(192); break;
    case 256: rc = DCVIC_FZ(256); break;
    default: return DCVIC_ERR_UNSUPPORTED;
  }
#undef DCVIC_FZ
  return rc;
}

}  // namespace dcvic
