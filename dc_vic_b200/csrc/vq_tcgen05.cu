// Wide-codebook nearest-codeword CANDIDATE search on the 5th-gen tensor cores (sm_100a).
//
// Computes, for every token row z_t (FP32, read straight from NCHW) and every codeword e_k,
//     s[t,k] = fp16(z_t) . fp16(e_k) - |e_k|^2 / 2        ( = (|z|^2 - d[t,k]) / 2 up to FP16 rounding )
// with tcgen05.mma (FP16 x FP16 -> FP32 in TMEM; FP16 rather than BF16 because its 11-bit significand keeps
// the proven margin 8x tighter, i.e. ~1.2 instead of ~2.7 FP32 re-rank candidates per token) and flags, per row, every k whose score is within a
// PROVEN error margin (vq_margin) of the row's running maximum.  The distance matrix never leaves
// the SM.  The FP32 re-rank of the few flagged codes (reference op order, lowest-index tie-break)
// happens in vq_finish_kernel, which streams z once more for the gather / STE / loss.
//
// Structure: persistent CTA pairs (cta_group::2, UMMA 256x256x16), each CTA owning 128 tokens of a
// 256-token pair tile and half of every codebook tile (CG = 1 is the same code on single CTAs).
//   warps 0-7   A producers: read z (FP32, 16-byte loads of 4 consecutive tokens, two 8-load sets in
//               flight per thread, running ahead into the next tile),
//               convert to FP16, write the K-major SWIZZLE_128B operand tile (double-buffered: tile i+1
//               loads while tile i multiplies), publish |z|^2 per token
//   warps 8-15  epilogue: warps 8-11 take columns 0-127 of every accumulator, warps 12-15 columns
//               128-255.  tcgen05.ld 32 scores per row at a time (software pipelined), running max, one
//               flag mask per 32 codes (FADD on the FMA pipe + funnel shift, branch-free), append
//               {chunk max | chunk id, mask} to the token's list in global memory when non-empty.  The
//               accumulator goes back to the MMA as soon as its last scores are in registers.
//   warp 16     TMA producer: this CTA's half of the FP16 codebook tile [256/CG codes x 64 ch]
//               (SWIZZLE_128B) into a 4-stage ring; completion is signalled on the LEADER's barrier
//   warp 17     TMEM allocator; in the leader CTA one thread issues every tcgen05.mma of the pair and
//               multicasts the commits (stage free, accumulator full, operand tile free) to both CTAs
// -|e|^2/2 enters through the contraction itself: the FP16 codebook carries one extra 64-column chunk
// whose first three columns are a 3-way FP16 split of -|e_k|^2/2, multiplied (one K=16 step, which
// also zero-initialises the accumulator) with a constant operand chunk of ones.
// Reference semantics: taming/modules/vqvae/quantize.py:280-284 (distance + argmin).
#include "vq_common.cuh"
#include "sm100_ptx.cuh"
#include <cuda.h>

namespace dcvic {

namespace tc {

constexpr int BM = 128;          // tokens per CTA tile (UMMA M = 128 * CG)
constexpr int BN = 256;          // codes per N-tile (UMMA N)
constexpr int BK = 64;           // channels per smem chunk: 64 fp16 = 128 B = one SWIZZLE_128B row
constexpr int MAX_KC = 4;        // e_dim <= 256
constexpr int A_CHUNK_BYTES = BM * BK * 2;            // 16 KB
constexpr int A_BUF_BYTES = MAX_KC * A_CHUNK_BYTES;   // 64 KB
constexpr int MAX_K = 4096;
#ifndef DCVIC_SEARCH_SETS
#define DCVIC_SEARCH_SETS 3       // register sets of operand loads in flight per producer thread (2: the older loop)
#endif
// Three producer sets: 20 warps (two idle ones complete the fifth warpgroup, so that every setmaxnreg is executed by a
// whole warpgroup) launch with 96 registers each = 61,440; producers take 136, epilogue warps keep 88, the
// TMA / MMA / idle warpgroup drops to 32:  8*32*136 + 8*32*88 + 4*32*32 = 61,440.
#ifndef DCVIC_SEARCH_PROD_REGS
#define DCVIC_SEARCH_PROD_REGS 136
#endif
#ifndef DCVIC_SEARCH_EPI_REGS
#define DCVIC_SEARCH_EPI_REGS 88
#endif
#ifndef DCVIC_SEARCH_AUX_REGS
#define DCVIC_SEARCH_AUX_REGS 32
#endif
#if DCVIC_SEARCH_SETS == 3
constexpr int NTHREADS = 640;
#else
constexpr int NTHREADS = 576;
#endif
constexpr int NPROD = 8;                       // A-producer warps (warps 0-7); epilogue warps 8-15
constexpr int WARP_TMA = 16, WARP_MMA = 17;
constexpr int ZZ_SLOTS = 4;

template <int CG>
struct Cfg {
  static constexpr int B_STAGE_BYTES = (BN / CG) * BK * 2;       // 32 KB (CG=1) / 16 KB (CG=2)
  static constexpr int NSTAGE = CG == 2 ? 4 : 2;
  // dynamic shared memory map (base aligned to 1024 B)
  static constexpr int OFF_A = 0;                                  // [2][MAX_KC][BM x 128 B]
  static constexpr int OFF_APAD = OFF_A + 2 * A_BUF_BYTES;         // [BM x 128 B] constant: ones in columns 0-2
  static constexpr int OFF_B = OFF_APAD + A_CHUNK_BYTES;           // [NSTAGE][BN/CG x 128 B]
  static constexpr int OFF_ZZ = OFF_B + NSTAGE * B_STAGE_BYTES;    // [ZZ_SLOTS][2 channel halves][BM] float
  static constexpr int OFF_BAR = OFF_ZZ + 2 * ZZ_SLOTS * BM * 4;
  // barrier slots (8 bytes each)
  static constexpr int BAR_B_FULL = 0;                       // [NSTAGE]   leader only
  static constexpr int BAR_B_EMPTY = BAR_B_FULL + NSTAGE;    // [NSTAGE]
  static constexpr int BAR_A_FULL = BAR_B_EMPTY + NSTAGE;    // [2][MAX_KC] leader only
  static constexpr int BAR_A_EMPTY = BAR_A_FULL + 2 * MAX_KC;  // [2]
  static constexpr int BAR_T_FULL = BAR_A_EMPTY + 2;         // [2]
  static constexpr int BAR_T_EMPTY = BAR_T_FULL + 2;         // [2]        leader only
  static constexpr int BAR_ZZ = BAR_T_EMPTY + 2;             // [ZZ_SLOTS]
  static constexpr int BAR_COUNT = BAR_ZZ + ZZ_SLOTS;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + BAR_COUNT * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16;
  static_assert(SMEM_BYTES + 1024 <= 232448, "shared memory budget");
};

// kind::f16 instruction descriptor: D=F32 (bit 4), A=B=F16 (format 0), both K-major, N=256, M=128*CG
template <int CG>
struct Idesc {
  static constexpr uint32_t value =
      (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
};

template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  if constexpr (CG == 2) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(Idesc<2>::value), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(Idesc<1>::value), "r"(accumulate)
        : "memory");
  }
}
}  // namespace tc

using namespace tc;

// Optional role timing (build with -DDCVIC_TRACE; tools/trace_run.py): cycles each warp role spends
// waiting / working, summed over the launch, per CTA.
// Role timing (-DDCVIC_TRACE): the MMA issuer's waits always; the other roles only with -DDCVIC_TRACE_ALL (their
// counters cost registers that the setmaxnreg split does not have to spare, and perturb the kernel by 30 %).
#ifdef DCVIC_TRACE
__device__ unsigned long long g_trace[kNumSMs * 2][16];
#define TRM_NOW() clock64()
#define TRM_ADD(acc, t0) acc += clock64() - (t0)
#define TR_PUT(slot, v) g_trace[blockIdx.x][slot] = (v)
#else
#define TRM_NOW() 0ull
#define TRM_ADD(acc, t0) (void)(t0)
#define TR_PUT(slot, v)
#endif
#if defined(DCVIC_TRACE) && defined(DCVIC_TRACE_ALL)
#define TR_NOW() clock64()
#define TR_ADD(acc, t0) acc += clock64() - (t0)
#else
#define TR_NOW() 0ull
#define TR_ADD(acc, t0) (void)(t0)
#endif

template <int CG>
__global__ void __launch_bounds__(NTHREADS, 1)
vq_tensor_search_kernel(const __grid_constant__ CUtensorMap tmap_cb, const float* __restrict__ z,
                        const float* __restrict__ emax_ptr, int N, int D, int HW, int K, int num_ptiles,
                        int wait_first, int split, VqMeta* __restrict__ meta, uint2* __restrict__ list) {
  using C = Cfg<CG>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment (same adjustment in both CTAs of a pair)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + C::OFF_BAR;
  auto bar = [&](int slot) { return bar0 + slot * 8; };
  float* s_zz = reinterpret_cast<float*>(smem + C::OFF_ZZ);
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM_PTR);

  // PDL: this kernel may have started before its predecessor (codebook prepare, or the previous call's finish) ended;
  // the roles that read the predecessor's outputs wait for it below, the z loads do not.  Its own successor (finish)
  // may start as soon as SMs free up: it only touches z until it waits for this grid.
  // wait_first: the predecessor in the stream is not one of this library's kernels (frozen codebook, no prepare
  // launch) and may be the producer of z, so nothing here - and, through the trigger below, nothing in the finish
  // kernel - may read z before it has completed.  (The wait and the trigger sit behind the set-up below - barriers,
  // constant operand chunk, TMEM allocation, cluster sync touch nothing a predecessor wrote - so that in wait_first
  // mode the set-up still overlaps the predecessor's tail.)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int pair = blockIdx.x / CG, npairs = gridDim.x / CG;
  const int KC = D / BK;               // channel chunks
  const int NT = K / BN;               // N-tiles per token tile
  // Schedule: `rounds` full rounds of one token tile per pair; the remaining `rem` tiles (fewer than pairs) form
  // the last round.  With split == 2 (host: 2 * rem <= npairs) every one of them goes to TWO pairs, each scanning
  // half of the codebook's N-tiles, so the last round takes half as long (40 of 74 pairs would idle through it on
  // C2).  The two halves keep separate running maxima and list halves (VqMeta part 0 / part 1).
  const int rounds = num_ptiles / npairs, rem = num_ptiles % npairs;
  const int my_tiles = rounds + (pair < rem * split ? 1 : 0);
  auto unit_ptile = [&](int it) { return it < rounds ? pair + it * npairs : rounds * npairs + pair / split; };
  auto unit_part = [&](int it) { return it < rounds ? 0 : pair % split; };
  auto unit_nt0 = [&](int it) { return (it < rounds || split == 1) ? 0 : (pair % split) * (NT / split); };
  auto unit_nt1 = [&](int it) { return (it < rounds || split == 1) ? NT : (pair % split + 1) * (NT / split); };
  auto unit_split = [&](int it) { return it >= rounds && split > 1; };
  // barriers that live in the leader CTA, as shared::cluster addresses
  auto leader_bar = [&](int slot) { return CG == 2 ? map_to_cta(bar(slot), 0) : bar(slot); };
  auto arrive_leader = [&](uint32_t b) { if constexpr (CG == 2) mbar_arrive_cluster(b); else mbar_arrive(b); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(bar(C::BAR_B_FULL + s), 1); mbar_init(bar(C::BAR_B_EMPTY + s), 1); }
    for (int c = 0; c < 2 * MAX_KC; ++c) mbar_init(bar(C::BAR_A_FULL + c), NPROD * CG);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(C::BAR_A_EMPTY + b), 1);
      mbar_init(bar(C::BAR_T_FULL + b), 1);
      mbar_init(bar(C::BAR_T_EMPTY + b), 8 * CG);
    }
    for (int s = 0; s < ZZ_SLOTS; ++s) mbar_init(bar(C::BAR_ZZ + s), NPROD);
    fence_barrier_init();
  }
  // constant operand chunk for the -|e|^2/2 step: fp16 1.0 in columns 0-2 of every row.  Columns 0-7 sit in
  // 16-byte piece (0 ^ (row & 7)) of the row (SWIZZLE_128B), all other pieces are zero.
  for (int i = threadIdx.x; i < BM * 8; i += NTHREADS) {
    const int row = i >> 3, piece = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (piece == (row & 7)) { v.x = 0x3C003C00u; v.y = 0x00003C00u; }
    *reinterpret_cast<uint4*>(smem + C::OFF_APAD + row * 128 + piece * 16) = v;
  }
  fence_proxy_async();
  if (warp == WARP_MMA) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::OFF_TMEM_PTR),
                   "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::OFF_TMEM_PTR),
                   "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync();     // peer's barriers / constant chunk are in place before anyone uses them
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (wait_first) pdl_wait();
  pdl_launch_dependents();

  if (warp < NPROD) {
    // ===================== A producers: FP32 NCHW -> FP16 K-major SWIZZLE_128B =====================
    // Warp w owns tokens [32(w&3), +32) of the tile and channel half ch = w>>2 of every 64-channel chunk.
    // lane = tq*4 + cg: token quad tq (4 consecutive tokens, one 16-byte load per channel) and channel
    // group cg (every quarter warp then covers all 8 swizzled bank groups in its 16-byte stores).  One step = 8 channels x 4 tokens per thread (8 LDG.128, 512 contiguous bytes per
    // channel and warp); two steps are in flight per thread.  The smem store of one (token, 8 channels)
    // 16-byte piece hits 8 distinct swizzled bank groups across the warp (4 wavefronts, the minimum
    // for 512 bytes).
    const int cg = lane & 3, tq = lane >> 2, ch = warp >> 2;
    const int row0 = (warp & 3) * 32 + tq * 4;
    const int g = cg + 4 * ch;                  // 8-channel group inside a 64-channel chunk
    const size_t sHW = (size_t)HW;
    [[maybe_unused]] unsigned long long tr_wait = 0, tr_work = 0;
    float4 va[8], vb[8];
    auto tile_ptr = [&](int it, bool& valid) {
      const long long t = ((long long)unit_ptile(it) * CG + rank) * BM + row0;
      valid = it < my_tiles && t < N;           // N and HW are multiples of 4: a quad is valid as a whole
      return z + (valid ? ((size_t)(t / HW) * D * HW + (size_t)(t % HW) + (size_t)(g * 8) * sHW) : 0);
    };
    auto load_step = [&](float4 (&v)[8], const float* zc, bool valid, int kc) {
      const float* p = zc + (size_t)(BK * kc) * sHW;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        v[k] = valid ? ldg_stream(reinterpret_cast<const float4*>(p + (size_t)k * sHW)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
#if DCVIC_SEARCH_SETS == 3
    // Three register sets in flight per thread (96 KB per SM instead of 64: the producers, not the tensor pipe, set
    // the pace of this kernel).  The steps of all this pair's tiles form one stream s = 0, 1, 2, ...
    // (tile = s / KC, chunk = s % KC); step s lives in set s % 3: store it, then request step s + 3 into the same
    // registers.  The extra registers come from setmaxnreg (the epilogue warps give some of theirs up).
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DCVIC_SEARCH_PROD_REGS));
    float4 vc[8];
    const int total_steps = my_tiles * KC;
    // load cursor (three steps ahead of the store cursor)
    int lit = 0, lkc = 0;
    bool lvalid;
    const float* lzc = tile_ptr(0, lvalid);
    auto load_next = [&](float4 (&v)[8]) {
      load_step(v, lzc, lvalid, lkc);
      if (++lkc == KC) { lkc = 0; ++lit; lzc = tile_ptr(lit, lvalid); }
    };
    // store cursor
    int sit = 0, skc = 0;
    float zz4[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned long long tr0 = 0;
    auto store_next = [&](const float4 (&v)[8]) {
      const int abuf = sit & 1;
      if (skc == 0) {
        tr0 = TR_NOW();
        mbar_wait(bar(C::BAR_A_EMPTY + abuf), ((sit >> 1) & 1) ^ 1);
        TR_ADD(tr_wait, tr0);
        tr0 = TR_NOW();
#pragma unroll
        for (int i = 0; i < 4; ++i) zz4[i] = 0.f;
      }
      uint8_t* a = smem + C::OFF_A + abuf * A_BUF_BYTES + row0 * 128 + skc * A_CHUNK_BYTES;
      const float* f = reinterpret_cast<const float*>(v);     // f[4k + i] = channel k, token i
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 pk;
        pk.x = pack_f16x2(f[0 + i], f[4 + i]);
        pk.y = pack_f16x2(f[8 + i], f[12 + i]);
        pk.z = pack_f16x2(f[16 + i], f[20 + i]);
        pk.w = pack_f16x2(f[24 + i], f[28 + i]);
        const int r7 = (row0 + i) & 7;
        *reinterpret_cast<uint4*>(a + i * 128 + ((g ^ r7) << 4)) = pk;
#pragma unroll
        for (int k = 0; k < 8; ++k) zz4[i] = fmaf(f[4 * k + i], f[4 * k + i], zz4[i]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive_leader(leader_bar(C::BAR_A_FULL + abuf * MAX_KC + skc));
      if (skc == KC - 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          zz4[i] += __shfl_xor_sync(0xffffffffu, zz4[i], 1);
          zz4[i] += __shfl_xor_sync(0xffffffffu, zz4[i], 2);
        }
        if (cg == 0)
          *reinterpret_cast<float4*>(s_zz + ((sit & (ZZ_SLOTS - 1)) * 2 + ch) * BM + row0) =
              make_float4(zz4[0], zz4[1], zz4[2], zz4[3]);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(C::BAR_ZZ + (sit & (ZZ_SLOTS - 1))));
        TR_ADD(tr_work, tr0);
      }
      if (++skc == KC) { skc = 0; ++sit; }
    };
    load_next(va);
    load_next(vb);
    load_next(vc);
#pragma unroll 1
    for (int s0 = 0; s0 < total_steps; s0 += 3) {
      store_next(va);
      load_next(va);
      if (s0 + 1 < total_steps) { store_next(vb); load_next(vb); }
      if (s0 + 2 < total_steps) { store_next(vc); load_next(vc); }
    }
#else
    bool valid;
    const float* zc = tile_ptr(0, valid);
    load_step(va, zc, valid, 0);
    if (KC > 1) load_step(vb, zc, valid, 1);
    for (int it = 0; it < my_tiles; ++it) {
      const int abuf = it & 1;
      unsigned long long tr0 = TR_NOW();
      mbar_wait(bar(C::BAR_A_EMPTY + abuf), ((it >> 1) & 1) ^ 1);
      TR_ADD(tr_wait, tr0);
      tr0 = TR_NOW();
      float zz4[4] = {0.f, 0.f, 0.f, 0.f};
      uint8_t* abase = smem + C::OFF_A + abuf * A_BUF_BYTES + row0 * 128;
      auto store_step = [&](const float4 (&v)[8], int kc) {
        uint8_t* a = abase + kc * A_CHUNK_BYTES;
        const float* f = reinterpret_cast<const float*>(v);     // f[4k + i] = channel k, token i
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 pk;
          pk.x = pack_f16x2(f[0 + i], f[4 + i]);
          pk.y = pack_f16x2(f[8 + i], f[12 + i]);
          pk.z = pack_f16x2(f[16 + i], f[20 + i]);
          pk.w = pack_f16x2(f[24 + i], f[28 + i]);
          const int r7 = (row0 + i) & 7;
          *reinterpret_cast<uint4*>(a + i * 128 + ((g ^ r7) << 4)) = pk;
#pragma unroll
          for (int k = 0; k < 8; ++k) zz4[i] = fmaf(f[4 * k + i], f[4 * k + i], zz4[i]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) arrive_leader(leader_bar(C::BAR_A_FULL + abuf * MAX_KC + kc));
      };
      // the next tile's first two steps are requested before this tile's last two are stored, so the
      // memory pipe never drains between tiles (registers, not the smem buffer, are the landing zone)
      bool nvalid;
      const float* nzc = tile_ptr(it + 1, nvalid);
#pragma unroll 1
      for (int kc = 0; kc < KC; kc += 2) {
        store_step(va, kc);
        if (kc + 2 < KC) load_step(va, zc, valid, kc + 2); else load_step(va, nzc, nvalid, 0);
        if (kc + 1 < KC) store_step(vb, kc + 1);
        if (kc + 3 < KC) load_step(vb, zc, valid, kc + 3); else if (KC > 1) load_step(vb, nzc, nvalid, 1);
      }
      zc = nzc;
      valid = nvalid;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        zz4[i] += __shfl_xor_sync(0xffffffffu, zz4[i], 1);
        zz4[i] += __shfl_xor_sync(0xffffffffu, zz4[i], 2);
      }
      if (cg == 0)
        *reinterpret_cast<float4*>(s_zz + ((it & (ZZ_SLOTS - 1)) * 2 + ch) * BM + row0) =
            make_float4(zz4[0], zz4[1], zz4[2], zz4[3]);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(C::BAR_ZZ + (it & (ZZ_SLOTS - 1))));
      TR_ADD(tr_work, tr0);
    }
#endif
    if (threadIdx.x == 0) { TR_PUT(0, tr_wait); TR_PUT(1, tr_work); }
  } else if (warp < NPROD + 8) {
    // ===================== epilogue: flag masks per 32 codes, running max per row =====================
#if DCVIC_SEARCH_SETS == 3
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DCVIC_SEARCH_EPI_REGS));
#endif
    const int q = (warp - NPROD) >> 2;            // column half of every accumulator this warp quad drains
    const int part = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = part * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(part * 32) << 16) + q * (BN / 2);
    pdl_wait();                                   // emax (prepare kernel); the previous call's finish is done with meta/list
    const float emax = emax_ptr[0];
    const bool cb_unsafe = __float_as_uint(emax_ptr[1]) != 0u;   // codebook outside FP16's range: FP32 scan for all
    uint32_t g = 0;                               // running N-tile counter (same sequence as the MMA issuer)
    [[maybe_unused]] unsigned long long tr_zz = 0, tr_full = 0, tr_proc = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int ptile = unit_ptile(it);
      const int part = unit_part(it), nt0 = unit_nt0(it), nt1 = unit_nt1(it);
      const bool is_split = unit_split(it);
      const int cap = is_split ? kListCap / 2 : kListCap;   // entries this unit may write per list
      const long long t = ((long long)ptile * CG + rank) * BM + row;
      const bool valid = t < N;
      unsigned long long tr0 = TR_NOW();
      mbar_wait(bar(C::BAR_ZZ + (it & (ZZ_SLOTS - 1))), (it / ZZ_SLOTS) & 1);
      TR_ADD(tr_zz, tr0);
      const float zz = s_zz[(it & (ZZ_SLOTS - 1)) * 2 * BM + row] + s_zz[((it & (ZZ_SLOTS - 1)) * 2 + 1) * BM + row];
      const float margin = vq_margin(zz, emax);
      float m = -INFINITY;
      int n = 0;                                  // list entries written
      uint2* my_list = list + ((size_t)(valid ? t : 0) * 2 + q) * kListCap + part * (kListCap / 2);
      for (int nt = nt0; nt < nt1; ++nt, ++g) {
        const uint32_t buf = g & 1;
        tr0 = TR_NOW();
        mbar_wait(bar(C::BAR_T_FULL + buf), (g >> 1) & 1);
        TR_ADD(tr_full, tr0);
        tr0 = TR_NOW();
        tc_fence_after();
        const uint32_t taddr = tlane + buf * BN;
        const uint32_t chunk0 = (uint32_t)(nt * (BN / 32) + q * (BN / 64));
        auto emit = [&](uint32_t mask, float cm, uint32_t chunk) {
          if (mask != 0u && valid) {
            const uint32_t key = (__float_as_uint(cm) & 0xFFFFFF80u) | chunk;
            if ((n & 3) == 0 && n < cap) {
              // first entry of a 32-byte sector: write the whole sector (entry + zeros), so that the finish kernel's
              // read of it is an L2 hit and not a DRAM fill of the bytes nobody wrote
              uint4* p = reinterpret_cast<uint4*>(my_list + n);
              p[0] = make_uint4(key, mask, 0u, 0u);
              p[1] = make_uint4(0u, 0u, 0u, 0u);
            } else {
              my_list[min(n, cap - 1)] = make_uint2(key, mask);
            }
            ++n;
          }
          __syncwarp();
        };
        uint32_t ra[32], rb[32];
        float cm;
        uint32_t mask;
        TMEM_LD32(ra, taddr);
        TMEM_WAIT_LD32(ra);
        TMEM_LD32(rb, taddr + 32);
        mask = chunk_flags(ra, margin, m, cm);
        emit(mask, cm, chunk0);
        TMEM_WAIT_LD32(rb);
        TMEM_LD32(ra, taddr + 64);
        mask = chunk_flags(rb, margin, m, cm);
        emit(mask, cm, chunk0 + 1);
        TMEM_WAIT_LD32(ra);
        TMEM_LD32(rb, taddr + 96);
        mask = chunk_flags(ra, margin, m, cm);
        emit(mask, cm, chunk0 + 2);
        TMEM_WAIT_LD32(rb);
        // every score of this warp's slice is in registers: the accumulator can be overwritten
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(leader_bar(C::BAR_T_EMPTY + buf));
        mask = chunk_flags(rb, margin, m, cm);
        emit(mask, cm, chunk0 + 3);
        TR_ADD(tr_proc, tr0);
      }
      if (valid) {
        const short nn = (n > cap || cb_unsafe || !(zz < kVqFp16Zz2Max)) ? (short)-1 : (short)n;
        VqMeta* mp = meta + t;
        if (part == 0) {
          if (q == 0) { mp->m0 = m; mp->n0 = nn; mp->zz = zz; mp->split = is_split ? 1u : 0u; }
          else        { mp->m1 = m; mp->n1 = nn; }
          if (!is_split) {                       // no second half: its fields read as "nothing flagged"
            if (q == 0) { mp->mb0 = -INFINITY; mp->nb0 = 0; }
            else        { mp->mb1 = -INFINITY; mp->nb1 = 0; }
          }
        } else {
          if (q == 0) { mp->mb0 = m; mp->nb0 = nn; }
          else        { mp->mb1 = m; mp->nb1 = nn; }
        }
      }
    }
    if (part == 0 && lane == 0) { TR_PUT(2 + 4 * q, tr_zz); TR_PUT(3 + 4 * q, tr_full); TR_PUT(4 + 4 * q, tr_proc); }
  } else if (warp == WARP_TMA) {
    // ===================== TMA producer: this CTA's half of every codebook tile =====================
#if DCVIC_SEARCH_SETS == 3
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DCVIC_SEARCH_AUX_REGS));
#endif
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_cb) : "memory");
      pdl_wait();                                 // the FP16 codebook is written by the prepare kernel
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it)
        for (int nt = unit_nt0(it); nt < unit_nt1(it); ++nt)
          for (int kc = -1; kc < KC; ++kc) {      // kc == -1: the -|e|^2/2 chunk (columns D .. D+63)
            mbar_wait(bar(C::BAR_B_EMPTY + stage), phase ^ 1);
            if (leader) mbar_arrive_expect_tx(bar(C::BAR_B_FULL + stage), CG * C::B_STAGE_BYTES);
            // chunk-major FP16 codebook: rows [chunk * K, (chunk + 1) * K) of a 64-column matrix; the pad is chunk KC
            tma_load_2d<CG>(sbase + C::OFF_B + stage * C::B_STAGE_BYTES, &tmap_cb, 0,
                            (kc < 0 ? KC : kc) * K + nt * BN + (int)rank * (BN / CG), leader_bar(C::BAR_B_FULL + stage));
            if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
          }
    }
  } else {
    // ===================== MMA issuer (leader CTA, one thread); idle warps of the last warpgroup ==========
#if DCVIC_SEARCH_SETS == 3
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DCVIC_SEARCH_AUX_REGS));
#endif
    if (warp == WARP_MMA && leader) {
      // The whole warp walks the loop (warp-uniform control flow: barrier addresses, descriptors and TMEM addresses
      // stay in uniform registers); one elected lane issues the MMAs and the commits.  With a single active lane the
      // compiler wrapped every tcgen05 instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop and the ~90 dependent
      // instructions per 64-channel chunk were on the kernel's critical path (tools/trace_run.py: the issuer was busy
      // 2x the tensor pipe's own time).
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;   // running N-tile counter -> TMEM buffer g & 1
      const uint32_t rt_one = my_tiles > 0 ? 1u : 0u;   // 1, but not a compile-time constant
      const uint32_t rt_zero = my_tiles < 0 ? 1u : 0u;  // 0, likewise
      const uint64_t pad_desc = umma_desc_sw128(sbase + C::OFF_APAD);
      [[maybe_unused]] unsigned long long tr_te = 0, tr_af = 0, tr_bf = 0, tr0;
      [[maybe_unused]] const unsigned long long tr_start = TRM_NOW();
      for (int it = 0; it < my_tiles; ++it) {
        const int abuf = it & 1;
        const int nt0 = unit_nt0(it), nt1 = unit_nt1(it);
        const uint64_t a_desc0 = umma_desc_sw128(sbase + C::OFF_A + abuf * A_BUF_BYTES);
        for (int nt = nt0; nt < nt1; ++nt, ++g) {
          const uint32_t buf = g & 1;
          tr0 = TRM_NOW();
          if (g >= 2) mbar_wait(bar(C::BAR_T_EMPTY + buf), ((g >> 1) - 1) & 1);   // both CTAs drained this buffer
          TRM_ADD(tr_te, tr0);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * BN;
          // the -|e|^2/2 step (overwrites the buffer): constant operand chunk x pad columns of the codebook
          {
            tr0 = TRM_NOW();
            mbar_wait(bar(C::BAR_B_FULL + stage), phase);
            TRM_ADD(tr_bf, tr0);
            tc_fence_after();
            if (issuer) {
              umma_f16<CG>(tmem_d, pad_desc, umma_desc_sw128(sbase + C::OFF_B + stage * C::B_STAGE_BYTES), rt_zero);
              umma_commit<CG>(bar(C::BAR_B_EMPTY + stage));
            }
            if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
          }
          for (int kc = 0; kc < KC; ++kc) {
            tr0 = TRM_NOW();
            if (nt == nt0) mbar_wait(bar(C::BAR_A_FULL + abuf * MAX_KC + kc), (it >> 1) & 1);
            TRM_ADD(tr_af, tr0);
            tr0 = TRM_NOW();
            mbar_wait(bar(C::BAR_B_FULL + stage), phase);
            TRM_ADD(tr_bf, tr0);
            tc_fence_after();
            // Descriptors advance by 32 bytes (2 in the >>4 address field) per K=16 step.  The accumulate flags
            // are run-time values on purpose: with a literal 0 ptxas 12.9 emitted a predicated UTCHMMA whose
            // result was wrong in cta_group::2 (caught by the D1 / D1b parity tests).
            if (issuer) {
              const uint64_t ad = a_desc0 + (uint64_t)((kc * A_CHUNK_BYTES) >> 4);
              const uint64_t bd = umma_desc_sw128(sbase + C::OFF_B + stage * C::B_STAGE_BYTES);
              umma_f16<CG>(tmem_d, ad, bd, rt_one);
              umma_f16<CG>(tmem_d, ad + 2, bd + 2, rt_one);
              umma_f16<CG>(tmem_d, ad + 4, bd + 4, rt_one);
              umma_f16<CG>(tmem_d, ad + 6, bd + 6, rt_one);
              umma_commit<CG>(bar(C::BAR_B_EMPTY + stage));     // frees the codebook stage when these MMAs retire
            }
            if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
          }
          if (issuer) umma_commit<CG>(bar(C::BAR_T_FULL + buf));          // accumulator ready for the epilogue warps
        }
        if (issuer) umma_commit<CG>(bar(C::BAR_A_EMPTY + abuf));          // operand tile may be overwritten
        __syncwarp();
      }
      if (lane == 0) { TR_PUT(10, tr_te); TR_PUT(11, tr_af); TR_PUT(12, tr_bf); TR_PUT(13, TRM_NOW() - tr_start); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync();     // no CTA leaves while its peer may still touch its smem / barriers
  if (warp == WARP_MMA) {
    tc_fence_after();
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<EncodeTiledFn>(fn);
}

static bool device_is_sm100() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return major == 10;
}

// (callers additionally require H*W % 4 == 0: the operand producer reads 4 tokens per 16-byte load)
bool vq_tensor_supported(int D, int K) {
  if (D % BK != 0 || D < BK || D > MAX_KC * BK) return false;
  if (K % BN != 0 || K < BN || K > MAX_K) return false;
  return device_is_sm100();
}

template <int CG>
static int launch_search(const CUtensorMap& tmap, const float* z, const float* emax, int N, int D, int HW, int K,
                         int wait_first, bool allow_split, VqMeta* meta, uint2* list, cudaStream_t s) {
  using C = Cfg<CG>;
  const int num_ptiles = (N + BM * CG - 1) / (BM * CG);
  static const int pairs_env = getenv("DCVIC_VQ_PAIRS") ? atoi(getenv("DCVIC_VQ_PAIRS")) : 0;   // experiments
  const int max_pairs = (pairs_env > 0 && pairs_env < kNumSMs / CG) ? pairs_env : kNumSMs / CG;
  int npairs = num_ptiles < max_pairs ? num_ptiles : max_pairs;
  // last-round split (see the kernel): two pairs per leftover tile when they fit; small inputs split every tile
  static const bool split_off = getenv("DCVIC_VQ_SPLIT") && atoi(getenv("DCVIC_VQ_SPLIT")) == 0;
  const int NT = K / BN;
  const bool can_split = allow_split && !split_off && NT % 2 == 0;
  if (can_split && 2 * num_ptiles <= max_pairs) npairs = 2 * num_ptiles;
  const int rem = num_ptiles % npairs;
  const int split = (can_split && rem > 0 && 2 * rem <= npairs) ? 2 : 1;
  const int grid = CG * npairs;
  const int smem = C::SMEM_BYTES + 1024;
  if (cudaFuncSetAttribute(vq_tensor_search_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
      cudaSuccess)
    return DCVIC_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (cudaLaunchKernelEx(&cfg, vq_tensor_search_kernel<CG>, tmap, z, emax, N, D, HW, K, num_ptiles, wait_first, split, meta,
                         list) != cudaSuccess)
    return DCVIC_ERR_CUDA;
  return dcvic_launch_status();
}

#ifdef DCVIC_TRACE
}  // namespace dcvic
extern "C" int dcvic_debug_read_trace(unsigned long long* host_out /* [296][16] */) {
  return cudaMemcpyFromSymbol(host_out, dcvic::g_trace, sizeof(dcvic::g_trace)) == cudaSuccess ? 0 : -4;
}
namespace dcvic {
#endif

int vq_tensor_search(const float* z, const __half* cb16, const float* emax, int B, int D, int HW, int K,
                     bool after_prepare, bool allow_split, VqMeta* meta, uint2* list, cudaStream_t s) {
  if (!vq_tensor_supported(D, K)) return DCVIC_ERR_UNSUPPORTED;
  static const int cta_group = [] {
    const char* e = getenv("DCVIC_VQ_CTA_GROUP");
    return (e && e[0] == '1') ? 1 : 2;
  }();
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) return DCVIC_ERR_DEVICE;
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)BK, (cuuint64_t)K * (cuuint64_t)((D + kCb16Pad) / BK)};
  const cuuint64_t gstride[1] = {(cuuint64_t)BK * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(BN / cta_group)};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(cb16), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return DCVIC_ERR_CUDA;
  const int N = B * HW;
  const int wait_first = after_prepare ? 0 : 1;
  return cta_group == 2 ? launch_search<2>(tmap, z, emax, N, D, HW, K, wait_first, allow_split, meta, list, s)
                        : launch_search<1>(tmap, z, emax, N, D, HW, K, wait_first, allow_split, meta, list, s);
}

}  // namespace dcvic
