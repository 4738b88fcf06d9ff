// Inline-PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarriers, cluster addressing, TMA loads, UMMA
// descriptors / commits, tensor-memory loads and stores, and the per-chunk flag computation of the VQ epilogue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace dcvic {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier given by its shared::cluster address (possibly in the peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_bar) {   // cluster-scope release
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (The polling loop stays inside the asm block: a C++ loop around try_wait measured 2 us slower per launch.)
// The suspend-time hint matters: without it try_wait comes back after a few tens of cycles and the waiting warps
// (most of a warp-specialised CTA at any moment) spend the schedulers' issue slots on TRYWAIT / BRA pairs - half of
// all instructions the single-pass VQ kernel executed.  With the hint the warp sleeps until the phase completes.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  if constexpr (CG == 2) {
    // `bar` is a shared::cluster address (the leader CTA's barrier)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
        "%3}], [%4];" ::"r"(smem_dst),
        "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_dst),
        "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
  }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
// rows are 128 B, 8-row groups 1024 B apart (SBO), LBO unused for swizzled K-major (canonical 1).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// arrive on the barrier at shared::cta offset `bar` (in BOTH CTAs of the pair when CG == 2) once every
// tcgen05.mma issued so far by this thread has retired
// cta_mask (CG == 2): the CTAs of the cluster whose barrier at this offset gets the arrival (default: CTAs 0 and 1)
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar, uint16_t cta_mask = 3) {
  if constexpr (CG == 2) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(cta_mask)
        : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
}

#define TMEM_LD32(r, taddr)                                                                                        \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18," \
      "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                               \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
      : "r"(taddr)                                                                                                 \
      : "memory")

// tcgen05.wait::ld that also "touches" the 32 destination registers so that no use of them can be
// scheduled above the wait
#define TMEM_WAIT_LD32(r)                                                                                          \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                    \
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),   \
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),          \
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),        \
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),        \
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])                                                             \
               :                                                                                                   \
               : "memory")


// 16 consecutive 32-bit columns of this thread's TMEM lane
#define TMEM_LD16(r, taddr)                                                                                        \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"      \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                    \
      : "r"(taddr)                                                                                                 \
      : "memory")
#define TMEM_WAIT_LD16(r)                                                                                          \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                    \
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),   \
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),          \
                 "+r"(r[15])                                                                                       \
               :                                                                                                   \
               : "memory")

// chunk_flags for a 16-code slice: maximum, running maximum, 16-bit flag mask (bit j <=> s[j] >= m - margin)
__device__ __forceinline__ uint32_t chunk_flags16(const uint32_t (&r)[16], float margin, float& m, float& cm_out) {
  auto f = [&](int i) { return __uint_as_float(r[i]); };
  auto max3 = [](float x, float y, float z) { return fmaxf(fmaxf(x, y), z); };
  const float a0 = max3(f(0), f(1), f(2)), a1 = max3(f(3), f(4), f(5)), a2 = max3(f(6), f(7), f(8)),
              a3 = max3(f(9), f(10), f(11)), a4 = max3(f(12), f(13), f(14));
  const float cm = fmaxf(max3(a0, a1, a2), max3(a3, a4, f(15)));
  m = fmaxf(m, cm);
  cm_out = cm;
  const float thr = m - margin;
  uint32_t neg[2] = {0u, 0u};
#pragma unroll
  for (int j = 7; j >= 0; --j)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float d = __fsub_rn(__uint_as_float(r[c * 8 + j]), thr);
      neg[c] = __funnelshift_l(__float_as_uint(d), neg[c], 1);
    }
  return ~(neg[0] | (neg[1] << 8)) & 0xFFFFu;
}

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// One 32-code chunk of one row: chunk maximum, running maximum, flag mask of the scores within
// `margin` of the running maximum (which already includes this chunk).  Flags: d = s - thr on the FMA
// pipe, sign bits collected with funnel shifts (bit j of the result <=> s[j] >= thr).
__device__ __forceinline__ uint32_t chunk_flags(const uint32_t (&r)[32], float margin, float& m, float& cm_out) {
  // 32 -> 1 with three-input maxima (FMNMX3): 11 + 4 + 2 instructions
  auto f = [&](int i) { return __uint_as_float(r[i]); };
  auto max3 = [](float x, float y, float z) { return fmaxf(fmaxf(x, y), z); };
  float a[11];
#pragma unroll
  for (int j = 0; j < 10; ++j) a[j] = max3(f(3 * j), f(3 * j + 1), f(3 * j + 2));
  a[10] = fmaxf(f(30), f(31));
  const float b0 = max3(a[0], a[1], a[2]), b1 = max3(a[3], a[4], a[5]), b2 = max3(a[6], a[7], a[8]),
              b3 = fmaxf(a[9], a[10]);
  const float cm = fmaxf(max3(b0, b1, b2), b3);
  m = fmaxf(m, cm);
  cm_out = cm;
  const float thr = m - margin;
  uint32_t neg[4] = {0u, 0u, 0u, 0u};   // four independent chains of 8 sign bits
#pragma unroll
  for (int j = 7; j >= 0; --j)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float d = __fsub_rn(__uint_as_float(r[c * 8 + j]), thr);
      neg[c] = __funnelshift_l(__float_as_uint(d), neg[c], 1);   // (neg << 1) | sign(d)
    }
  const uint32_t below = neg[0] | (neg[1] << 8) | (neg[2] << 16) | (neg[3] << 24);
  return ~below;
}


// tcgen05.st: 16 consecutive 32-bit columns of this thread's TMEM lane (32x32b shape: lane = 32 * (warp % 4) + laneid)
#define TMEM_ST16(taddr, r)                                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),   \
               "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])           \
               : "memory")
#define TMEM_ST32(taddr, r)                                                                                          \
  asm volatile(                                                                                                      \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"  \
      "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                      \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),   \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),     \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),     \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                 \
      : "memory")
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace tc
}  // namespace dcvic
