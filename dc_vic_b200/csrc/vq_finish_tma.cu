// Finish kernel of the wide VQ path, TMA-pipelined version: FP32 re-rank of the candidates the tensor search
// flagged, codebook gather, straight-through value z + (e - z), loss partials, indices.
//
// Replaces (for H*W % 32 == 0 and e_dim in {64, 128, 192, 256}) vq_finish_v5_kernel, whose CTAs walk through
// "load tile -> lists -> codebook rows -> store tile" one dependent round trip after the other, so that on average
// only a quarter of the resident CTAs have loads in flight (2.5 TB/s on C2).  Here the z / z_q traffic is
// decoupled from the arithmetic:
//   * one persistent CTA per SM; a ring of NST shared-memory stages, each one 32-token tile [e_dim][32] FP32
//     exactly as it lies in the NCHW tensor (32 consecutive tokens of one image x all channels);
//   * a producer thread keeps the ring full with TMA loads (cp.async.bulk.tensor.2d, SWIZZLE_128B) and drains it
//     with TMA stores of the same stages after the consumers turned z into z_q in place - no registers are
//     the landing zone of any global load of z, so 5 tiles (160 KB) per SM can be in flight;
//   * 15 consumer warps take (tile, token quad) units round-robin.  A warp owns 4 tokens x 8 channel lanes
//     (lane = 8*token + channel lane, channel c = lane + 8i): with the 128-byte swizzle the 32 lanes of every
//     shared-memory access fall on 32 distinct banks.  Per unit a warp reads its tokens' meta record and list
//     entries (prefetched one unit ahead), expands them into candidate codes, requests the first candidate's
//     codebook row BEFORE the tile has arrived, re-ranks in FP32 where more than one code was flagged, and writes
//     z + (e - z) over z.
// Arithmetic (reference: taming/modules/vqvae/quantize.py:268-281 / :56-80): d = (|z|^2 + |e|^2) - 2 z.e in FP32,
// minimum over (d, index); z_q = z + (e - z); loss partial = sum (e - z)^2.
#include <cuda.h>
#include <float.h>
#include <stdlib.h>

#include "vq_common.cuh"

namespace dcvic {
namespace {

constexpr int FT = 32;                       // tokens per tile
#ifndef DCVIC_FIN_NCW
#define DCVIC_FIN_NCW 12
#endif
#ifndef DCVIC_FIN_NEW
#define DCVIC_FIN_NEW 3
#endif
constexpr int F_NCW = DCVIC_FIN_NCW;         // consumer warps
constexpr int F_NEW = DCVIC_FIN_NEW;         // list-expander warps
constexpr int F_THREADS = (F_NCW + F_NEW + 1) * 32;  // + the producer warp (16 warps: 128 registers per thread)

__device__ __forceinline__ uint32_t f_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void f_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void f_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "F_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra F_WAIT_DONE;\n\t"
      "bra F_WAIT_LOOP;\n\t"
      "F_WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void f_tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void f_tma_store_2d(const CUtensorMap* map, int x, int y, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x),
               "r"(y), "r"(src)
               : "memory");
}
__device__ __forceinline__ void f_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void f_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void f_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void f_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float4 f_lds4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void f_sts4(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#ifdef DCVIC_TRACE
__device__ unsigned long long g_trace_ftma[148][32][8];
#define FTM_DECL unsigned long long ftm_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ftm_t = clock64(), ftm_t0 = ftm_t
#define FTM_MARK(i)                            \
  do {                                         \
    const unsigned long long now = clock64();  \
    ftm_acc[i] += now - ftm_t;                 \
    ftm_t = now;                               \
  } while (0)
#define FTM_PUT(w)                                                                           \
  do {                                                                                       \
    if (lane == 0 && blockIdx.x < 148) {                                                     \
      ftm_acc[7] = clock64() - ftm_t0;                                                       \
      for (int i_ = 0; i_ < 8; ++i_) g_trace_ftma[blockIdx.x][w][i_] = ftm_acc[i_];          \
    }                                                                                        \
  } while (0)
#else
#define FTM_DECL
#define FTM_MARK(i)
#define FTM_PUT(w)
#endif

template <int D, int NST>
__global__ void __launch_bounds__(F_THREADS, 1)
vq_finish_tma_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_zq,
                     const float* __restrict__ E, const float* __restrict__ ee, const float* __restrict__ emax_ptr,
                     const int* __restrict__ cand, const VqMeta* __restrict__ meta, const uint2* __restrict__ list,
                     int N, int HW, int K, int contig, int64_t* __restrict__ idx, double* __restrict__ partials,
                     unsigned* __restrict__ counters) {
  constexpr int STAGE_BYTES = D * 128;
  extern __shared__ uint8_t f_smem_raw[];
  __shared__ __align__(8) unsigned long long s_bar[2 * NST];
  __shared__ __align__(8) unsigned long long s_cbar[NST];
  __shared__ unsigned short s_ck[NST][FT][kCandMax];   // candidate codes of the tile in each stage
  __shared__ int s_nc[NST][FT];                // flagged codes per token, -1: scan the whole codebook
  __shared__ unsigned s_stat[4];
  // SWIZZLE_128B stages need 1024-byte alignment
  const uint32_t stage0 = (f_smem_u32(f_smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = f_smem_u32(s_bar);
  auto full_bar = [&](int st) { return bar0 + st * 8; };
  auto done_bar = [&](int st) { return bar0 + (NST + st) * 8; };
  const uint32_t cbar0 = f_smem_u32(s_cbar);
  auto cand_bar = [&](int st) { return cbar0 + st * 8; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = N / FT;
  // tiles of this CTA: a contiguous range (its loads in flight then cover neighbouring 128-byte pieces of the same
  // channel rows), or strided by the grid
  int nloc, tile_first, tile_step;
  if (contig) {
    const int base = ntiles / (int)gridDim.x, rem = ntiles % (int)gridDim.x, b = (int)blockIdx.x;
    nloc = base + (b < rem ? 1 : 0);
    tile_first = b * base + min(b, rem);
    tile_step = 1;
  } else {
    nloc = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    tile_first = (int)blockIdx.x;
    tile_step = (int)gridDim.x;
  }

  if (threadIdx.x == 0) {
    for (int st = 0; st < NST; ++st) {
      f_mbar_init(full_bar(st), 1);
      f_mbar_init(done_bar(st), 8);
      f_mbar_init(cand_bar(st), 1);
    }
    s_stat[0] = 0; s_stat[1] = 0; s_stat[2] = 0; s_stat[3] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == F_NCW + F_NEW) {
    // ------------------------------------------------------------ producer: TMA loads of z, TMA stores of z_q
    // z is an input of the whole call: loads start before the tensor search (the predecessor under programmatic
    // dependent launch) has finished.
    if (lane == 0) {
      auto coords = [&](int j, int& x, int& y) {
        const int t0 = (tile_first + j * tile_step) * FT;
        x = t0 % HW;
        y = (t0 / HW) * D;
      };
      auto load = [&](int j) {
        const int st = j % NST;
        int x, y;
        coords(j, x, y);
        f_mbar_arrive_expect_tx(full_bar(st), STAGE_BYTES);
        f_tma_load_2d(stage0 + st * STAGE_BYTES, &tm_z, x, y, full_bar(st));
      };
      FTM_DECL;
      int nissued = 0;
      for (; nissued < nloc && nissued < NST; ++nissued) load(nissued);
      FTM_MARK(0);
      for (int j = 0; j < nloc; ++j) {
        const int st = j % NST;
        f_mbar_wait(done_bar(st), (j / NST) & 1);
        FTM_MARK(1);
        int x, y;
        coords(j, x, y);
        f_tma_store_2d(&tm_zq, x, y, stage0 + st * STAGE_BYTES);
        f_bulk_commit();
        // the stage stored one iteration ago is the next one to refill, once the TMA has read it out
        if (j >= 1 && nissued < nloc) {
          f_bulk_wait_read<1>();
          FTM_MARK(2);
          load(nissued);
          ++nissued;
        }
        FTM_MARK(3);
      }
      f_bulk_wait_all();
      FTM_MARK(4);
      FTM_PUT(F_NCW + F_NEW);
    }
    return;
  }

  if (warp >= F_NCW) {
    // ------------------------------------------------------------ list expanders: one lane per token
    // The tiles of stage st are expanded by warp (st mod F_NEW), each while its z is still on the way: meta record
    // and the first 4 or 8 entries of both lists (32-byte sectors, written whole by the search; both prefetched),
    // the rare longer tails, -> candidate codes s_ck[stage][token][], count s_nc[stage][token].
    // (One warp per STAGE, not per tile: the parity wait on the stage's done barrier below is only valid if this
    // warp has itself seen the previous phase complete, i.e. expanded the stage's previous tile.)
    const int e = warp - F_NCW;
    FTM_DECL;
    pdl_wait();                                          // meta / lists / emax come from the preceding kernels
    FTM_MARK(0);
    const float emax = cand ? 0.f : __ldg(emax_ptr);
    auto tile_token = [&](int j) { return (tile_first + j * tile_step) * FT + lane; };
    // software pipeline over this warp's tiles: meta records two tiles ahead, list entries one tile ahead
    auto load_meta = [&](int j, VqMeta& m, int& ck) {
      if (j >= nloc) return;
      if (cand) ck = __ldg(cand + tile_token(j));
      else m = meta[tile_token(j)];
    };
    auto load_entries = [&](int j, const VqMeta& m, uint4 (&en)[2][4]) {   // entries 0-7 of both lists
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) en[q][i] = make_uint4(0u, 0u, 0u, 0u);
      if (j >= nloc || cand) return;
      const int t = tile_token(j);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4* lp = reinterpret_cast<const uint4*>(list + ((size_t)t * 2 + q) * kListCap);
        const int n = q == 0 ? m.n0 : m.n1;
        if (n > 0) { en[q][0] = __ldg(lp); en[q][1] = __ldg(lp + 1); }        // one 32-byte sector
        if (n > 4) { en[q][2] = __ldg(lp + 2); en[q][3] = __ldg(lp + 3); }    // (a quarter of the lists)
      }
    };
    auto next_tile = [&](int j) {                          // this warp's next tile after j (or >= nloc)
      do { ++j; } while (j < nloc && (j % NST) % F_NEW != e);
      return j;
    };
    VqMeta mt = {}, mt_n = {};
    int cand_k = 0, cand_n = 0;
    uint4 en[2][4], en_n[2][4];
    const int j0 = next_tile(-1);
    int j1 = next_tile(j0);
    load_meta(j0, mt, cand_k);
    load_entries(j0, mt, en);
    load_meta(j1, mt_n, cand_n);
    for (int j = j0; j < nloc;) {
      const int st = j % NST;
      const int j2 = next_tile(j1);
      const int t = tile_token(j);
      int w = 0;                                          // candidates found
      // overflowed list / FP16-unsafe token or codebook
      bool full = !cand && (mt.n0 < 0 || mt.n1 < 0 || mt.nb0 < 0 || mt.nb1 < 0);
      // next tile's entries (its meta record arrived during the previous tile), the meta record after that
      VqMeta mt_nn = {};
      int cand_nn = 0;
      load_entries(j1, mt_n, en_n);
      load_meta(j2, mt_nn, cand_nn);
      // the stage's previous tile must be through with s_ck / s_nc
      if (j >= NST) f_mbar_wait(done_bar(st), ((j / NST) - 1) & 1);
      FTM_MARK(1);
      unsigned short* ck = &s_ck[st][lane][0];
      if (cand) {
        ck[0] = (unsigned short)min(max(cand_k, 0), K - 1);
        w = 1;
      } else if (!full) {
        const float thr = fmaxf(fmaxf(mt.m0, mt.m1), fmaxf(mt.mb0, mt.mb1)) - vq_margin(mt.zz, emax);
        auto take = [&](unsigned key, unsigned mask) {
          if (vq_key_upper(key) < thr) return;               // chunk maximum (rounded towards +inf) below the threshold
          const int c0 = (int)(key & 0x7Fu) * kChunk;
          while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            if (w < kCandMax) ck[w] = (unsigned short)(c0 + b);
            ++w;
          }
        };
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int n = q == 0 ? mt.n0 : mt.n1;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (n > 2 * i) take(en[q][i].x, en[q][i].y);
            if (n > 2 * i + 1) take(en[q][i].z, en[q][i].w);
          }
          for (int i = 8; i < n; ++i) {                   // (rare; never on split tiles: a half holds 8 entries)
            const uint2 en2 = __ldg(list + ((size_t)t * 2 + q) * kListCap + i);
            take(en2.x, en2.y);
          }
        }
        if (mt.split) {
          // second half of a split tile (one tile in eight on C2): entries 8.. of both lists, not prefetched
          uint4 eb[2][4];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint4* lp = reinterpret_cast<const uint4*>(list + ((size_t)t * 2 + q) * kListCap + kListCap / 2);
            const int n = q == 0 ? mt.nb0 : mt.nb1;
#pragma unroll
            for (int i = 0; i < 4; ++i) eb[q][i] = make_uint4(0u, 0u, 0u, 0u);
            if (n > 0) { eb[q][0] = __ldg(lp); eb[q][1] = __ldg(lp + 1); }
            if (n > 4) { eb[q][2] = __ldg(lp + 2); eb[q][3] = __ldg(lp + 3); }
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int n = q == 0 ? mt.nb0 : mt.nb1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (n > 2 * i) take(eb[q][i].x, eb[q][i].y);
              if (n > 2 * i + 1) take(eb[q][i].z, eb[q][i].w);
            }
          }
        }
        if (w > kCandMax || w <= 0) full = true;
      }
      if (full) ck[0] = 0;
      s_nc[st][lane] = full ? -1 : w;
      mt = mt_n; mt_n = mt_nn;
      cand_k = cand_n; cand_n = cand_nn;
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) en[q][i] = en_n[q][i];
      __syncwarp();
      if (lane == 0) f_mbar_arrive(cand_bar(st));
      FTM_MARK(2);
      j = j1;
      j1 = j2;
    }
    FTM_PUT(warp);
    return;
  }

  // -------------------------------------------------------------- consumers
  // Every lane holds 4 consecutive channels per 128-channel block (c = 4*lane + 128h) of all 4 tokens of a quad: one
  // LDG.128 per codebook row and block, one LDS.128 / STS.128 per channel (the four tokens are the four words of a
  // 16-byte piece of the swizzled stage).  Lane l visits its 4 channels in the order (j + (l >> 1)) & 3, j = 0..3, so
  // that the 8 lanes of a quarter warp touch 8 different pieces (rows 4l + jj: (row & 7) = 4(l & 1) + jj); codebook
  // values are rotated to that order where they meet z.  Everything per token (candidate count, codes, best
  // distance) is warp-uniform, so tokens that need no re-rank cost nothing in the re-rank loop.  Units (tile, quad)
  // are handed out in order through a shared counter: a warp that drew a long re-rank does not hold up its CTA.
  constexpr int NH = (D + 127) / 128;                    // 128-channel blocks
  const int cw = warp;
  const int rot = (lane >> 1) & 3;
  const bool hv[2] = {4 * lane < D, 4 * lane + 128 < D}; // which blocks this lane has (e_dim 64 / 192: not all)
  FTM_DECL;
  pdl_wait();                                            // ee comes from the prepare kernel
  FTM_MARK(0);
  const int total_units = nloc * 8;
  double dsq = 0.0;
  unsigned n_rr = 0, n_fs = 0;
  auto unit_token0 = [&](int u) { return (tile_first + (u >> 3) * tile_step) * FT + 4 * (u & 7); };
  // out[j] = v[(j + r) & 3]
  auto rotl = [](float4 v, int r) {
    if (r & 1) v = make_float4(v.y, v.z, v.w, v.x);
    if (r & 2) v = make_float4(v.z, v.w, v.x, v.y);
    return v;
  };
  // codebook row k, this lane's channels (natural order: rotating here would wait for the data)
  auto load_row = [&](float4 (&r)[NH], int k) {
    const float* row = E + (size_t)k * D + 4 * lane;
#pragma unroll
    for (int h = 0; h < NH; ++h)
      r[h] = hv[h] ? __ldg(reinterpret_cast<const float4*>(row + 128 * h)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };

  for (;;) {
    int u = 0;
    if (lane == 0) u = (int)atomicAdd(&s_stat[3], 1u);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= total_units) break;
    const int j = u >> 3, q = u & 7, st = j % NST;
    const int t0 = unit_token0(u);
    // shared-memory offsets of this lane's 4 channel rows (visiting order), block 0; block h adds h * 16384
    uint32_t zo[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int row = 4 * lane + ((jj + rot) & 3);
      zo[jj] = stage0 + st * STAGE_BYTES + row * 128 + ((q ^ (row & 7)) << 4);
    }
#ifdef DCVIC_FIN_COPYONLY
    {
      f_mbar_wait(full_bar(st), (j / NST) & 1);
#pragma unroll
      for (int h = 0; h < NH; ++h)
        if (hv[h])
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float4 a = f_lds4(zo[jj] + h * 16384);
            a.x += 1.f; a.y += 1.f; a.z += 1.f; a.w += 1.f;
            f_sts4(zo[jj] + h * 16384, a);
          }
      f_fence_proxy_async();
      __syncwarp();
      if (lane == 0) f_mbar_arrive(done_bar(st));
      continue;
    }
#endif
    // ---- candidates of the quad's tokens (from the expander warps)
    f_mbar_wait(cand_bar(st), (j / NST) & 1);
    FTM_MARK(1);
    int nc[4], bk[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      nc[g] = s_nc[st][4 * q + g];
      bk[g] = s_ck[st][4 * q + g][0];
    }
    // ---- first candidates' codebook rows (the winners for three tokens out of four), requested ahead of the tile
    float4 er[4][NH];
#pragma unroll
    for (int g = 0; g < 4; ++g) load_row(er[g], bk[g]);

    // ---- the tile
    FTM_MARK(2);
    f_mbar_wait(full_bar(st), (j / NST) & 1);
    FTM_MARK(3);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (nc[g] == 1) continue;                           // warp-uniform
      // |z|^2 and z.e of token g: per-lane partials over its channels, then over the warp
      float4 zg[NH];
      float zz = 0.f;
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (hv[h]) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float4 a = f_lds4(zo[jj] + h * 16384);
            v[jj] = g == 0 ? a.x : g == 1 ? a.y : g == 2 ? a.z : a.w;
          }
        }
        zg[h] = rotl(make_float4(v[0], v[1], v[2], v[3]), (4 - rot) & 3);   // back to channel order
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) zz = __fadd_rn(zz, __fmul_rn(v[jj], v[jj]));
      }
      zz = warp_sum(zz);
      auto dot = [&](const float4 (&r)[NH]) {
        float dp = 0.f;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          dp = fmaf(zg[h].x, r[h].x, dp); dp = fmaf(zg[h].y, r[h].y, dp);
          dp = fmaf(zg[h].z, r[h].z, dp); dp = fmaf(zg[h].w, r[h].w, dp);
        }
        return warp_sum(dp);
      };
      float bd = FLT_MAX;
      int kb = 0x7fffffff;
      if (nc[g] > 1) {
        ++n_rr;
        bd = fmaf(-2.f, dot(er[g]), __fadd_rn(zz, __ldg(ee + bk[g])));
        kb = bk[g];
#pragma unroll 1
        for (int ci = 1; ci < nc[g]; ci += 2) {            // two candidate rows in flight
          const int k0 = s_ck[st][4 * q + g][ci], k1 = s_ck[st][4 * q + g][min(ci + 1, nc[g] - 1)];
          float4 e0[NH], e1[NH];
          load_row(e0, k0);
          load_row(e1, k1);
          const float d0 = fmaf(-2.f, dot(e0), __fadd_rn(zz, __ldg(ee + k0)));
          const float d1 = fmaf(-2.f, dot(e1), __fadd_rn(zz, __ldg(ee + k1)));
          if (d0 < bd || (d0 == bd && k0 < kb)) {
            bd = d0;
            kb = k0;
#pragma unroll
            for (int h = 0; h < NH; ++h) er[g][h] = e0[h];
          }
          if (d1 < bd || (d1 == bd && k1 < kb)) {           // (k1 == k0 on an odd tail: no effect)
            bd = d1;
            kb = k1;
#pragma unroll
            for (int h = 0; h < NH; ++h) er[g][h] = e1[h];
          }
        }
      } else {
        // whole-codebook scan (overflowed list or FP16-unsafe input; rare): four rows in flight
        ++n_fs;
#pragma unroll 1
        for (int k = 0; k < K; k += 4) {
          float4 e[4][NH];
          float d[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) load_row(e[r], min(k + r, K - 1));
#pragma unroll
          for (int r = 0; r < 4; ++r) d[r] = fmaf(-2.f, dot(e[r]), __fadd_rn(zz, __ldg(ee + min(k + r, K - 1))));
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int kr = min(k + r, K - 1);
            if (d[r] < bd || (d[r] == bd && kr < kb)) { bd = d[r]; kb = kr; }
          }
        }
        load_row(er[g], kb);
      }
      bk[g] = kb;
    }
    FTM_MARK(4);
    // ---- z_q = z + (e - z) in place, loss partial
    float sq = 0.f;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      if (!hv[h]) continue;
      const float4 v0 = rotl(er[0][h], rot), v1 = rotl(er[1][h], rot), v2 = rotl(er[2][h], rot),
                   v3 = rotl(er[3][h], rot);      // visiting order
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 a = f_lds4(zo[jj] + h * 16384);
        const float e0 = jj == 0 ? v0.x : jj == 1 ? v0.y : jj == 2 ? v0.z : v0.w;
        const float e1 = jj == 0 ? v1.x : jj == 1 ? v1.y : jj == 2 ? v1.z : v1.w;
        const float e2 = jj == 0 ? v2.x : jj == 1 ? v2.y : jj == 2 ? v2.z : v2.w;
        const float e3 = jj == 0 ? v3.x : jj == 1 ? v3.y : jj == 2 ? v3.z : v3.w;
        const float d0 = __fsub_rn(e0, a.x), d1 = __fsub_rn(e1, a.y), d2 = __fsub_rn(e2, a.z), d3 = __fsub_rn(e3, a.w);
        f_sts4(zo[jj] + h * 16384,
               make_float4(__fadd_rn(a.x, d0), __fadd_rn(a.y, d1), __fadd_rn(a.z, d2), __fadd_rn(a.w, d3)));
        sq = fmaf(d0, d0, sq); sq = fmaf(d1, d1, sq); sq = fmaf(d2, d2, sq); sq = fmaf(d3, d3, sq);
      }
    }
    dsq += (double)sq;
    if (lane < 4) idx[t0 + lane] = (int64_t)(lane == 0 ? bk[0] : lane == 1 ? bk[1] : lane == 2 ? bk[2] : bk[3]);
    f_fence_proxy_async();                      // generic-proxy writes of the stage -> visible to the TMA store
    __syncwarp();
    if (lane == 0) f_mbar_arrive(done_bar(st));
    FTM_MARK(5);
  }
  FTM_PUT(cw);
  // loss: one partial per consumer warp, summed in index order by vq_loss_finalize_kernel (deterministic)
  {
    const double wsum = warp_sum(dsq);
    if (lane == 0) partials[(size_t)blockIdx.x * F_NCW + cw] = wsum;
  }
  // statistics: one global atomic per CTA and counter, issued by the consumer warp that finishes last
  if (lane == 0) {
    if (n_rr) atomicAdd(&s_stat[0], n_rr);
    if (n_fs) atomicAdd(&s_stat[1], n_fs);
    __threadfence_block();
    if (atomicAdd(&s_stat[2], 1u) == F_NCW - 1) {
      __threadfence_block();
      const unsigned rr = atomicAdd(&s_stat[0], 0u), fs = atomicAdd(&s_stat[1], 0u);
      if (rr) atomicAdd(counters + kCtrRerank, rr);
      if (fs) atomicAdd(counters + kCtrOverflow, fs);
    }
  }
}

typedef CUresult (*FEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

FEncodeTiledFn f_encode_fn() {
  static FEncodeTiledFn cached = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return (FEncodeTiledFn) nullptr;
    }
    return reinterpret_cast<FEncodeTiledFn>(fn);
  }();
  return cached;
}

// NCHW FP32 tensor as a 2-D map: x = position in the image (H*W, contiguous), y = image * e_dim + channel
bool make_tile_map(CUtensorMap* tm, const float* base, int B, int D, int HW) {
  FEncodeTiledFn encode = f_encode_fn();
  if (!encode) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)B * D};
  const cuuint64_t gstride[1] = {(cuuint64_t)HW * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)FT, (cuuint32_t)D};
  const cuuint32_t estr[2] = {1, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int D>
int launch_finish_tma(const CUtensorMap& tz, const CUtensorMap& tq, const float* E, const float* ee,
                      const float* emax, const int* cand, const VqMeta* meta, const uint2* list, int N, int HW, int K,
                      float beta, int legacy, int64_t* idx, float* loss, double* partials, unsigned* counters,
                      cudaStream_t s) {
  constexpr int STAGE = D * 128;
#ifdef DCVIC_FIN_NST
  constexpr int NST = DCVIC_FIN_NST;
#else
  constexpr int NST = (192 * 1024 / STAGE) < 8 ? (192 * 1024 / STAGE) : 8;
#endif
  const int smem = NST * STAGE + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(vq_finish_tma_kernel<D, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return DCVIC_ERR_CUDA;
    attr_set = true;
  }
  const int ntiles = N / FT;
  const int grid = ntiles < kNumSMs ? ntiles : kNumSMs;
  static const int contig = (getenv("DCVIC_FIN_MAP") && atoi(getenv("DCVIC_FIN_MAP")) == 0) ? 0 : 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(F_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute pdl[1];
  pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = pdl;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, vq_finish_tma_kernel<D, NST>, tz, tq, E, ee, emax, cand, meta, list, N, HW, K, contig,
                         idx, partials, counters) != cudaSuccess)
    return DCVIC_ERR_CUDA;
  return vq_launch_loss_finalize(partials, grid * F_NCW, (long long)N * D, beta, legacy, loss, s);
}

}  // namespace

#ifdef DCVIC_TRACE
}  // namespace dcvic
extern "C" int dcvic_debug_read_ftma_trace(unsigned long long* host_out /* [148][32][8] */) {
  return cudaMemcpyFromSymbol(host_out, dcvic::g_trace_ftma, sizeof(dcvic::g_trace_ftma)) == cudaSuccess ? 0 : -4;
}
namespace dcvic {
#endif

bool vq_finish_tma_supported(const float* z, const float* zq, int D, int HW, int K) {
  static const bool disabled = getenv("DCVIC_FINISH_TMA") && atoi(getenv("DCVIC_FINISH_TMA")) == 0;
  if (disabled) return false;
  if (!(D == 64 || D == 128 || D == 192 || D == 256)) return false;
  if (HW % FT != 0 || K > 65535) return false;
  if ((reinterpret_cast<uintptr_t>(z) & 15) || (reinterpret_cast<uintptr_t>(zq) & 15)) return false;
  static const bool sm100 = [] {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return major == 10;
  }();
  return sm100 && f_encode_fn() != nullptr;
}

int vq_finish_tma(const float* z, const float* E, const float* ee, const float* emax, const int* cand,
                  const VqMeta* meta, const uint2* list, int B, int D, int HW, int K, float beta, int legacy,
                  float* zq, int64_t* idx, float* loss, double* partials, unsigned* counters, cudaStream_t s) {
  CUtensorMap tz, tq;
  if (!make_tile_map(&tz, z, B, D, HW) || !make_tile_map(&tq, zq, B, D, HW)) return DCVIC_ERR_CUDA;
  const int N = B * HW;
#define DCVIC_FIN_TMA(DD)                                                                                          \
  launch_finish_tma<DD>(tz, tq, E, ee, emax, cand, meta, list, N, HW, K, beta, legacy, idx, loss, partials, counters, s)
  switch (D) {
    case 64: return DCVIC_FIN_TMA(64);
    case 128: return DCVIC_FIN_TMA(128);
    case 192: return DCVIC_FIN_TMA(192);
    case 256: return DCVIC_FIN_TMA(256);
  }
#undef DCVIC_FIN_TMA
  return DCVIC_ERR_UNSUPPORTED;
}

}  // namespace dcvic
