// Read-bandwidth probe for the access pattern of the tcgen05 search's operand producer
// (tools only; not part of the library).   nvcc -arch=sm_100a -O3 -o tools/bw_probe tools/bw_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__global__ void linear_read(const float4* __restrict__ p, size_t n4, float* out) {
  float acc = 0.f;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = ldg_stream(p + i), b = ldg_stream(p + i + stride), c = ldg_stream(p + i + 2 * stride), d = ldg_stream(p + i + 3 * stride);
    acc += a.x + b.y + c.z + d.w;
  }
  if (acc == 123.456f) *out = acc;
}
// tile pattern: CTA b handles token tiles b, b+grid, ... of 128 tokens x 256 channels (NCHW, HW=1024)
template <int NW, int SETS>
__global__ void __launch_bounds__(NW * 32) tile_read(const float* __restrict__ z, int N, int D, int HW, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tq = lane & 7, cg = lane >> 3;
  const int tokgrp = warp & 3, part = warp >> 2;           // NW/4 channel parts
  constexpr int PARTS = NW / 4;
  const int ntiles = N / 128;
  float acc = 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long t = (long long)tile * 128 + tokgrp * 32 + tq * 4;
    const float* zc = z + (size_t)(t / HW) * D * HW + (t % HW);
    // channels: groups of 8; this thread takes groups cg + 4*part + 4*PARTS*s
    float4 v[SETS][8];
    const int nsteps = D / (32 * PARTS);
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += SETS) {
#pragma unroll
      for (int u = 0; u < SETS; ++u) {
        const int g = cg + 4 * part + 4 * PARTS * (s0 + u);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[u][k] = ldg_stream(reinterpret_cast<const float4*>(zc + (size_t)(g * 8 + k) * HW));
      }
#pragma unroll
      for (int u = 0; u < SETS; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[u][k].x + v[u][k].w;
    }
  }
  if (acc == 123.456f) *out = acc;
}
// tile pattern + L2 prefetch of the tile `dist` iterations ahead (MODE 1: cp.async.bulk.prefetch.L2 512 B per
// channel row, MODE 2: prefetch.global.L2 per 128 B line)
template <int NW, int MODE>
__global__ void __launch_bounds__(NW * 32) tile_read_pf(const float* __restrict__ z, int N, int D, int HW, int dist, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = lane & 3, tq = lane >> 2;
  const int tokgrp = warp & 3, part = warp >> 2;
  constexpr int PARTS = NW / 4;
  const int ntiles = N / 128;
  float acc = 0.f;
  auto pf = [&](int tile) {
    if (tile >= ntiles) return;
    const long long t0 = (long long)tile * 128;
    const float* zb = z + (size_t)(t0 / HW) * D * HW + (t0 % HW);
    for (int c = threadIdx.x; c < D; c += NW * 32) {
      const float* pp = zb + (size_t)c * HW;
      if (MODE == 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pp), "r"(512u) : "memory");
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + 32 * j) : "memory");
      }
    }
  };
  for (int d = 0; d < dist; ++d) pf(blockIdx.x + d * gridDim.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    pf(tile + dist * gridDim.x);
    const long long t = (long long)tile * 128 + tokgrp * 32 + tq * 4;
    const float* zc = z + (size_t)(t / HW) * D * HW + (t % HW);
    float4 v[2][8];
    const int nsteps = D / (32 * PARTS);
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int g = cg + 4 * part + 4 * PARTS * (s0 + u);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[u][k] = ldg_stream(reinterpret_cast<const float4*>(zc + (size_t)(g * 8 + k) * HW));
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[u][k].x + v[u][k].w;
    }
  }
  if (acc == 123.456f) *out = acc;
}
// linear, U float4 loads in flight per thread, then consume (same issue structure as tile_read)
template <int U>
__global__ void linear_read_u(const float4* __restrict__ p, size_t n4, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldg_stream(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].w;
  }
  if (acc == 123.456f) *out = acc;
}
// finish-kernel staging pattern: CTA of 256 threads reads (and optionally writes back) TT tokens x 256 channels,
// i.e. 256 pieces of TT*4 bytes at 4 KB stride
template <int TT, bool WRITE>
__global__ void __launch_bounds__(256) fin_pattern(const float* __restrict__ z, float* __restrict__ o, int N, int D, int HW, float* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int TQ = TT / 4, CQ = 32 / TQ;
  const int tq = lane / CQ, cq = lane % CQ;
  const long long t = (long long)blockIdx.x * TT + 4 * tq;
  const size_t base = (size_t)(t / HW) * D * HW + (t % HW);
  float acc = 0.f;
  float4 v[2][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = h * 128 + wid * (4 * CQ) + cq * 4 + k;
      v[h][k] = ldg_stream(reinterpret_cast<const float4*>(z + base + (size_t)c * HW));
    }
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = h * 128 + wid * (4 * CQ) + cq * 4 + k;
      if (WRITE) {
        float4 w = v[h][k]; w.x += 1.f;
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + base + (size_t)c * HW), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
      } else acc += v[h][k].x + v[h][k].w;
    }
  if (!WRITE && acc == 123.456f) *out = acc;
}
template <class F> float timeit(F f, int iters = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a); for (int i = 0; i < iters; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / iters;
}
int main() {
  const int B = 64 * 4, D = 256, HW = 1024, N = B * HW;   // 268 MB (> L2)
  const size_t n = (size_t)N * D;
  float *z, *out; cudaMalloc(&z, n * 4); cudaMalloc(&out, 4); cudaMemset(z, 0, n * 4);
  const double gb = n * 4 / 1e9;
  float ms = timeit([&] { linear_read<<<148 * 8, 256>>>((const float4*)z, n / 4, out); });
  printf("linear 148x8x256            : %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { linear_read<<<148 * 2, 512>>>((const float4*)z, n / 4, out); });
  printf("linear 148x2x512            : %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { tile_read<4, 2><<<148, 128>>>(z, N, D, HW, out); });
  printf("tile 4 warps, 2 sets (32KB) : %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { tile_read<8, 2><<<148, 256>>>(z, N, D, HW, out); });
  printf("tile 8 warps, 2 sets (64KB) : %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { tile_read<8, 4><<<148, 256>>>(z, N, D, HW, out); });
  printf("tile 8 warps, 4 sets (128KB): %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { tile_read<16, 2><<<148, 512>>>(z, N, D, HW, out); });
  printf("tile 16 warps, 2 sets(128KB): %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { tile_read<16, 4><<<148, 512>>>(z, N, D, HW, out); });
  printf("tile 16 warps, 4 sets(256KB): %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { tile_read<8, 2><<<296, 256>>>(z, N, D, HW, out); });
  printf("tile 8 warps, 2 sets, 2 CTA/SM: %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { linear_read_u<16><<<148, 128>>>((const float4*)z, n / 4, out); });
  printf("linear 148x128 thr, 16 in flight (32KB/SM): %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { linear_read_u<16><<<148, 256>>>((const float4*)z, n / 4, out); });
  printf("linear 148x256 thr, 16 in flight (64KB/SM): %.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { linear_read_u<16><<<148, 512>>>((const float4*)z, n / 4, out); });
  printf("linear 148x512 thr, 16 in flight (128KB/SM): %.1f GB/s\n", gb / ms * 1e3);
  {  // L2-resident footprint (32 MB): same kernels
    const int Ns = 32 * HW; const size_t ns = (size_t)Ns * D; const double gbs = ns * 4 / 1e9;
    ms = timeit([&] { linear_read_u<16><<<148, 128>>>((const float4*)z, ns / 4, out); }, 50);
    printf("L2-resident linear 128 thr 16 in flight: %.1f GB/s\n", gbs / ms * 1e3);
    ms = timeit([&] { linear_read_u<16><<<148, 256>>>((const float4*)z, ns / 4, out); }, 50);
    printf("L2-resident linear 256 thr 16 in flight: %.1f GB/s\n", gbs / ms * 1e3);
    ms = timeit([&] { tile_read<4, 2><<<148, 128>>>(z, Ns, D, HW, out); }, 50);
    printf("L2-resident tile 4 warps 2 sets: %.1f GB/s\n", gbs / ms * 1e3);
    ms = timeit([&] { tile_read<8, 2><<<148, 256>>>(z, Ns, D, HW, out); }, 50);
    printf("L2-resident tile 8 warps 2 sets: %.1f GB/s\n", gbs / ms * 1e3);
  }
  {
    float* o; cudaMalloc(&o, n * 4);
    ms = timeit([&] { fin_pattern<32, false><<<N / 32, 256>>>(z, o, N, D, HW, out); });
    printf("finish pattern read  32-token tiles (128 B pieces): %.1f GB/s\n", gb / ms * 1e3);
    ms = timeit([&] { fin_pattern<16, false><<<N / 16, 256>>>(z, o, N, D, HW, out); });
    printf("finish pattern read  16-token tiles ( 64 B pieces): %.1f GB/s\n", gb / ms * 1e3);
    ms = timeit([&] { fin_pattern<32, true><<<N / 32, 256>>>(z, o, N, D, HW, out); });
    printf("finish pattern read+write 32-token tiles: %.1f GB/s (r+w bytes)\n", 2 * gb / ms * 1e3);
    cudaFree(o);
  }
  for (int dist = 1; dist <= 0; ++dist) {
    ms = timeit([&] { tile_read_pf<4, 1><<<148, 128>>>(z, N, D, HW, dist, out); });
    printf("tile 4 warps 2 sets + bulk prefetch dist %d : %.1f GB/s\n", dist, gb / ms * 1e3);
    ms = timeit([&] { tile_read_pf<4, 2><<<148, 128>>>(z, N, D, HW, dist, out); });
    printf("tile 4 warps 2 sets + line prefetch dist %d : %.1f GB/s\n", dist, gb / ms * 1e3);
    ms = timeit([&] { tile_read_pf<8, 1><<<148, 256>>>(z, N, D, HW, dist, out); });
    printf("tile 8 warps 2 sets + bulk prefetch dist %d : %.1f GB/s\n", dist, gb / ms * 1e3);
    ms = timeit([&] { tile_read_pf<8, 2><<<148, 256>>>(z, N, D, HW, dist, out); });
    printf("tile 8 warps 2 sets + line prefetch dist %d : %.1f GB/s\n", dist, gb / ms * 1e3);
  }
  cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
