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
  cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
