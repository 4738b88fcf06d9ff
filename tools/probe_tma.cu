// Probe (measurement aid): TMA load throughput / latency from L2 as seen by one SM's ring of S stages.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_tma tools/probe_tma.cu -lcuda && tools/probe_tma
// Every CTA (one per SM) streams the SAME 512 KB FP16 "codebook" (L2 resident after the first pass) through a ring of
// S stages of BYTES each: one thread issues cp.async.bulk.tensor.2d, one thread waits and frees.  Reports bytes per
// clock per SM and the average time from issue to arrival.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra W_DONE;\n\tbra W_LOOP;\n\tW_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

// box = 64 fp16 columns x ROWS rows
template <int ROWS>
__global__ void __launch_bounds__(64, 1) ring_kernel(const __grid_constant__ CUtensorMap tm, int S, int iters, int krows,
                                                     unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) unsigned long long bars[64];
  __shared__ long long t_issue[32];
  constexpr int BYTES = ROWS * 128;
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * S; ++i) mbar_init(b0 + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  long long lat = 0;
  if (warp == 0 && lane == 0) {
    int y = (blockIdx.x * 64) % krows, x = 0;
    for (int i = 0; i < iters; ++i) {
      const int st = i % S;
      mbar_wait(b0 + (S + st) * 8, ((i / S) & 1) ^ 1);
      t_issue[st] = clock64();
      mbar_expect(b0 + st * 8, BYTES);
      tma_load(smem_u32(smem) + st * BYTES, &tm, x, y, b0 + st * 8);
      y += ROWS; if (y >= krows) { y = 0; x = (x + 64) % 256; }
    }
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      const int st = i % S;
      mbar_wait(b0 + st * 8, (i / S) & 1);
      lat += clock64() - t_issue[st];
      mbar_arrive(b0 + (S + st) * 8);
    }
    out[blockIdx.x * 2] = clock64() - t0;
    out[blockIdx.x * 2 + 1] = lat / iters;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS>
void run(EncodeFn enc, void* buf, int krows, int S, int grid, unsigned long long* dT) {
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {256, (cuuint64_t)krows};
  const cuuint64_t gstr[1] = {512};
  const cuuint32_t box[2] = {64, ROWS};
  const cuuint32_t es[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); exit(1); }
  const int smem = S * ROWS * 128 + 1024, iters = 4000;
  CK(cudaFuncSetAttribute(ring_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) { ring_kernel<ROWS><<<grid, 64, smem>>>(tm, S, iters, krows, dT); CK(cudaDeviceSynchronize()); }
  unsigned long long h[296];
  CK(cudaMemcpy(h, dT, sizeof(h), cudaMemcpyDeviceToHost));
  double tot = 0, lat = 0;
  for (int b = 0; b < grid; ++b) { tot += (double)iters * ROWS * 128 / (double)h[2 * b]; lat += (double)h[2 * b + 1]; }
  printf("grid %3d  stage %5d B x %2d stages (%3d KB in flight): %.1f B/clk/SM  (%.0f B/clk chip, %.2f TB/s at 1.965 GHz), issue->arrival %.0f cycles\n",
         grid, ROWS * 128, S, S * ROWS * 128 / 1024, tot / grid, tot, tot * 1.965e9 / 1e12, lat / grid);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  const int krows = 1024;                 // 1024 x 256 fp16 = 512 KB: L2 resident, shared by every SM
  void* buf;
  CK(cudaMalloc(&buf, (size_t)krows * 512));
  CK(cudaMemset(buf, 0, (size_t)krows * 512));
  unsigned long long* dT;
  CK(cudaMalloc(&dT, 296 * 8));
  for (int grid : {1, 148}) {
    for (int S : {2, 4, 8, 16}) run<64>(enc, buf, krows, S, grid, dT);
    for (int S : {2, 4, 8}) run<128>(enc, buf, krows, S, grid, dT);
    for (int S : {2, 4}) run<256>(enc, buf, krows, S, grid, dT);
  }
  // a big buffer (1 GB > L2): the DRAM-bound rate through the same ring
  const int big = 2 * 1024 * 1024;
  void* buf2;
  CK(cudaMalloc(&buf2, (size_t)big * 512));
  CK(cudaMemset(buf2, 0, (size_t)big * 512));
  for (int S : {4, 8, 16}) run<64>(enc, buf2, big, S, 148, dT);
  for (int S : {4, 6}) run<256>(enc, buf2, big, S, 148, dT);
  return 0;
}
