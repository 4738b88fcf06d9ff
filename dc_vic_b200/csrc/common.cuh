// Shared device/host helpers for libdcvic_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/dcvic_b200.h"

#define DCVIC_CHECK_ARG(cond) \
  do {                        \
    if (!(cond)) return DCVIC_ERR_BAD_ARG; \
  } while (0)

static inline int dcvic_launch_status() {
  return cudaGetLastError() == cudaSuccess ? DCVIC_OK : DCVIC_ERR_CUDA;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

namespace dcvic {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum in double; result valid in thread 0.  `scratch` >= 32 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? scratch[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// "Last block finishes": every block publishes one double partial; the block that arrives
// last reduces them in a fixed order (deterministic) and returns true in ALL its threads
// with the total in *total (thread 0 only).  `counter` must be 0 on entry and is reset.
__device__ __forceinline__ bool publish_and_elect_last(double block_partial, double* partials, unsigned* counter,
                                                       unsigned nblocks, unsigned block_linear, double* scratch,
                                                       double* total) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    partials[block_linear] = block_partial;
    __threadfence();
    unsigned prev = atomicAdd(counter, 1u);
    s_last = (prev == nblocks - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double s = 0.0;
  for (unsigned i = threadIdx.x; i < nblocks; i += blockDim.x) s += __ldcg(partials + i);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    *total = s;
    *counter = 0u;
  }
  return true;
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running; pdl_wait() blocks until that predecessor has completed
// and its writes are visible, pdl_launch_dependents() lets the successor start early.  Everything a kernel reads
// from its predecessor's outputs must come after pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace dcvic
