// Shared device helpers for the cacfe kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cacfe {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}
// NaN-propagating forms (numpy's np.min / np.max, which predict_utils.normalize_data uses: a NaN sample makes the whole
// clip NaN; fminf / fmaxf would skip it).  Same FMNMX instruction with the .NAN modifier.  Used by the clip normalisation only.
__device__ __forceinline__ float min_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float warp_min_nan(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min_nan(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}
__device__ __forceinline__ float warp_max_nan(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max_nan(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// Block-wide (min, max) into lane 0 of warp 0.  `scratch` holds 2*32 floats.
__device__ __forceinline__ void block_minmax(float& mn, float& mx, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  mn = warp_min(mn);
  mx = warp_max(mx);
  if (lane == 0) {
    scratch[warp] = mn;
    scratch[32 + warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    mn = lane < nwarp ? scratch[lane] : INFINITY;
    mx = lane < nwarp ? scratch[32 + lane] : -INFINITY;
    mn = warp_min(mn);
    mx = warp_max(mx);
  }
  __syncthreads();
}

__device__ __forceinline__ void block_minmax_nan(float& mn, float& mx, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  mn = warp_min_nan(mn);
  mx = warp_max_nan(mx);
  if (lane == 0) {
    scratch[warp] = mn;
    scratch[32 + warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    mn = lane < nwarp ? scratch[lane] : INFINITY;
    mx = lane < nwarp ? scratch[32 + lane] : -INFINITY;
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
  }
  __syncthreads();
}

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// MUFU square root (2 ulp; subnormal inputs flush to 0 -- far below the 1e-5 absolute tolerance of every feature).
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Streaming (read-once) global load that does not allocate in L1.
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

}  // namespace cacfe
