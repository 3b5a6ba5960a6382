// Micro-benchmark (sm_100a): issue rate of scalar FADD/FFMA against the packed FADD2/FFMA2 forms, alone and mixed
// with shared-memory loads, at the occupancy of the K1 kernel (12 warps per SM) and at 32 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu && ./ubench_f32x2
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float addv(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmav(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr int CH = 16;  // independent chains per thread
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float acc = 0.f;
  if (MODE == 0 || MODE == 4) {  // scalar FADD (+ LDS for 4)
    float v[CH];
    for (int c = 0; c < CH; ++c) v[c] = seed + c + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int c = 0; c < CH; ++c) v[c] = addv(v[c], seed);
        if (MODE == 4) {
#pragma unroll
          for (int c = 0; c < 4; ++c) acc += sm[(threadIdx.x + 32 * (4 * r + c) + it) & 4095];
        }
      }
    }
    for (int c = 0; c < CH; ++c) acc += v[c];
  } else if (MODE == 1 || MODE == 5) {  // FADD2 (+ LDS for 5): same number of instructions, twice the flops
    u64 v[CH];
    const u64 s2 = pk(seed, seed);
    for (int c = 0; c < CH; ++c) v[c] = pk(seed + c + threadIdx.x, seed - c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int c = 0; c < CH; ++c) v[c] = add2(v[c], s2);
        if (MODE == 5) {
#pragma unroll
          for (int c = 0; c < 4; ++c) acc += sm[(threadIdx.x + 32 * (4 * r + c) + it) & 4095];
        }
      }
    }
    for (int c = 0; c < CH; ++c) { float a, b; upk(v[c], a, b); acc += a + b; }
  } else if (MODE == 2) {  // scalar FFMA (3 register operands)
    float v[CH];
    for (int c = 0; c < CH; ++c) v[c] = seed + c + threadIdx.x;
    const float m = seed * 0.5f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < CH; ++c) v[c] = fmav(v[c], m, seed);
    }
    for (int c = 0; c < CH; ++c) acc += v[c];
  } else if (MODE == 3) {  // FFMA2
    u64 v[CH];
    const u64 s2 = pk(seed, seed), m2 = pk(seed * 0.5f, seed * 0.25f);
    for (int c = 0; c < CH; ++c) v[c] = pk(seed + c + threadIdx.x, seed - c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < CH; ++c) v[c] = fma2(v[c], m2, s2);
    }
    for (int c = 0; c < CH; ++c) { float a, b; upk(v[c], a, b); acc += a + b; }
  }
  if (acc == 12345.678f) out[0] = acc;
}

template <int MODE>
void run(const char* name, int threads, int blocks_per_sm) {
  float* d; cudaMalloc(&d, 4);
  const int iters = 2000, grid = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, threads>>>(d, 10, 1.0f);
  cudaEventRecord(e0);
  k<MODE><<<grid, threads>>>(d, iters, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fp_instr = (double)iters * 8 * CH * (threads / 32) * grid;  // warp-level FP instructions
  const double per_clk_sm = fp_instr / (ms * 1e-3 * 1.965e9) / 148.0;
  printf("%-28s threads=%4d x%d: %.3f ms  %.3f FP warp-instr/clk/SM (%s)\n", name, threads, blocks_per_sm, ms, per_clk_sm,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int threads = cfg == 0 ? 384 : 1024;
    run<0>("FADD", threads, 1);
    run<1>("FADD2", threads, 1);
    run<2>("FFMA rrr", threads, 1);
    run<3>("FFMA2", threads, 1);
    run<4>("FADD + 1 LDS per 4", threads, 1);
    run<5>("FADD2 + 1 LDS per 4", threads, 1);
  }
  return 0;
}
