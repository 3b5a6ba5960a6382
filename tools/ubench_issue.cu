// Micro-benchmark (sm_100a): what one scheduler (SMSP) sustains for the instruction kinds the fused STFT kernel is made of,
// at that kernel's occupancy (3 warps per scheduler) -- the cost model behind DESIGN.md section 3.1a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_issue tools/ubench_issue.cu && tools/ubench_issue
// Every mode runs 16 independent dependency chains per thread with DISTINCT source registers (no operand reuse), so that what
// is measured is the pipe / register-file limit and not a latency chain.  Output: cycles per warp instruction per scheduler.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
#define ADD2(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define MUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define ADD1(d, a, b) asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define MUL1(d, a, b) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define FMA1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define FMAI(d, a, c) asm volatile("fma.rn.f32 %0, %1, 0f3F8CCCCD, %2;" : "=f"(d) : "f"(a), "f"(c))
#define IADD(d, a, b) asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define LOPX(d, a, b) asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))

constexpr int CH = 16;
enum { M_FADD, M_FADD2, M_FMUL2, M_FFMA_RRR, M_FFMA_IMM, M_FFMA2, M_FADD2_LDS, M_FADD2_ALU, M_FADD2_FADD, M_FFMA2_ALU, M_LDS, M_FADD_LDS,
       M_FADD2_STS, M_ALU, M_FADD2_2ALU, M_FFMA2_UR, M_FFMA2_UR16, M_COUNT };
const char* kNames[M_COUNT] = {"FADD r,r", "FADD2", "FMUL2", "FFMA r,r,r", "FFMA r,imm,r", "FFMA2 r,r,r", "FADD2 + LDS.32 (1:1)", "FADD2 + IADD (1:1)",
                               "FADD2 + FADD (1:1)", "FFMA2 + IADD (1:1)", "LDS.32 alone", "FADD + LDS.32 (1:1)", "FADD2 + STS.32 (1:1)", "IADD alone",
                               "FADD2 + 2 ALU (1:2)", "FFMA2 r,const,r (1 const)", "FFMA2 r,const,r (16 const)"};

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(float* out, int iters, float seed) {
  __shared__ float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  u64 v[CH], a[CH], b[CH];
  float s[CH], sa[CH], sb[CH];
  int n[CH], m[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    v[c] = pk(seed + c + threadIdx.x, seed - c);
    a[c] = pk(seed * 0.999f + c, seed * 1.001f);
    b[c] = pk(seed * 0.5f - c, seed * 0.25f);
    s[c] = seed + c;
    sa[c] = seed * 0.999f + c * 1e-3f;
    sb[c] = seed * 0.5f - c;
    n[c] = threadIdx.x + c;
    m[c] = c * 7 + (int)seed;
  }
  float acc = 0.f;
  const float* base = sm + (threadIdx.x & 31);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (MODE == M_FADD) ADD1(s[c], s[c], sa[c]);
        if (MODE == M_FADD2) ADD2(v[c], v[c], a[c]);
        if (MODE == M_FMUL2) MUL2(v[c], v[c], a[c]);
        if (MODE == M_FFMA_RRR) FMA1(s[c], s[c], sa[c], sb[c]);
        if (MODE == M_FFMA_IMM) FMAI(s[c], s[c], sa[c]);
        if (MODE == M_FFMA2) FMA2(v[c], v[c], a[c], b[c]);
        if (MODE == M_FADD2_LDS) { ADD2(v[c], v[c], a[c]); acc += base[((c + 16 * r + it) & 255) * 32]; }
        if (MODE == M_FADD2_ALU) { ADD2(v[c], v[c], a[c]); IADD(n[c], n[c], m[c]); }
        if (MODE == M_FADD2_FADD) { ADD2(v[c], v[c], a[c]); ADD1(s[c], s[c], sa[c]); }
        if (MODE == M_FFMA2_ALU) { FMA2(v[c], v[c], a[c], b[c]); IADD(n[c], n[c], m[c]); }
        if (MODE == M_LDS) acc += base[((c + 16 * r + it) & 255) * 32];
        if (MODE == M_FADD_LDS) { ADD1(s[c], s[c], sa[c]); acc += base[((c + 16 * r + it) & 255) * 32]; }
        if (MODE == M_FADD2_STS) { ADD2(v[c], v[c], a[c]); sm[((c + 16 * r) & 255) * 32 + threadIdx.x % 32 + 32 * (threadIdx.x / 32) * 0] = s[c]; }
        if (MODE == M_ALU) IADD(n[c], n[c], m[c]);
        if (MODE == M_FFMA2_UR) FMA2(v[c], v[c], pk(0.999f, 1.001f), b[c]);                    // constant pair: a uniform-register operand
        if (MODE == M_FFMA2_UR16) FMA2(v[c], v[c], pk(0.999f + 1e-3f * c, 1.001f - 1e-3f * c), b[c]);   // 16 different constant pairs
        if (MODE == M_FADD2_2ALU) { ADD2(v[c], v[c], a[c]); IADD(n[c], n[c], m[c]); LOPX(m[c], m[c], n[(c + 1) % CH]); }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += lo(v[c]) + s[c] + (float)n[c] + (float)m[c];
  if (acc == 12345.678f) out[0] = acc;
}

template <int MODE>
void run(int threads) {
  float* d;
  cudaMalloc(&d, 4);
  const int iters = 4000, grid = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<grid, threads>>>(d, 10, 1.0f);
  cudaEventRecord(e0);
  k<MODE><<<grid, threads>>>(d, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double groups = (double)iters * 4 * CH * (threads / 32) / 4.0;   // instruction groups per scheduler
  const double cyc = ms * 1e-3 * 1.965e9 / groups;
  printf("%-24s warps/SM=%2d: %8.3f ms  %.3f cycles per group per scheduler (%s)\n", kNames[MODE], threads / 32, ms, cyc,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

template <int MODE>
void both() {
  run<MODE>(384);
  run<MODE>(128);
}

int main() {
  both<M_FADD>(); both<M_FADD2>(); both<M_FMUL2>(); both<M_FFMA_RRR>(); both<M_FFMA_IMM>(); both<M_FFMA2>();
  both<M_ALU>(); both<M_LDS>(); both<M_FADD_LDS>(); both<M_FADD2_LDS>(); both<M_FADD2_STS>(); both<M_FADD2_ALU>(); both<M_FADD2_2ALU>();
  both<M_FADD2_FADD>(); both<M_FFMA2_ALU>(); both<M_FFMA2_UR>(); both<M_FFMA2_UR16>();
  return 0;
}
