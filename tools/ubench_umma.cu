// What a tensor-core DFT stage would cost on this SM, measured: the dispatch time of small-N tcgen05.mma (kind::f16, K = 16 per
// instruction) with the operands in shared memory or A in tensor memory, and what the operand reads do to the LDS bandwidth the
// rest of a fused kernel needs.  DESIGN.md section 3.1a uses the numbers.  Build and run (one GPU):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I audio-training_b200/csrc -o tools/ubench_umma tools/ubench_umma.cu
//   tools/ubench_umma
// One CTA per SM.  Warp 0 owns tensor memory and (lane 0) issues `reps` x 4 MMAs back to back, one commit at the end; warps 1..8
// run a conflict-free LDS.128 loop over a 16 KB region.  Three runs per shape: MMA alone, LDS alone, both together;
// each with all MMAs accumulating into one tensor-memory tile (a dependent chain) and with four tiles in rotation.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#include "cacfe_async.cuh"
using namespace cacfe;

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc_f16(int m, int n) {   // D = F32, A = B = F16, both K-major
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

constexpr int kLdsWarps = 8;
constexpr int kOperandBytes = 64 * 1024;   // A region 32 KB, B region 32 KB
constexpr int kLdsBytes = 16 * 1024;

template <int a_in_tmem>
__global__ void __launch_bounds__(32 * (1 + kLdsWarps), 1)
ubench(int M, int N, int reps, int mma_on, int lds_iters, int n_acc, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOperandBytes + kLdsBytes);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (kOperandBytes + kLdsBytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;
  unsigned long long t_mma = 0, t_lds = 0;
  if (warp == 0) {
    if (mma_on) {   // the whole warp walks the loop converged; one elected lane issues (no per-instruction election loop in the SASS)
      const uint32_t sa = smem_u32(smem), sb = sa + kOperandBytes / 2;
      const uint32_t idesc = idesc_f16(M, N);
      const uint32_t lbo_a = (uint32_t)(M / 8) * 128u, lbo_b = (uint32_t)(N / 8) * 128u;
      const uint32_t a_tm = tmem + 480;   // A in tensor memory: M lanes x 8 columns per K = 16 step
      const unsigned long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t d_tm = tmem + (uint32_t)((r % n_acc) * N);   // n_acc independent accumulators (output tiles) in rotation
#pragma unroll
        for (int j = 0; j < 4; ++j) {     // four K steps = one K = 64 block of a 64-point DFT stage
          const uint64_t bd = smem_desc(sb + j * 2 * lbo_b, lbo_b, 128);
          if (a_in_tmem) {
            asm volatile(
                "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tm),
                "r"(a_tm + 8 * j), "l"(bd), "r"(idesc), "r"(1u)
                : "memory");
          } else {
            const uint64_t ad = smem_desc(sa + j * 2 * lbo_a, lbo_a, 128);
            asm volatile(
                "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tm),
                "l"(ad), "l"(bd), "r"(idesc), "r"(1u)
                : "memory");
          }
        }
      }
      asm volatile(
          "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
          "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
          : "memory");
      mbar_wait(smem_u32(bar), 0);
      t_mma = clock64() - t0;
    }
  } else if (lds_iters > 0) {
    const float4* p = reinterpret_cast<const float4*>(smem + kOperandBytes) + lane;   // 32 lanes x 16 B = 512 B per access
    float acc = 0.0f;
    const unsigned long long t0 = clock64();
    for (int i = 0; i < lds_iters; ++i) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(smem_u32(p + ((i + u) & 31) * 32)));
        acc += v.x + v.y + v.z + v.w;
      }
    }
    t_lds = clock64() - t0;
    if (acc == 123.456f) out[0] = 1;   // keep the loads
  }
  if (lane == 0) {
    if (warp == 0) out[1 + blockIdx.x * 2] = t_mma;
    if (warp == 1) out[2 + blockIdx.x * 2] = t_lds;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem_bytes = kOperandBytes + kLdsBytes + 64;
  cudaFuncSetAttribute(ubench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(ubench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  unsigned long long* d_out;
  cudaMalloc(&d_out, (2 * sms + 1) * sizeof(unsigned long long));
  std::vector<unsigned long long> h(2 * sms + 1);
  auto run = [&](int M, int N, int reps, int a_tm, int mma_on, int lds_iters, int n_acc, double* mma_cyc, double* lds_cyc) {
    cudaMemset(d_out, 0, h.size() * sizeof(unsigned long long));
    if (a_tm) ubench<1><<<sms, 32 * (1 + kLdsWarps), smem_bytes>>>(M, N, reps, mma_on, lds_iters, n_acc, d_out);
    else ubench<0><<<sms, 32 * (1 + kLdsWarps), smem_bytes>>>(M, N, reps, mma_on, lds_iters, n_acc, d_out);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("launch failed: %s\n", cudaGetErrorString(e));
      exit(1);
    }
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double sm = 0, sl = 0;
    for (int b = 0; b < sms; ++b) {
      sm += (double)h[1 + 2 * b];
      sl += (double)h[2 + 2 * b];
    }
    *mma_cyc = mma_on ? sm / sms / (4.0 * reps) : 0.0;                        // cycles per MMA instruction
    *lds_cyc = lds_iters ? sl / sms / (16.0 * lds_iters * kLdsWarps) : 0.0;   // SM cycles per warp-wide LDS.128
  };
  printf("# one CTA per SM on %d SMs; kind::f16, K = 16 per instruction, FP32 accumulate\n", sms);
  printf("# M N A_operand accumulators | cycles/MMA alone | cycles per LDS.128 alone | cycles/MMA with LDS | cycles per LDS.128 with MMA | "
         "operand bytes per MMA\n");
  const int shapes[][2] = {{64, 32}, {64, 64}, {64, 128}, {64, 256}, {128, 32}, {128, 64}, {128, 128}, {128, 256}};
  for (int a_tm = 0; a_tm < 2; ++a_tm)
    for (auto& s : shapes)
     for (int n_acc = 1; n_acc <= 4; n_acc *= 4) {
      const int M = s[0], N = s[1];
      if (n_acc * N > 448) continue;
      const int reps = 4096;
      double m0, l0, m1, l1, dummy;
      run(M, N, reps, a_tm, 1, 0, n_acc, &m0, &dummy);
      // size the LDS loop to the MMA run: 8 warps x 16 loads per iteration at ~4 cycles each
      const int lds_iters = (int)(m0 * 4.0 * reps / (16.0 * kLdsWarps * 4.0)) + 64;
      run(M, N, reps, a_tm, 0, lds_iters, n_acc, &dummy, &l0);
      run(M, N, reps, a_tm, 1, lds_iters, n_acc, &m1, &l1);
      const int bytes = (a_tm ? 0 : M * 32) + N * 32;
      printf("%4d %4d %s %d | %8.1f | %6.2f | %8.1f | %6.2f | %d\n", M, N, a_tm ? "tmem" : "smem", n_acc, m0, l0, m1, l1, bytes);
    }
  return 0;
}
