// Causal second-order-section IIR filter along the last axis (scipy.signal.sosfilt), parallel in time, sm_100a.
//
// Replaces butter_function -> tf.numpy_function(butter_bandpass_filter) (tfdataset.py:2062-2077), the CPU detour that
// raw_to_mel_dual takes before its STFTs (tfdataset.py:1821), and the per-track band-pass of predict_utils.load_samples
// (predict_utils.py:103-115).  scipy runs the direct-form-II-transposed recurrence of every section sample by sample in
// float64
//     y = b0 x + z1;   z1 = b1 x - a1 y + z2;   z2 = b2 x - a2 y
// and the reference casts the result to float32.  The recurrence is linear in its state s = (z1, z2):
//     s' = A s + B x,   A = [[-a1, 1], [-a2, 0]],   B = (b1 - a1 b0, b2 - a2 b0),
// so a run of L samples maps its start state affinely:  s_end = A^L s_start + c,  c = the run's end state from s = 0.
// One CTA streams one row (clip) in tiles of 256 threads x 32 samples; per tile and section
//   1. every thread runs its 32 samples from the zero state                        -> c          (sequential, in registers)
//   2. the 256 affine maps are composed by a warp shuffle scan (P = A^32 and its powers, host-built in long double) and a
//      short walk over the 8 warp totals; the carry of the previous tile enters at thread 0     -> every run's start state
//   3. every thread runs its 32 samples again from that state with scipy's own operation sequence (FP64, no FMA
//      contraction) and hands the outputs to the next section in registers.
// Loads and stores go through a padded shared-memory tile (coalesced 16-byte global accesses, conflict-free stride-33
// per-thread access).  FP64 work: ~16 instructions per sample and section; HBM: one read and one write of the row.
// The start states carry a relative error of ~1e-16 against the purely sequential evaluation, so the float32 results
// equal scipy's except where the float64 value sits within that distance of a rounding boundary (<= 1 ulp, ~1e-8 of the
// samples).  The round-1 kernel (one thread per clip, 35 ms for 2048 clips) is gone.
#pragma once
#include <cuda_runtime.h>

#include "cacfe_common.cuh"

namespace cacfe {

constexpr int kSosMaxSections = 8;
constexpr int kSosRun = 32;                 // samples per thread and tile
constexpr int kSosThreads = 256;
constexpr int kSosTile = kSosRun * kSosThreads;          // 8192 samples
constexpr int kSosStride = kSosRun + 1;                  // padded run: bank-conflict free per-thread access

struct SosSection {
  double b0, b1, b2, a1, a2;    // normalised by a0
  double c1, c2;                // B = (b1 - a1 b0, b2 - a2 b0)
  double pw[kSosRun + 1][4];    // (A^32)^m, m = 0..32, row-major 2 x 2
};

struct SosArgs {
  const float* in;
  float* out;
  long long rows, n;
  int n_sections;
  SosSection sec[kSosMaxSections];   // by value: 8.9 KB of kernel parameters (the 32 KB parameter space of CUDA >= 12.1)
};

__device__ __forceinline__ void sos_apply(const double (&m)[4], double x1, double x2, double& y1, double& y2) {
  y1 = fma(m[0], x1, m[1] * x2);
  y2 = fma(m[2], x1, m[3] * x2);
}

__global__ void __launch_bounds__(kSosThreads, 2) sosfilt_scan_kernel(const __grid_constant__ SosArgs a) {
  __shared__ float s_tile[kSosThreads * kSosStride];
  __shared__ double s_tot[kSosThreads / 32][2];
  __shared__ double s_carry[kSosMaxSections][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row = blockIdx.x;
  const float* x = a.in + row * a.n;
  float* y = a.out + row * a.n;
  const bool vec = (a.n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (tid < kSosMaxSections) s_carry[tid][0] = s_carry[tid][1] = 0.0;
  for (long long base = 0; base < a.n; base += kSosTile) {
    const long long left = a.n - base;
    const int count = left < kSosTile ? (int)left : kSosTile;
    __syncthreads();                                       // previous tile stored; carries written
    // ---- coalesced load into the padded tile, zeros beyond the row's end --------------------------------------------
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(x + base);
#pragma unroll
      for (int u = 0; u < kSosTile / 4 / kSosThreads; ++u) {
        const int i4 = tid + u * kSosThreads, i = 4 * i4;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (i < count) v = ld_stream4(x4 + i4);
        float* d = s_tile + i + (i >> 5);                  // i % 32 <= 28: the four values stay inside one padded run
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      for (int i = tid; i < kSosTile; i += kSosThreads) s_tile[i + (i >> 5)] = i < count ? x[base + i] : 0.0f;
    }
    __syncthreads();
    double v[kSosRun];
#pragma unroll
    for (int j = 0; j < kSosRun; ++j) v[j] = (double)s_tile[tid * kSosStride + j];

    for (int s = 0; s < a.n_sections; ++s) {
      const SosSection& q = a.sec[s];
      const double b0 = q.b0, b1 = q.b1, b2 = q.b2, a1 = q.a1, a2 = q.a2, c1 = q.c1, c2 = q.c2;
      // 1. zero-state run: state only
      double z1 = 0.0, z2 = 0.0;
#pragma unroll
      for (int j = 0; j < kSosRun; ++j) {
        const double n1 = fma(c1, v[j], fma(-a1, z1, z2));
        z2 = fma(c2, v[j], -a2 * z1);
        z1 = n1;
      }
      // the carry of the previous tile enters through thread 0:  c_0 += P s_carry
      double carry1 = 0.0, carry2 = 0.0;                   // thread 0: the state the previous tile ended in = its own start state
      if (tid == 0) {
        carry1 = s_carry[s][0];
        carry2 = s_carry[s][1];
        double m[4] = {q.pw[1][0], q.pw[1][1], q.pw[1][2], q.pw[1][3]}, p1, p2;
        sos_apply(m, carry1, carry2, p1, p2);
        z1 += p1;
        z2 += p2;
      }
      // 2. inclusive scan of the affine maps inside the warp:  C_k <- P^(d) C_(k-d) + C_k
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double u1 = __shfl_up_sync(kFullMask, z1, d), u2 = __shfl_up_sync(kFullMask, z2, d);
        if (lane >= d) {
          double m[4] = {q.pw[d][0], q.pw[d][1], q.pw[d][2], q.pw[d][3]}, p1, p2;
          sos_apply(m, u1, u2, p1, p2);
          z1 += p1;
          z2 += p2;
        }
      }
      if (lane == 31) {
        s_tot[warp][0] = z1;
        s_tot[warp][1] = z2;
      }
      // state at the end of the previous thread's run, before the earlier warps are folded in
      double e1 = __shfl_up_sync(kFullMask, z1, 1), e2 = __shfl_up_sync(kFullMask, z2, 1);
      if (lane == 0) e1 = e2 = 0.0;
      __syncthreads();
      // state at the start of this warp: E_w = P^32 E_(w-1) + T_(w-1)
      double w1 = 0.0, w2 = 0.0;
      {
        double m[4] = {q.pw[32][0], q.pw[32][1], q.pw[32][2], q.pw[32][3]};
        for (int u = 0; u < warp; ++u) {
          double p1, p2;
          sos_apply(m, w1, w2, p1, p2);
          w1 = p1 + s_tot[u][0];
          w2 = p2 + s_tot[u][1];
        }
      }
      {
        double m[4] = {q.pw[lane][0], q.pw[lane][1], q.pw[lane][2], q.pw[lane][3]}, p1, p2;
        sos_apply(m, w1, w2, p1, p2);                      // P^lane E_w
        z1 = e1 + p1 + carry1;                             // (the carry is already inside every later run's end state)
        z2 = e2 + p2 + carry2;
      }
      // 3. the run again from its true start state, scipy's operation sequence; the last thread leaves the tile's carry
#pragma unroll
      for (int j = 0; j < kSosRun; ++j) {
        const double xin = v[j];
        const double o = __dadd_rn(__dmul_rn(b0, xin), z1);
        z1 = __dadd_rn(__dsub_rn(__dmul_rn(b1, xin), __dmul_rn(a1, o)), z2);
        z2 = __dsub_rn(__dmul_rn(b2, xin), __dmul_rn(a2, o));
        v[j] = o;
      }
      __syncthreads();                                     // s_tot and s_carry[s] have been read by everyone
      if (tid == kSosThreads - 1) {
        s_carry[s][0] = z1;
        s_carry[s][1] = z2;
      }
    }
    // ---- outputs back through the tile --------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < kSosRun; ++j) s_tile[tid * kSosStride + j] = (float)v[j];
    __syncthreads();
    if (vec) {
      float4* y4 = reinterpret_cast<float4*>(y + base);
#pragma unroll
      for (int u = 0; u < kSosTile / 4 / kSosThreads; ++u) {
        const int i4 = tid + u * kSosThreads, i = 4 * i4;
        if (i < count) {
          const float* d = s_tile + i + (i >> 5);
          y4[i4] = make_float4(d[0], d[1], d[2], d[3]);
        }
      }
    } else {
      for (int i = tid; i < count; i += kSosThreads) y[base + i] = s_tile[i + (i >> 5)];
    }
  }
}

}  // namespace cacfe
