// Per-thread building blocks of the fused frame/window/FFT/power kernel (K1).
//
// Reference path replaced: tf.signal.stft(4096, 281, pad_end) -> pow(z,2) -> abs
// (tfdataset.py:2026-2046) and np.abs(librosa.stft(...))**power (predict_utils.py:194,
// custommel.py:59).
//
// Two real frames A and B ride one 4096-point complex FFT:  z[n] = w[n] (xA[n] + i xB[n]).
// 4096 = 64 x 64.  With n = 64 n1 + n2 and k = k1 + 64 k2:
//     Z[k1 + 64 k2] = sum_n2 W64^(n2 k2) * W4096^(n2 k1) * sum_n1 W64^(n1 k1) z[64 n1 + n2]
//   stage 1: thread n2 holds z[64 n1 + n2] (n1 = 0..63) in registers, runs cacfe_fft64,
//            multiplies by W4096^(n2 k1) (table tw[k1][n2]) and stores row k1 / column n2 of the exchange tile;
//   stage 2: thread j (= k1) loads row j, runs cacfe_fft64 and owns Z[j + 64 q], q = 0..63.
// The frames are separated without another pass over shared memory:
//     XA[k] = (Z[k] + conj Z[N-k]) / 2      XB[k] = (Z[k] - conj Z[N-k]) / (2i)
// and Z[N-k] lives in thread 64-j, slot 63-q  (thread 0: own slot 64-q), fetched by shuffle.
//
// Everything here is __host__ __device__ so tests/test_kernel_emulation.py can execute the
// exact index arithmetic on the CPU (as a checker of the kernel source, not as a fallback).
#pragma once
#include "fft64_gen.cuh"

#ifndef __CUDACC__
#include <cmath>
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
#endif

namespace cacfe {

constexpr int kFft = 4096;         // complex FFT length == frame length
constexpr int kRadix = 64;
constexpr int kExchStride = 66;    // float2 per exchange row: 528 B = 4*128 + 16 -> LDS.128 rows conflict free
constexpr int kExchFloat2 = kRadix * kExchStride;

// Hann window by symmetry: w[n] = w[4096 - n], table holds n = 0..2048.
CACFE_HD int win_index(int n) { return n <= kFft / 2 ? n : kFft - n; }

// Stage 1 for the thread owning column n2.  `fa`/`fb` point at sample 0 of frames A / B inside the
// (already normalised, already zero padded) sample tile.
CACFE_HD void stage1(const float* fa, const float* fb, const float* win, const float2* tw, int n2,
                     float2* exch) {
  float re[64], im[64];
#pragma unroll
  for (int a = 0; a < 64; ++a) {
    const int n = 64 * a + n2;
    const float w = win[a < 32 ? n : kFft - n];
    re[a] = fa[n] * w;
    im[a] = fb[n] * w;
  }
  cacfe_fft64(re, im);
#pragma unroll
  for (int k1 = 0; k1 < 64; ++k1) {
    const int s = CACFE_FFT64_SLOT(k1);
    const float2 t = tw[k1 * 64 + n2];  // table laid out [k1][n2]: lanes (n2) read consecutive words
    float2 y;
    y.x = re[s] * t.x - im[s] * t.y;
    y.y = re[s] * t.y + im[s] * t.x;
    exch[k1 * kExchStride + n2] = y;
  }
}

// Stage 2 load + FFT for the thread owning row j: afterwards Z[j + 64 q] is in slot CACFE_FFT64_SLOT(q).
CACFE_HD void stage2_load(const float2* exch, int j, float (&re)[64], float (&im)[64]) {
  const float4* row = reinterpret_cast<const float4*>(exch + j * kExchStride);  // 528 B rows: 16 B aligned
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float4 v = row[i];
    re[2 * i] = v.x;
    im[2 * i] = v.y;
    re[2 * i + 1] = v.z;
    im[2 * i + 1] = v.w;
  }
}

// Stage-2 thread slot -> row j.  Lane l of warp 0 owns {0..15, 32, 63..49}, warp 1 {16..31, 48..33}, so the
// owner of row 64-j is always lane^16 of the same warp (rows 0 and 32 are their own partners).
CACFE_HD int stage2_row(int tid64) {
  const int warp = tid64 >> 5, lane = tid64 & 31;
  if (warp == 0) return lane < 16 ? lane : (lane == 16 ? 32 : 80 - lane);
  return lane < 16 ? 16 + lane : 64 - lane;
}

// Power (|X|^2, or |X| when power == 1) of frames A and B at one bin from Z[k] = (zr, zi) and Z[N-k] = (pr, pi).
CACFE_HD void split_power(float zr, float zi, float pr, float pi, int power, float& pa, float& pb) {
  const float ar = zr + pr, ai = zi - pi;   // 2 XA
  const float br = zr - pr, bi = zi + pi;   // 2i XB
  pa = 0.25f * (ar * ar + ai * ai);
  pb = 0.25f * (br * br + bi * bi);
  if (power == 1) {
    pa = sqrtf(pa);
    pb = sqrtf(pb);
  }
}

}  // namespace cacfe
