// Path C: stored spectrogram [B][K][T] (t contiguous) -> mel image, banded FP32 SpMM, sm_100a.
//
// Replaces tf.tensordot(MEL_WEIGHTS, spectrogram(2049x513), 1) (tfdataset.py:1082-1090; power 1 because
// pcen=True skips the squaring, Q6) and, with power 2, custommel.mel_spec (custommel.py:57-61).
//
// Only rows bin_lo..bin_hi of the spectrogram carry non-zero weights (930 of 2049 for the default bank),
// so those rows are the compulsory traffic: 930*513*4 B in, 160*513*4 B out per clip.  Each thread owns
// one time column; lanes read consecutive t -> every load is a coalesced 128 B row segment.  A bin feeds
// at most two adjacent bands, so the second read of a row hits L1.  HBM bound.
#pragma once
#include "cacfe_common.cuh"

namespace cacfe {

struct MelSpecArgs {
  const float* spec;      // [B][n_bins][T]
  float* out;             // [B][M][T][C] or [B][T][M]
  const float* band_w;
  const int* band_start;  // relative to bin_lo
  const int* band_ofs;
  int n_bins, T, n_mels, nnz, bin_lo, power, channels, layout;
};

constexpr int kMelSpecThreads = 128;

// grid = (ceil(T / 128), B); dynamic smem = nnz floats + (2 M + 1) ints
__global__ void __launch_bounds__(kMelSpecThreads) melspec_banded_kernel(const MelSpecArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* s_w = reinterpret_cast<float*>(smem);
  int* s_start = reinterpret_cast<int*>(s_w + ((a.nnz + 3) & ~3));
  int* s_ofs = s_start + a.n_mels;
  for (int i = threadIdx.x; i < a.nnz; i += kMelSpecThreads) s_w[i] = a.band_w[i];
  for (int i = threadIdx.x; i < a.n_mels; i += kMelSpecThreads) s_start[i] = a.band_start[i];
  for (int i = threadIdx.x; i <= a.n_mels; i += kMelSpecThreads) s_ofs[i] = a.band_ofs[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * kMelSpecThreads + threadIdx.x;
  if (t >= a.T) return;
  const float* col = a.spec + ((size_t)b * a.n_bins + a.bin_lo) * a.T + t;
  for (int m = 0; m < a.n_mels; ++m) {
    const int o0 = s_ofs[m], o1 = s_ofs[m + 1];
    const float* p = col + (size_t)s_start[m] * a.T;
    float acc = 0.0f;
    int i = o0;
    for (; i + 4 <= o1; i += 4) {  // 4 independent row loads in flight
      float v0 = p[(size_t)(i - o0) * a.T], v1 = p[(size_t)(i - o0 + 1) * a.T], v2 = p[(size_t)(i - o0 + 2) * a.T],
            v3 = p[(size_t)(i - o0 + 3) * a.T];
      if (a.power == 2) {
        v0 *= v0; v1 *= v1; v2 *= v2; v3 *= v3;
      }
      acc = fmaf(s_w[i], v0, acc);
      acc = fmaf(s_w[i + 1], v1, acc);
      acc = fmaf(s_w[i + 2], v2, acc);
      acc = fmaf(s_w[i + 3], v3, acc);
    }
    for (; i < o1; ++i) {
      float v = p[(size_t)(i - o0) * a.T];
      if (a.power == 2) v *= v;
      acc = fmaf(s_w[i], v, acc);
    }
    if (a.layout == 1) {
      a.out[((size_t)b * a.T + t) * a.n_mels + m] = acc;
    } else {
      float* o = a.out + (((size_t)b * a.n_mels + m) * a.T + t) * a.channels;
      for (int c = 0; c < a.channels; ++c) o[c] = acc;
    }
  }
}

}  // namespace cacfe
