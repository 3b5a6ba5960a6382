// Causal second-order-section IIR filter along the last axis (scipy.signal.sosfilt), sm_100a.
//
// Replaces butter_function -> tf.numpy_function(butter_bandpass_filter) (tfdataset.py:2062-2077), the CPU detour that
// raw_to_mel_dual takes before its STFTs (tfdataset.py:1821).  scipy runs the direct-form-II-transposed recurrence
//     y = b0 x + z1;  z1 = b1 x - a1 y + z2;  z2 = b2 x - a2 y
// in float64 and the reference casts the result to float32; the kernel does the same arithmetic in FP64 (no FMA
// contraction), one thread per clip, sections cascaded sample by sample.  The recurrence is sequential in time, so the
// parallelism is the batch: this is a correctness path for a variant that is not on the benchmarked path.
#pragma once
#include <cuda_runtime.h>

namespace cacfe {

constexpr int kSosMaxSections = 8;

struct SosArgs {
  const float* in;
  float* out;
  long long rows, n;
  int n_sections;
  double sos[kSosMaxSections][6];  // b0 b1 b2 a0 a1 a2 (a0 == 1 after scipy's normalisation)
};

__global__ void __launch_bounds__(64) sosfilt_kernel(const SosArgs a) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.rows) return;
  const float* x = a.in + row * a.n;
  float* y = a.out + row * a.n;
  double z1[kSosMaxSections], z2[kSosMaxSections];
#pragma unroll
  for (int s = 0; s < kSosMaxSections; ++s) z1[s] = z2[s] = 0.0;
  for (long long i = 0; i < a.n; ++i) {
    double v = (double)x[i];
#pragma unroll
    for (int s = 0; s < kSosMaxSections; ++s) {
      if (s < a.n_sections) {
        const double o = __dadd_rn(__dmul_rn(a.sos[s][0], v), z1[s]);
        z1[s] = __dadd_rn(__dsub_rn(__dmul_rn(a.sos[s][1], v), __dmul_rn(a.sos[s][4], o)), z2[s]);
        z2[s] = __dsub_rn(__dmul_rn(a.sos[s][2], v), __dmul_rn(a.sos[s][5], o));
        v = o;
      }
    }
    y[i] = (float)v;
  }
}

}  // namespace cacfe
