// CPU emulation of one K1 frame pair (test infrastructure: checks the kernel's index arithmetic on the host;
// it includes the very header the CUDA kernel is built from).  Built by tests/test_kernel_emulation.py.
#include <cmath>
#include <vector>
#include "frontend_core.cuh"

using namespace cacfe;

extern "C" int k1_emul_pair(const float* fa, const float* fb, int power, float* pa_out, float* pb_out /* [2049] each */) {
  std::vector<float2> tw(4096);
  for (int k1 = 0; k1 < 64; ++k1)
    for (int n2 = 0; n2 < 64; ++n2) {
      const double ang = -2.0 * M_PI * (double)((k1 * n2) % 4096) / 4096.0;
      tw[k1 * 64 + n2] = float2{(float)std::cos(ang), (float)std::sin(ang)};
    }
  std::vector<float> win(2049);
  for (int n = 0; n <= 2048; ++n) win[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / 4096.0));
  alignas(16) static float2 exch[kExchFloat2];
  for (int n2 = 0; n2 < 64; ++n2) stage1(fa, fb, win.data(), tw.data(), n2, exch);
  static float re[64][64], im[64][64];
  int row_of[64];
  for (int t = 0; t < 64; ++t) {
    row_of[t] = stage2_row(t);
    stage2_load(exch, row_of[t], re[t], im[t]);
    cacfe_fft64(re[t], im[t]);
  }
  // every row owned exactly once
  int seen[64] = {0};
  for (int t = 0; t < 64; ++t) seen[row_of[t]]++;
  for (int j = 0; j < 64; ++j) if (seen[j] != 1) return 1;
  for (int t = 0; t < 64; ++t) {
    const int j = row_of[t];
    const bool self = (j == 0) || (j == 32);
    const int pt = self ? t : ((t & 32) | ((t & 31) ^ 16));  // __shfl_sync source: lane ^ 16 of the same warp
    if (!self && row_of[pt] != 64 - j) return 2;
    for (int q = 0; q < 33; ++q) {
      float pr = re[pt][CACFE_FFT64_SLOT(63 - q)], pi = im[pt][CACFE_FFT64_SLOT(63 - q)];
      if (j == 0) {
        pr = re[t][CACFE_FFT64_SLOT((64 - q) & 63)];
        pi = im[t][CACFE_FFT64_SLOT((64 - q) & 63)];
      }
      const int k = j + 64 * q;
      if (k <= 2048) split_power(re[t][CACFE_FFT64_SLOT(q)], im[t][CACFE_FFT64_SLOT(q)], pr, pi, power, pa_out[k], pb_out[k]);
    }
  }
  return 0;
}
