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

// ---- mel job tables (mel_jobs.h): run the kernel's quad loop on the host ------------------------------------------
#include "mel_jobs.h"

// power: [2 * n_chunks] bins of one frame (frame B of the pair is set to 2x frame A); out: [n_mels] for frame A and
// [n_mels] for frame B.  Returns 0, or a positive code when the tables are inconsistent.
extern "C" int mel_jobs_emul(const float* bank, int n_mels, int n_bins, int n_chunks, const float* power, float* out_a,
                             float* out_b, int* total_quads) {
  const MelJobs J = build_mel_jobs(bank, n_mels, n_bins, n_chunks);
  if (!J.ok) return 1;
  *total_quads = J.total_quads;
  // the kernel's power buffer: float2 (4 |XA|^2, 4 |XB|^2) per bin (the weights carry the 1/4)
  std::vector<float> pbuf((size_t)4 * n_chunks, 0.0f);
  for (int k = 0; k < 2 * n_chunks; ++k) {
    pbuf[2 * k] = 4.0f * power[k];
    pbuf[2 * k + 1] = 8.0f * power[k];
  }
  std::vector<int> written(n_mels, 0);
  int qbase = 0;
  for (int sg = 0; sg < kMelMaxSeg; ++sg) {
    const int nq = J.nq[sg];
    if (nq == 0) continue;
    float acc_a[64], acc_b[64];
    for (int t = 0; t < 64; ++t) {
      const int d = J.desc[sg * 64 + t];
      int c = mel_desc_chunk(d);
      acc_a[t] = acc_b[t] = 0.0f;
      for (int i = 0; i < nq; ++i, c += 2) {
        if (c < 0 || c + 1 >= n_chunks) return 3;
        const float* wv = &J.w[(((size_t)(qbase + i)) * 64 + t) * 4];
        const float* p01 = &pbuf[4 * c];
        const float* p23 = &pbuf[4 * (c + 1)];
        acc_a[t] = std::fmaf(wv[0], p01[0], acc_a[t]);
        acc_b[t] = std::fmaf(wv[0], p01[1], acc_b[t]);
        acc_a[t] = std::fmaf(wv[1], p01[2], acc_a[t]);
        acc_b[t] = std::fmaf(wv[1], p01[3], acc_b[t]);
        acc_a[t] = std::fmaf(wv[2], p23[0], acc_a[t]);
        acc_b[t] = std::fmaf(wv[2], p23[1], acc_b[t]);
        acc_a[t] = std::fmaf(wv[3], p23[2], acc_a[t]);
        acc_b[t] = std::fmaf(wv[3], p23[3], acc_b[t]);
      }
    }
    if (sg == J.split_seg)
      for (int t = 0; t < 64; t += 2) {
        const float sa = acc_a[t] + acc_a[t + 1], sb = acc_b[t] + acc_b[t + 1];
        acc_a[t] = acc_a[t + 1] = sa;
        acc_b[t] = acc_b[t + 1] = sb;
      }
    for (int t = 0; t < 64; ++t) {
      const int d = J.desc[sg * 64 + t];
      if (!mel_desc_stores(d)) continue;
      const int m = mel_desc_band(d);
      if (m >= n_mels) return 4;
      out_a[m] = acc_a[t];
      out_b[m] = acc_b[t];
      written[m]++;
    }
    qbase += nq;
  }
  for (int m = 0; m < n_mels; ++m)
    if (written[m] != 1) return 5;  // every band stored exactly once
  return 0;
}
