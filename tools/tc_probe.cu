// Probe for k_melspec_tc.cuh: one tile, small shapes, prints the first mismatches.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I audio-training_b200/csrc -o tools/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>
#include "k_melspec_tc.cuh"
using namespace cacfe;

int main(int argc, char** argv) {
  const int n_bins = argc > 1 ? atoi(argv[1]) : 32, n_mels = argc > 2 ? atoi(argv[2]) : 16, T = argc > 3 ? atoi(argv[3]) : 128;
  const int B = 1;
  std::vector<float> spec((size_t)B * n_bins * T), bank((size_t)n_mels * n_bins, 0.0f);
  for (int k = 0; k < n_bins; ++k)
    for (int t = 0; t < T; ++t) spec[(size_t)k * T + t] = 1.0f + 0.001f * k + 0.01f * t;
  for (int m = 0; m < n_mels; ++m)
    for (int k = 0; k < n_bins; ++k) bank[(size_t)m * n_bins + k] = ((k + m) % 5 == 0) ? 0.1f * (m + 1) + 0.003f * k : 0.0f;
  // chunks: dense over all bands (n0 = 0, nc = n_mels rounded to 16)
  std::vector<MelTcChunk> chunks;
  std::vector<float> wpk;
  const int nc = (n_mels + 15) & ~15;
  for (int k0 = 0; k0 < n_bins; k0 += kTcK) {
    MelTcChunk ch{k0, 0, nc, (int)wpk.size()};
    const size_t block = (size_t)nc * kTcK;
    wpk.resize(wpk.size() + 2 * block, 0.0f);
    for (int n = 0; n < nc; ++n)
      for (int kk = 0; kk < kTcK; ++kk) {
        const int k = k0 + kk;
        const float w = (k < n_bins && n < n_mels) ? bank[(size_t)n * n_bins + k] : 0.0f;
        uint32_t bits; memcpy(&bits, &w, 4); bits &= 0xffffe000u; float hi; memcpy(&hi, &bits, 4);
        const size_t idx = (size_t)(kk / 4) * (nc / 8) * 32 + (size_t)(n / 8) * 32 + (n % 8) * 4 + (kk % 4);
        wpk[ch.w_ofs + idx] = hi;
        wpk[ch.w_ofs + block + idx] = w - hi;
      }
    chunks.push_back(ch);
  }
  float *d_spec, *d_out, *d_w; MelTcChunk* d_ch;
  cudaMalloc(&d_spec, spec.size() * 4); cudaMalloc(&d_out, (size_t)B * nc * T * 4); cudaMalloc(&d_w, wpk.size() * 4);
  cudaMalloc(&d_ch, chunks.size() * sizeof(MelTcChunk));
  cudaMemcpy(d_spec, spec.data(), spec.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_w, wpk.data(), wpk.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_ch, chunks.data(), chunks.size() * sizeof(MelTcChunk), cudaMemcpyHostToDevice);
  cudaMemset(d_out, 0xff, (size_t)B * nc * T * 4);
  MelTcArgs a{d_spec, d_out, d_w, d_ch, (int)chunks.size(), n_bins, T, nc, 1, 1, 0, (T + kTcM - 1) / kTcM, 0};
  a.tile_frames = (T + a.tiles_per_clip - 1) / a.tiles_per_clip;
  cudaFuncSetAttribute(melspec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
  melspec_tc_kernel<<<B * a.tiles_per_clip, kTcThreads, kTcSmemBytes>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> out((size_t)B * nc * T);
  cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double worst = 0;
  for (int m = 0; m < n_mels; ++m)
    for (int t = 0; t < T; ++t) {
      double want = 0;
      for (int k = 0; k < n_bins; ++k) want += (double)bank[(size_t)m * n_bins + k] * spec[(size_t)k * T + t];
      const float got = out[(size_t)m * T + t];
      const double err = fabs(got - want) / (fabs(want) + 1e-6);
      if (err > worst) worst = err;
      if (err > 1e-5 && bad < 12) { printf("m=%d t=%d got %.7g want %.7g\n", m, t, got, want); ++bad; }
    }
  { // value histogram (debug)
    int shown = 0; float last = -12345.f;
    for (int m = 0; m < nc && shown < 24; ++m) for (int t = 0; t < T && shown < 24; ++t) { float v = out[(size_t)m * T + t]; if (v != last) { printf("  [m=%d t=%d] %.6g\n", m, t, v); last = v; ++shown; } }
  }
  printf("worst rel err %.3g (n_bins %d n_mels %d T %d)\n", worst, n_bins, n_mels, T);
  return 0;
}
