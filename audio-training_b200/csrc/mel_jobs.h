// Host-side layout of the banded mel projection for the persistent K1 kernel (k_frontend_v3.cuh).
//
// custommel.mel_f (custommel.py:18-54) gives a [n_mels][n_bins] bank whose rows are short runs of non-zeros (3..30
// taps for the reference's 160-band bank).  A 64-thread FFT group owns the power of one frame pair; its threads
// share the bands out so that every thread sees about the same number of taps:
//   segment s < n_mels / 64 : band 64 s + t (s even) or 64 s + 63 - t (s odd)      -- widths grow with the band index
//   last, short segment     : (n_mels % 64 <= 32) each band is cut in two halves for lanes 2i / 2i+1, summed by shuffle
// The weights are stored times 1/4 (exact): the kernel keeps 4 |X|^2 (or 4 |X|), see k_frontend_v3.cuh.
// Taps are consumed four bins at a time ("quads": one 16-byte weight load, two 16-byte power loads, 8 FMAs).  Every
// thread of a segment runs the same number of quads; the surplus taps carry weight 0.
// Plain C++ (no CUDA): tests/emul exercises this builder against the dense bank on the CPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace cacfe {

constexpr int kMelMaxSeg = 3;

struct MelJobs {
  bool ok = false;
  int nseg = 0;
  int nq[kMelMaxSeg] = {0, 0, 0};  // quads per thread in segment s (uniform over the 64 threads)
  int split_seg = -1;              // segment whose bands are shared by lane pairs, or -1
  int total_quads = 0;
  std::vector<float> w;            // [total_quads][64][4]
  std::vector<int32_t> desc;       // [kMelMaxSeg][64]: band | first 16-byte chunk << 8 | valid << 24 | stores << 25
};

inline int mel_desc_band(int32_t d) { return d & 0xff; }
inline int mel_desc_chunk(int32_t d) { return (d >> 8) & 0xffff; }
inline bool mel_desc_valid(int32_t d) { return (d >> 24) & 1; }
inline bool mel_desc_stores(int32_t d) { return (d >> 25) & 1; }

// n_chunks: 16-byte chunks (2 bins) of the power buffer the kernel fills = 32 * NQ.
inline MelJobs build_mel_jobs(const float* bank, int n_mels, int n_bins, int n_chunks) {
  MelJobs J;
  if (n_mels < 1 || n_mels > 64 * kMelMaxSeg) return J;
  const int n_full = n_mels / 64, rem = n_mels % 64;
  J.nseg = n_full + (rem ? 1 : 0);
  J.split_seg = (rem > 0 && rem <= 32) ? n_full : -1;
  struct Job { int m = -1, ka = 0, kb = -1; bool valid = false, stores = false; };
  std::vector<Job> jobs(kMelMaxSeg * 64);
  auto run_of = [&](int m, int& a, int& b) {
    a = 0; b = -1;
    const float* row = bank + (size_t)m * n_bins;
    int first = -1, last = -1;
    for (int k = 0; k < n_bins; ++k)
      if (row[k] != 0.0f) { if (first < 0) first = k; last = k; }
    if (first >= 0) { a = first; b = last; }
  };
  for (int s = 0; s < J.nseg; ++s)
    for (int t = 0; t < 64; ++t) {
      Job& j = jobs[s * 64 + t];
      if (s < n_full) {
        j.m = 64 * s + ((s & 1) ? 63 - t : t);
        j.valid = j.stores = true;
        run_of(j.m, j.ka, j.kb);
      } else if (J.split_seg == s) {
        if ((t >> 1) < rem) {
          j.m = 64 * s + (t >> 1);
          j.valid = true;
          j.stores = (t & 1) == 0;
          int a, b;
          run_of(j.m, a, b);
          const int n = b - a + 1, h0 = (n + 1) / 2;
          if (t & 1) { j.ka = a + h0; j.kb = b; } else { j.ka = a; j.kb = a + h0 - 1; }
        }
      } else if (t < rem) {
        j.m = 64 * s + t;
        j.valid = j.stores = true;
        run_of(j.m, j.ka, j.kb);
      }
    }
  // quads per segment
  for (int s = 0; s < J.nseg; ++s) {
    int q = 0;
    for (int t = 0; t < 64; ++t) {
      const Job& j = jobs[s * 64 + t];
      if (!j.valid || j.kb < j.ka) continue;
      if (j.kb >= 2 * n_chunks) return J;  // a tap beyond the bins the kernel computes
      const int c0 = j.ka >> 1;
      q = std::max(q, (j.kb - 2 * c0 + 1 + 3) / 4);
    }
    if (2 * q > n_chunks) return J;
    J.nq[s] = q;
    J.total_quads += q;
  }
  J.w.assign((size_t)std::max(J.total_quads, 1) * 64 * 4, 0.0f);
  J.desc.assign(kMelMaxSeg * 64, 0);
  int qbase = 0;
  for (int s = 0; s < J.nseg; ++s) {
    for (int t = 0; t < 64; ++t) {
      const Job& j = jobs[s * 64 + t];
      int c0 = 0;
      if (j.valid && j.kb >= j.ka) {
        c0 = j.ka >> 1;
        if (c0 + 2 * J.nq[s] > n_chunks) c0 = n_chunks - 2 * J.nq[s];  // stay inside the buffer: pad in front instead
        for (int i = 0; i < J.nq[s]; ++i)
          for (int e = 0; e < 4; ++e) {
            const int bin = 2 * c0 + 4 * i + e;
            if (bin >= j.ka && bin <= j.kb)
              J.w[(((size_t)(qbase + i)) * 64 + t) * 4 + e] = 0.25f * bank[(size_t)j.m * n_bins + bin];  // exact
          }
      }
      J.desc[s * 64 + t] = (j.valid ? (j.m & 0xff) : 0) | (c0 << 8) | ((j.valid ? 1 : 0) << 24) | ((j.stores ? 1 : 0) << 25);
    }
    qbase += J.nq[s];
  }
  J.ok = true;
  return J;
}


}  // namespace cacfe
