// Host-side layout of the banded mel projection for the persistent K1 kernel (k_frontend_v3.cuh).
//
// custommel.mel_f (custommel.py:18-54) gives a [n_mels][n_bins] bank whose rows are short runs of non-zeros (3..30
// taps for the reference's 160-band bank).  A 64-thread FFT group owns the power of one frame pair; its threads
// share the bands out so that every thread sees about the same number of taps:
//   segment s < n_mels / 64 : bands 64 s .. 64 s + 63, one per thread
//   last, short segment     : (n_mels % 64 <= 32) each band is cut in two halves for a lane pair 2i / 2i+1, summed by shuffle
// Which lane of its warp a band sits on, and at which 16-byte chunk its reads start, is chosen to keep the power loads free of
// shared-memory bank conflicts (see below).
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
  // ---- shared-memory bank conflicts of the power loads ---------------------------------------------------------------
  // A thread reads 16-byte chunks c0, c0 + 1, ... of the power buffer; the eight lanes of a quarter-warp are served
  // together and collide unless their chunk indices differ mod 8 (ncu: 55 % of those wavefronts were replays with every
  // band starting at its own first bin).  Two freedoms, both free of arithmetic: a band may start up to `slack` chunks
  // early (the extra taps carry weight 0, the quad count of the segment is what its widest band needs anyway), and the
  // bands of a segment may sit on any lane of their warp (the kernel takes band and chunk from the descriptor; the halves of
  // a split band stay on a lane pair).  Greedy: units with the fewest feasible residues first, each into the first
  // quarter-warp that still has distinct residues for it.
  std::vector<int> c0_of(kMelMaxSeg * 64, 0);
  for (int s = 0; s < J.nseg; ++s) {
    const int span = n_chunks - 2 * J.nq[s];
    auto range_of = [&](const Job& j, int& lo, int& hi) {
      lo = 0;
      hi = span;
      if (j.valid && j.kb >= j.ka) {
        hi = std::min(j.ka >> 1, span);
        lo = std::max(0, (j.kb + 1 - 4 * J.nq[s] + 1) >> 1);
        if (lo > hi) lo = hi;
      }
    };
    const int unit = (J.split_seg == s) ? 2 : 1;
    for (int w = 0; w < 2; ++w) {
      const int base = s * 64 + 32 * w;
      std::vector<int> order(32 / unit);
      for (int u = 0; u < 32 / unit; ++u) order[u] = u;
      auto n_feasible = [&](int u) {
        int n = 0;
        for (int e = 0; e < unit; ++e) {
          int lo, hi;
          range_of(jobs[base + u * unit + e], lo, hi);
          n += std::min(8, hi - lo + 1);
        }
        return n;
      };
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n_feasible(a) < n_feasible(b); });
      int members[4] = {0, 0, 0, 0};
      bool used[4][8] = {};
      Job placed[32];
      int placed_c0[32];
      int slot_of_group[4][8];
      for (int u : order) {
        int pick[2] = {0, 0}, group = -1;
        for (int g = 0; g < 4 && group < 0; ++g) {
          if (members[g] + unit > 8) continue;
          bool taken[8];
          for (int r = 0; r < 8; ++r) taken[r] = used[g][r];
          bool ok = true;
          for (int e = 0; e < unit && ok; ++e) {
            int lo, hi;
            range_of(jobs[base + u * unit + e], lo, hi);
            int found = -1;
            for (int c = hi; c >= lo && c > hi - 8; --c)
              if (!taken[c & 7]) {
                found = c;
                break;
              }
            if (found < 0) ok = false;
            else {
              taken[found & 7] = true;
              pick[e] = found;
            }
          }
          if (ok) group = g;
        }
        if (group < 0) {  // no conflict-free place left: the emptiest quarter-warp, at the band's own first chunk
          for (int g = 0; g < 4; ++g)
            if (members[g] + unit <= 8 && (group < 0 || members[g] < members[group])) group = g;
          for (int e = 0; e < unit; ++e) {
            int lo, hi;
            range_of(jobs[base + u * unit + e], lo, hi);
            pick[e] = hi;
          }
        }
        for (int e = 0; e < unit; ++e) {
          used[group][pick[e] & 7] = true;
          slot_of_group[group][members[group]] = u * unit + e;
          const int lane = 8 * group + members[group]++;
          placed[lane] = jobs[base + u * unit + e];
          placed_c0[lane] = pick[e];
        }
      }
      (void)slot_of_group;
      for (int l = 0; l < 32; ++l) {
        jobs[base + l] = placed[l];
        c0_of[base + l] = placed_c0[l];
      }
    }
  }
  J.w.assign((size_t)std::max(J.total_quads, 1) * 64 * 4, 0.0f);
  J.desc.assign(kMelMaxSeg * 64, 0);
  int qbase = 0;
  for (int s = 0; s < J.nseg; ++s) {
    for (int t = 0; t < 64; ++t) {
      const Job& j = jobs[s * 64 + t];
      const int c0 = c0_of[s * 64 + t];
      if (j.valid && j.kb >= j.ka) {
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
