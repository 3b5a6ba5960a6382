// Path C, streaming form: stored spectrogram [B][K][T] (t contiguous) -> mel image [B][M][T][C], FP32, sm_100a.
//
// Replaces tf.tensordot(MEL_WEIGHTS, spectrogram(2049x513), 1) (tfdataset.py:1082-1090; magnitude in, Q6) and, with
// power 2, custommel.mel_spec (custommel.py:57-61).  HBM bound: rows bin_lo..bin_hi of a clip (930 x 513 x 4 B) are ONE
// contiguous 1.9 MB block, so the kernel streams it as whole-row chunks with TMA bulk copies instead of the 412..512-byte
// row pieces at a 2052-byte pitch that the per-column kernels (k_melspec.cuh, k_melspec_tc.cuh) fetch:
//   * work item = (clip, band segment); a persistent CTA walks its items; a segment is the run of rows that feed bands
//     [m0, m1) (the plan cuts the bank into 1, 2 or 4 segments of about equal row count; rows of the one band that
//     straddles a cut are read by both sides);
//   * a producer lane issues one cp.async.bulk per chunk of 8 rows (16.4 KB for T = 513) into a ring of shared-memory
//     stages (full / empty mbarriers), two CTAs per SM with five stages each.  A chunk starts wherever row k of clip b
//     starts -- 4-byte aligned only -- so the copy starts at the 16-byte boundary below it and the consumers index with the 0..3 float offset.  A chunk whose
//     rounded copy would leave the caller's array (first / last bytes of the tensor) is loaded with plain LDGs instead;
//   * six consumer warps own the time columns (thread c -> t = c, c + 192, ...).  A triangular bank feeds each bin to at
//     most two adjacent bands, so a thread keeps two accumulators per column (band `lo` on its falling slope, band
//     `lo + 1` on its rising slope); the per-row record (w_lo, w_hi, lo) is one broadcast LDS.128.  When `lo` advances
//     the finished band is stored (coalesced along t) and the pair shifts.  Accumulation runs along k in FP32, the
//     order of the per-column kernel.
// Banks that are not of that form (a bin feeding three bands or two non-adjacent ones), T > 768 and the [B][T][M] layout
// stay on melspec_banded_kernel.
#pragma once
#include "cacfe_common.cuh"
#include "cacfe_async.cuh"

namespace cacfe {

constexpr int kMsConsumers = 192;
constexpr int kMsThreads = kMsConsumers + 32;
constexpr int kMsRows = 8;         // rows per chunk
constexpr int kMsMaxCols = 4;      // columns per thread: T <= 768
constexpr int kMsMaxStages = 8;
constexpr int kMsMaxSegs = 4;

struct MelStreamSeg {
  int row0, row1;   // spectrogram rows [row0, row1) feed the segment's bands
  int m0, m1;       // bands [m0, m1) are stored by this segment
};

struct MelStreamArgs {
  const float* spec;        // [B][n_bins][T]
  float* out;               // [B][M][T][C]
  const float4* rows;       // per row bin_lo + i: (w_lo, w_hi, lo as int bits, 0)
  MelStreamSeg segs[kMsMaxSegs];
  int n_segs, n_items;      // items = B * n_segs
  int n_rows, bin_lo, n_bins, T, n_mels, channels;
  int stages, stage_floats; // ring geometry (host: melstream_geometry)
  size_t total_bytes;       // B * n_bins * T * 4: bulk copies stay inside [spec, spec + total_bytes)
};

struct MelStreamGeom {
  int stages, stage_floats;
  size_t smem_bytes;
};

inline MelStreamGeom melstream_geometry(int T, int n_rows, size_t smem_limit) {
  MelStreamGeom g;
  g.stage_floats = (kMsRows * T + 8 + 31) & ~31;               // + 8: alignment offset and the rounded-up copy length
  const size_t fixed = (size_t)n_rows * sizeof(float4) + 2 * kMsMaxStages * sizeof(uint64_t) +
                       (size_t)kMsConsumers * kMsMaxCols * sizeof(float);  // tail slack for the unguarded column reads
  const size_t per_stage = (size_t)g.stage_floats * sizeof(float);
  long long s = smem_limit > fixed ? (long long)((smem_limit - fixed) / per_stage) : 0;
  g.stages = (int)(s > kMsMaxStages ? kMsMaxStages : s);
  g.smem_bytes = fixed + (size_t)g.stages * per_stage;
  return g;
}

// Is the 16-byte-rounded copy of [p, p + bytes) inside [base, base + total)?
__device__ __forceinline__ bool ms_bulk_ok(const float* p, uint32_t bytes, const float* base, size_t total) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p), b0 = reinterpret_cast<uintptr_t>(base);
  const uintptr_t lo = a & ~uintptr_t(15), hi = (a + bytes + 15) & ~uintptr_t(15);
  return lo >= b0 && hi <= b0 + total;
}

template <int C, bool POW2, int CH>   // CH: channel repeat 1 or 3 unrolled, 0 = a.channels
__global__ void __launch_bounds__(kMsThreads, 2) melspec_stream_kernel(const MelStreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* s_stage = reinterpret_cast<float*>(smem);
  float4* s_rows = reinterpret_cast<float4*>(s_stage + (size_t)a.stages * a.stage_floats + kMsConsumers * kMsMaxCols);
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_rows + a.n_rows);
  uint64_t* s_empty = s_full + kMsMaxStages;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), kMsConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < a.n_rows; i += kMsThreads) s_rows[i] = a.rows[i];
  __syncthreads();

  const int T = a.T;
  if (warp == kMsConsumers / 32) {
    // ---- producer: one lane walks the same (item, chunk) sequence as the consumers ---------------------------------
    if (lane != 0) return;
    int n = 0, j = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++j) {
      const int b = item / a.n_segs;
      // the j-th item of a CTA takes segment (item + j) % n_segs of its clip: every CTA cycles through the segments (the low
      // bands end every ~3 rows and cost more), and -- the grid being a multiple of n_segs -- each (clip, segment) is taken once
      const MelStreamSeg sg = a.segs[(item - b * a.n_segs + j) % a.n_segs];
      for (int r0 = sg.row0; r0 < sg.row1; r0 += kMsRows, ++n) {
        const int s = n % a.stages;
        const uint32_t ph = (uint32_t)((n / a.stages) & 1);
        mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);
        const int nr = min(kMsRows, sg.row1 - r0);
        const float* src = a.spec + ((size_t)b * a.n_bins + r0) * T;
        const uint32_t bytes = (uint32_t)(nr * T) * 4u;
        const uint32_t bar = smem_u32(&s_full[s]);
        if (ms_bulk_ok(src, bytes, a.spec, a.total_bytes)) {
          const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
          const uint32_t head = (uint32_t)(addr & 15);
          const uint32_t len = (head + bytes + 15u) & ~15u;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(bar, len);
          bulk_g2s(smem_u32(s_stage + (size_t)s * a.stage_floats), reinterpret_cast<const void*>(addr - head), len, bar);
        } else {
          mbar_arrive(bar);   // the stage is free: the consumers fill it themselves
        }
      }
    }
    return;
  }

  // ---- consumers ---------------------------------------------------------------------------------------------------
  const int chn = CH > 0 ? CH : a.channels;
  const size_t m_stride = (size_t)T * chn;
  bool col_ok[C];
#pragma unroll
  for (int c = 0; c < C; ++c) col_ok[c] = tid + c * kMsConsumers < T;
  int n = 0, j = 0;
  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++j) {
    const int b = item / a.n_segs;
    const MelStreamSeg sg = a.segs[(item - b * a.n_segs + j) % a.n_segs];
    float acc_lo[C], acc_hi[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc_lo[c] = acc_hi[c] = 0.0f;
    int cur = sg.m0;
    if (sg.row0 < sg.row1) cur = min(cur, __float_as_int(s_rows[sg.row0 - a.bin_lo].z));
    float* o_cur = a.out + ((size_t)b * a.n_mels + cur) * m_stride + (size_t)tid * chn;   // this thread's first column of band `cur`

    // band `cur` is complete: store it if this segment owns it, then the pair moves up by one band
    auto advance = [&]() {
      if (cur >= sg.m0 && cur < sg.m1) {
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (col_ok[c]) {
            float* o = o_cur + (size_t)c * kMsConsumers * chn;
            if (CH > 0) {
#pragma unroll
              for (int ch = 0; ch < CH; ++ch) o[ch] = acc_lo[c];
            } else {
#pragma unroll 1
              for (int ch = 0; ch < chn; ++ch) o[ch] = acc_lo[c];
            }
          }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        acc_lo[c] = acc_hi[c];
        acc_hi[c] = 0.0f;
      }
      ++cur;
      o_cur += m_stride;
    };

    for (int r0 = sg.row0; r0 < sg.row1; r0 += kMsRows, ++n) {
      const int s = n % a.stages;
      const uint32_t ph = (uint32_t)((n / a.stages) & 1);
      const int nr = min(kMsRows, sg.row1 - r0);
      const float* src = a.spec + ((size_t)b * a.n_bins + r0) * T;
      float* stage = s_stage + (size_t)s * a.stage_floats;
      mbar_wait(smem_u32(&s_full[s]), ph);
      const float* x;
      if (ms_bulk_ok(src, (uint32_t)(nr * T) * 4u, a.spec, a.total_bytes)) {   // CTA-uniform
        x = stage + ((reinterpret_cast<uintptr_t>(src) & 15) >> 2);
      } else {
        for (int i = tid; i < nr * T; i += kMsConsumers) stage[i] = ld_stream(src + i);
        group_barrier(1, kMsConsumers);
        x = stage;
      }
      x += tid;
      const float4* rec = s_rows + (r0 - a.bin_lo);
      // four rows per trip: all 4 (C + 1) shared-memory loads are issued before the first band check
#pragma unroll 1
      for (int r = 0; r < nr; r += 4) {
        float4 w[4];
        float v[4][C];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = min(r + u, nr - 1);
          w[u] = rec[rr];
#pragma unroll
          for (int c = 0; c < C; ++c) v[u][c] = x[rr * T + c * kMsConsumers];   // columns t >= T read on into the next row: never stored
        }
        const int left = nr - r;
        if (left >= 4 && __float_as_int(w[3].z) == cur) {   // CTA-uniform: no band ends inside these rows (the common case)
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const float p = POW2 ? v[u][c] * v[u][c] : v[u][c];
              acc_lo[c] = fmaf(w[u].x, p, acc_lo[c]);
              acc_hi[c] = fmaf(w[u].y, p, acc_hi[c]);
            }
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (u < left) {
              const int lo = __float_as_int(w[u].z);
              while (cur < lo) advance();
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float p = POW2 ? v[u][c] * v[u][c] : v[u][c];
                acc_lo[c] = fmaf(w[u].x, p, acc_lo[c]);
                acc_hi[c] = fmaf(w[u].y, p, acc_hi[c]);
              }
            }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));
    }
    while (cur < sg.m1) advance();
  }
}

}  // namespace cacfe
