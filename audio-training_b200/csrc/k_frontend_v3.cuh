// K1 (persistent form, round-1 third generation): fused normalise / frame / Hann / rFFT / power / mel for sm_100a.
//
// Same arithmetic as stft_mel_kernel (k_frontend.cuh).  What changed, and the ncu evidence for it
// (profiles/r01_k1_v2_*: stft_mel_stream_kernel, 3.70 ms for 1024 clips, IPC 2.03):
//   * the largest stall was `no_instruction` (1.76 of 5.92 warp-cycles per issue): that kernel is ~6 300 SASS
//     instructions (100 KB) of straight-line code with 12 warps at 6 unrelated positions -- it does not fit any
//     level of the instruction cache.  Here the two 64-point FFT passes of a frame pair run through ONE copy of the
//     unrolled cacfe_fft64 (a two-trip loop around it), there is one stage-1 load path instead of two, and the mel
//     loop is rolled: the hot loop is ~40 % of the old size;
//   * the per-clip normalisation is applied once per staged sample, in place in shared memory (x - min is exact
//     for clips whose DC dwarfs their range, which the spectrum-side trick of the previous kernel was not), and the
//     same pass writes the zero padding (Q4: padding happens after normalisation) or the reflected samples, so every
//     frame -- partial or not -- takes the same code path: 2 LDS + 1 LDS (window) + 2 FMUL per complex point;
//   * a tile is 12 frames = 6 frame pairs = one pair per 64-thread FFT group per trip, the CTA meets at two
//     barriers per tile (buffer hand-over), so all 12 warps stay within one trip of each other and share
//     instruction-cache lines;
//   * the mel projection is balanced across the 64 threads of a group (serpentine band order, the short last
//     round split across lane pairs) -- the old order left one warp with 48 taps and the other with 22.
// Shorter transforms ride the same kernel: a frame of n_fft = 4096 / r samples, zero padded to 4096, has its DFT at every
// r-th bin of the 4096-point one, so the plan zero-pads the window and places the filterbank weights on bins r k
// (tf.signal.stft(1024 | 2048, ...) of raw_to_mel_rgb / raw_to_mel_dual, tfdataset.py:1818-2004).  r x the FFT work of a
// dedicated kernel; those variants are not on the benchmarked path.
// Sample tiles still arrive by TMA bulk copy (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) into a 2-deep ring.
//
// Round 2 (DESIGN.md 3.1a'): what bound this kernel was again the instruction fetch path (47.6 KB of loop body against a 32 KB
// instruction cache) and spilled registers (168 per thread, and with 220 KB of shared memory no L1 to catch a spill).  Three
// changes without any arithmetic in them took it from 9.67 to 8.2 ms per 4096 clips: the FFT-group index goes through a lane-0
// broadcast so that ptxas keeps everything derived from it in uniform registers; the benchmarked configurations have their own
// instantiations (template parameter HOT) with the run-time switches resolved; the library is built with ptxas
// --register-usage-level=7.  The kernel is sensitive to both: any edit must be checked with -Xptxas -v (spills) and timed
// (tools/ab_k1.py) -- shorter source is not faster code here.
#pragma once
#include "cacfe_common.cuh"
#include "frontend_core.cuh"
#include "k_frontend.cuh"
#include "cacfe_async.cuh"  // mbarrier / bulk-copy wrappers
#include "fft64x2_gen.cuh"
#include "mel_jobs.h"
#include "win_consts.cuh"

namespace cacfe {

#ifndef CACFE_VGROUPS          // FFT groups (64 threads each) per CTA; a tile is 2 * groups frames.  A/B switch: tools/ab_k1.py
#define CACFE_VGROUPS 6
#endif
constexpr int kVGroups = CACFE_VGROUPS;
constexpr int kHalfStride = 68;                    // floats per half-exchange row: 272 B = 2*128 + 16
constexpr int kHalfFloats = 64 * kHalfStride;      // 17408 B per group
constexpr int kVThreads = kVGroups * 64;
constexpr int kVTileFrames = 2 * kVGroups;
#ifndef CACFE_K1_NORM_ITERS
#define CACFE_K1_NORM_ITERS 6
#endif
constexpr int kNormIters = CACFE_K1_NORM_ITERS;
#ifndef CACFE_K1_NORM_UNROLL   // A/B switch; rolled (1) is 112 instructions shorter than fully unrolled (6) and 0.9 % faster
#define CACFE_K1_NORM_UNROLL 1
#endif
constexpr int kNormUnroll = CACFE_K1_NORM_UNROLL;    // 16-byte groups per thread in the normalise pass: tiles up to 6 * 384 * 4 samples (hop <= 464)

struct VSmem {
  int tile_len, tile_pad, mel_quads;
  size_t off_win, off_tile, off_exch, off_melw, off_desc, off_sync, total;
};

__host__ __device__ inline VSmem v3_smem_layout(int hop, int mel_quads) {
  VSmem s;
  s.mel_quads = mel_quads;
  s.tile_len = kFft + hop * (kVTileFrames - 1);
  s.tile_pad = (s.tile_len + 3 + 4) & ~3;          // room for the copy length rounded up to 16 B
  size_t o = sizeof(float2) * 4096;
  s.off_win = o;    o += sizeof(float) * kFft;      // full window: Hann(n_fft) zero padded to 4096 (n_fft < 4096)
  s.off_tile = o;   o += sizeof(float) * s.tile_pad * 2;
  s.off_exch = o;   o += sizeof(float) * kHalfFloats * kVGroups;
  s.off_melw = o;   o += sizeof(float4) * 64 * (mel_quads > 0 ? mel_quads : 1);
  s.off_desc = o;   o += sizeof(int) * 64 * kMelMaxSeg;
  s.off_sync = o;   o += 96;                         // full/norm/pre mbarriers, done counters, normalisation pairs
  s.total = o;
  return s;
}

#ifdef CACFE_K1_JITTER
// Race hunting without compute-sanitizer (tests/test_gpu_parity.py::test_k1_jitter): a pseudo-random pause before every
// hand-over operation.  A schedule-dependent result shows up as a run that differs from the others.
__device__ __forceinline__ void k1_jitter(unsigned salt) {
  unsigned h = (unsigned)clock() * 2654435761u + salt * 40503u + threadIdx.x * 97u + blockIdx.x * 1315423911u;
  h ^= h >> 15;
  __nanosleep(h & 0x7ffu);   // 0 .. 2 us
}
#else
__device__ __forceinline__ void k1_jitter(unsigned) {}
#endif

// LAYOUT_SPECT store phase: every thread takes one frame tt of the tile and the rows k0, k0 + 32, ...; twelve neighbouring lanes
// write one 48-byte run.  A function of its own, called once per tile when the FFT's registers are dead, so that ptxas allocates
// it apart from the loop (inlined: 92 / 108 bytes of spill traffic in the loop and 4.16 ms per 1024 clips; called: 20 / 40, 3.92 ms).
__device__ __noinline__ void k1_spect_store(const float* s_exch, float* obase, int n_frames, int n_out, int rsh, int n_t, int tid) {
  __syncthreads();
  const int tt = tid % kVTileFrames, k0 = tid / kVTileFrames;
  const float* src = s_exch + (tt >> 1) * kHalfFloats + (tt & 1);
  float* dst = obase + tt;
  if (tt < n_t) {
#pragma unroll 4
    for (int ko = k0; ko < n_out; ko += kVThreads / kVTileFrames) dst[(size_t)ko * n_frames] = src[2 * (ko << rsh)];
  }
  __syncthreads();
}

// Mel job tables of the plan (mel_jobs.h), device copies.
struct MelArgs {
  const float4* tw4;  // [32][64] stage twiddles, packed per output pair: (cos k, cos k+1, sin k, sin k+1)
  const float* win;   // [4096] periodic Hann of length n_fft, zero beyond n_fft
  const float4* w;    // [total_quads][64]
  const int* desc;    // [kMelMaxSeg][64]
  int nq[kMelMaxSeg];
  int split_seg, total_quads;
  int spec_ratio, spec_bins;   // LAYOUT_SPEC: 4096 / n_fft and n_fft / 2 + 1
  int tile_len, tile_pad;      // v3_smem_layout's tile geometry (read from the constant bank in the loop instead of re-derived)
};

// WINC: the Hann(4096) window is computed per thread by angle addition (two FFMA with immediates per value) instead of being
// read from shared memory (one LDS.64 per two values): the kernel is shared-memory-wavefront bound, not FMA bound.  Shorter
// transforms (zero-padded window) keep the table.
// HOT = 1: the instantiation of the benchmarked path (per-clip normalisation on, no reflect padding, power 2; for the spectrogram
// layout power 1 = the stored magnitude of audiodataset.load_data) with those three run-time
// switches resolved at compile time: the magnitude loop with its sqrt calls, the mirror pass and the un-normalised form leave
// the code the twelve warps fetch.
template <int NQ, int LAYOUT, bool WINC = false, int HOT = 0>   // HOT: 0 generic, 1 normalisation on, 2 normalisation off
__global__ void __launch_bounds__(kVThreads, 1) stft_mel_v3_kernel(const FrontendArgs a, const MelArgs mj,
                                                                      const int total_tiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const VSmem L = v3_smem_layout(a.hop, mj.total_quads);
  // (HOT = 2: the same with the normalisation off -- raw_to_mel / get_spect on clips the caller has normalised, tfdataset.py:913-915)
  const bool has_norm = HOT == 1 || (HOT == 0 && a.norm != nullptr), reflect = !HOT && a.reflect, magnitude = !HOT && a.power == 1;
  const bool spec_magnitude = HOT || a.power == 1;   // LAYOUT_SPEC only
  float4* s_tw4 = reinterpret_cast<float4*>(smem);   // [32 output pairs][64 n2]
  float2* s_win2 = reinterpret_cast<float2*>(smem + L.off_win);
  float* s_tile = reinterpret_cast<float*>(smem + L.off_tile);
  float* s_exch = reinterpret_cast<float*>(smem + L.off_exch);
  float4* s_melw = reinterpret_cast<float4*>(smem + L.off_melw);
  int* s_desc = reinterpret_cast<int*>(smem + L.off_desc);
  uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + L.off_sync);  // [2]

  uint64_t* s_norm = s_full + 2;                                      // [2] tile normalised (12 warp arrivals)
  uint64_t* s_pre = s_full + 4;                                       // [2] reflect framing: first pass done
  int* s_done = reinterpret_cast<int*>(s_full + 6);                   // [2] groups that have read the tile
  float2* s_nrm = reinterpret_cast<float2*>(s_full + 8);              // [2][2] (range, min) of clips b & ~1, b | 1

#ifdef CACFE_K1_WARP_PERM   // A/B switch (tools/ab_k1.py): put both warps of FFT groups 0..3 on one scheduler (warps w, w + 4)
  const int pw_ = (int)(threadIdx.x >> 5);
  const int lw_ = pw_ < 8 ? ((pw_ & 3) * 2 + (pw_ >> 2)) : pw_;
  const int tid = lw_ * 32 + (int)(threadIdx.x & 31);
#else
  const int tid = threadIdx.x;
#endif
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // The group index goes through a lane-0 broadcast: ptxas then knows it is warp-uniform and keeps everything derived from it
  // (frame pointers, the exchange tile, frame numbers, store predicates) in uniform registers.  Measured (tools/ab_k1.py, 4096
  // clips): 4096 -> 3672 SASS instructions, 64 / 124 bytes of spill stores / loads -> none, 9.66 -> 9.14 ms, bit-identical.
#ifdef CACFE_K1_VECTOR_G   // A/B switch: the plain per-thread form
  const int g = tid >> 6;
#else
  const int g = __shfl_sync(kFullMask, tid >> 6, 0);
#endif
  const int t64 = tid & 63, lane = tid & 31;

  // Arms buffer i&1 with this CTA's i-th tile: one bulk copy of the samples that lie inside the clip, one of the
  // clip's normalisation pair.
  auto issue_tile = [&](int i) {
    const int s = i & 1;
    const int w = (int)blockIdx.x + i * (int)gridDim.x;
    const int b = w / a.tiles_per_clip;
    const int t0 = (w - b * a.tiles_per_clip) * kVTileFrames;
    const int s_lo = a.origin + a.hop * t0;          // multiple of 4: hop * 12 and the origin both are
    const int c0 = max(s_lo, 0);
    int c1 = min(s_lo + L.tile_len, a.n_samples);
    c1 = (c1 + 3) & ~3;                              // n_samples % 4 == 0 on this path, so this never leaves the clip
    const uint32_t bytes = (uint32_t)(c1 - c0) * 4u;
    const uint32_t bar = smem_u32(&s_full[s]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses vs the async write
    mbar_expect_tx(bar, bytes + (has_norm ? 16u : 0u));
    bulk_g2s(smem_u32(s_tile + (size_t)s * L.tile_pad + (c0 - s_lo)), a.in + (size_t)b * a.n_samples + c0, bytes, bar);
    if (has_norm) bulk_g2s(smem_u32(s_nrm + 2 * s), a.norm + (b & ~1), 16u, bar);
  };

  // Every warp normalises its slice of tile i in place (and writes the padding), then arrives on s_norm[i&1].
  auto normalise_tile = [&](int i, int b, int t0) {
    const int s = i & 1;
    const int s_lo = a.origin + a.hop * t0;
    float* tile = s_tile + (size_t)s * L.tile_pad;
    const uint32_t parity = (uint32_t)((i >> 1) & 1);
    mbar_wait_relaxed(smem_u32(&s_full[s]), parity);
    float mn = 0.0f, sc = 1.0f, of = 0.0f;
    if (has_norm) {  // ((x - mn) / range + 1e-6 - 0.5) * 2, the reference's order with one rounding less
      const float2 nrm = s_nrm[2 * s + (b & 1)];
      mn = nrm.y;
      sc = 2.0f / nrm.x;      // range 0 -> inf -> (x - mn) * inf = NaN: constant clips give NaN features (Q1)
      of = -0.999998f;
    }
    // s_lo and n_samples are multiples of 4: a 16-byte group of samples lies entirely inside or outside the clip
    float4* t4 = reinterpret_cast<float4*>(tile);
    const int n4 = L.tile_pad >> 2;
    const int e_lo = s_lo < 0 ? (-s_lo) >> 2 : 0, e_hi = (a.n_samples - s_lo) >> 2;
    const cacfe_f2 mn2 = cacfe_pk(mn, mn), sc2 = cacfe_pk(sc, sc), of2 = cacfe_pk(of, of);
#ifdef CACFE_K1_NORM_ONEPATH   // A/B switch: one unrolled path for every tile (a select per 16-byte group), no second loop
#pragma unroll
    for (int u = 0; u < kNormIters; ++u) {
      const int e = tid + u * kVThreads;
      if (e < n4) {
        const float4 v = t4[e];
        const cacfe_f2 lo = cacfe_fma2(cacfe_sub2(cacfe_pk(v.x, v.y), mn2), sc2, of2);
        const cacfe_f2 hi = cacfe_fma2(cacfe_sub2(cacfe_pk(v.z, v.w), mn2), sc2, of2);
        const bool inside = e >= e_lo && e < e_hi;
        t4[e] = inside ? make_float4(cacfe_lo(lo), cacfe_hi(lo), cacfe_lo(hi), cacfe_hi(hi)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
    }
#else
    if (e_lo == 0 && e_hi >= n4) {  // the whole tile lies inside the clip (41 of 43 tiles): no padding to write
      // (measured alternatives, tools/ab_k1.py: all six loads before the first store, a software pipeline of depth 1 - 3, the
      // loop as a function of its own -- each is 4 - 8 % slower: more instructions in the loop body, or a stack frame)
#pragma unroll kNormUnroll
      for (int u = 0; u < kNormIters; ++u) {
        const int e = tid + u * kVThreads;
        if (e < n4) {
          const float4 v = t4[e];
          const cacfe_f2 lo = cacfe_fma2(cacfe_sub2(cacfe_pk(v.x, v.y), mn2), sc2, of2);
          const cacfe_f2 hi = cacfe_fma2(cacfe_sub2(cacfe_pk(v.z, v.w), mn2), sc2, of2);
          t4[e] = make_float4(cacfe_lo(lo), cacfe_hi(lo), cacfe_lo(hi), cacfe_hi(hi));
        }
      }
    } else {
#pragma unroll 1
      for (int e = tid; e < n4; e += kVThreads) {
        const float4 v = t4[e];
        const cacfe_f2 lo = cacfe_fma2(cacfe_sub2(cacfe_pk(v.x, v.y), mn2), sc2, of2);
        const cacfe_f2 hi = cacfe_fma2(cacfe_sub2(cacfe_pk(v.z, v.w), mn2), sc2, of2);
        const bool inside = e >= e_lo && e < e_hi;
        t4[e] = inside ? make_float4(cacfe_lo(lo), cacfe_hi(lo), cacfe_lo(hi), cacfe_hi(hi))
                       : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
    }
#endif
    if (reflect) {  // numpy 'reflect' (no edge repeat): copy the already normalised mirror samples
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_pre[s]));
      mbar_wait(smem_u32(&s_pre[s]), parity);
      for (int e = tid; e < L.tile_len; e += kVThreads) {
        const int p = s_lo + e;
        int r = p;
        if (p < 0) r = -p;
        if (p >= a.n_samples) r = 2 * (a.n_samples - 1) - p;
        if (r != p && r >= s_lo && r >= 0 && r < a.n_samples) tile[e] = tile[r - s_lo];  // host checked: true for samples a frame reads
      }
    }
    __syncwarp();
    k1_jitter(5u * (unsigned)i + 2u);
    if (lane == 0) mbar_arrive(smem_u32(&s_norm[s]));
  };

  // A group has read tile i into registers; the last of the six re-arms the buffer with tile i + 2.
  auto release_tile = [&](int i) {
    group_barrier(1 + g, 64);
    if (t64 == 0) {
      const int s = i & 1;
      k1_jitter(5u * (unsigned)i + 1u);
      __threadfence_block();
      const int old = atomicAdd(&s_done[s], 1);
      if (old == kVGroups - 1) {
        __threadfence_block();
        s_done[s] = 0;
        if (i + 2 < my_tiles) issue_tile(i + 2);
      }
    }
  };

  if (tid == 0) {
    for (int q = 0; q < 2; ++q) {
      mbar_init(smem_u32(&s_full[q]), 1);
      mbar_init(smem_u32(&s_norm[q]), kVThreads / 32);
      mbar_init(smem_u32(&s_pre[q]), kVThreads / 32);
      s_done[q] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- tables, once per CTA (L2 resident) ------------------------------------------------------------------
  {
    for (int i = tid; i < 2048; i += kVThreads) s_tw4[i] = mj.tw4[i];
    // window pre-paired for the packed multiply: s_win2[(q / 2) * 64 + t] = (w[64 q + t], w[64 (q + 1) + t]), q even
    for (int i = tid; i < kFft / 2; i += kVThreads) {
      const int qh = i >> 6, t = i & 63;
      s_win2[i] = make_float2(mj.win[128 * qh + t], mj.win[128 * qh + 64 + t]);
    }
    for (int i = tid; i < 64 * mj.total_quads; i += kVThreads) s_melw[i] = mj.w[i];
    for (int i = tid; i < 64 * kMelMaxSeg; i += kVThreads) s_desc[i] = mj.desc[i];
  }
  __syncthreads();
  if (tid == 0) {
    if (my_tiles > 0) issue_tile(0);
    if (my_tiles > 1) issue_tile(1);
  }
  // tile coordinates (clip, first frame) advance by gridDim.x tiles per trip: kept incrementally, no division in the loop
  const int step_b = (int)gridDim.x / a.tiles_per_clip, step_t = ((int)gridDim.x % a.tiles_per_clip) * kVTileFrames;
  const int wrap_t = a.tiles_per_clip * kVTileFrames;
  int b = (int)blockIdx.x / a.tiles_per_clip, t0 = ((int)blockIdx.x % a.tiles_per_clip) * kVTileFrames;
  if (my_tiles > 0) normalise_tile(0, b, t0);

  float* ex = s_exch + g * kHalfFloats;
  float2* pbuf = reinterpret_cast<float2*>(ex);  // (power A, power B) per bin; aliases the exchange tile
  const int j = stage2_row(t64);
  const bool self = (j == 0) || (j == 32);
  const int plane = self ? lane : (lane ^ 16);
  // The six groups run free: there is no CTA-wide barrier in the loop, only the two mbarrier hand-overs per tile.
  for (int i = 0; i < my_tiles; ++i) {
    const int s = i & 1;
    float* tile = s_tile + (size_t)s * L.tile_pad;
    int b_next = b + step_b, t_next = t0 + step_t;
    if (t_next >= wrap_t) {
      t_next -= wrap_t;
      ++b_next;
    }

    if (i + 1 < my_tiles) normalise_tile(i + 1, b_next, t_next);   // buffer s^1: its TMA was issued when tile i-1 had been read by all
    mbar_wait_relaxed(smem_u32(&s_norm[s]), (uint32_t)((i >> 1) & 1));

    const int ta = t0 + 2 * g;
    const int b_cur = b;
    b = b_next;
    t0 = t_next;
    // LAYOUT_SPECT: the six groups leave (|XA|, |XB|) of their frame pairs in their exchange tiles [k], the CTA meets, and all
    // threads write the tile's 12 frames as 48-byte runs of the stored [b][k][t] array (the neighbouring tiles of the clip are in
    // flight on the neighbouring CTAs: L2 sees whole lines); a second rendezvous hands the exchange tiles back to the FFT.
    auto spect_store = [&]() {
      const int t_first = ta - 2 * g;
      k1_spect_store(s_exch, a.out + (size_t)b_cur * mj.spec_bins * a.n_frames + t_first, a.n_frames, mj.spec_bins,
                     31 - __clz(mj.spec_ratio), min(kVTileFrames, a.n_frames - t_first), tid);
    };
    // (LAYOUT_SPECT: a group whose frames lie past the end of the clip -- one or two groups in the last of a clip's 43 tiles --
    // transforms whatever its part of the tile holds and stores nothing: every thread reaches the one spect_store below)
    if (LAYOUT != LAYOUT_SPECT && ta >= a.n_frames) {  // group-uniform
      release_tile(i);
      continue;
    }
    const bool store_b = ta + 1 < a.n_frames;

    float re[64], im[64];
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
      if (ph == 0) {
        // ---- stage 1 input: window; z[n] = w[n] (xA[n] + i xB[n]), thread n2 = t64 holds n = 64 q + n2 ----------------
        const float* fa = tile + (2 * g) * a.hop;
        const float* fb = fa + a.hop;
        float win_c = 0.0f, win_s = 0.0f;   // (cos, sin) of phi_t = 2 pi t / 4096: re-read per pair (two registers that must not
        if (WINC) {                         // stay live across the DFT) from the twiddle row of k = 1: (cos, -sin)(2 pi n2 / 4096)
          const float4 t1 = s_tw4[t64];
          win_c = t1.y;
          win_s = -t1.w;
        }
#ifndef CACFE_K1_WIN_FULL   // A/B switch: compute all 64 window values per thread
        if (WINC) {
          // Hann: w[n + 2048] = 1 - w[n].  Thirty-two window values per thread (two FFMAs each), the other thirty-two products as
          // x - x w: one packed FMA where the plain form has a packed multiply, and 64 scalar FFMAs fewer per frame pair.
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const int n = 64 * q + t64;
            const cacfe_f2 wv = cacfe_pk(fmaf(kWinA[q], win_c, fmaf(kWinB[q], win_s, 0.5f)),
                                         fmaf(kWinA[q + 1], win_c, fmaf(kWinB[q + 1], win_s, 0.5f)));
            const cacfe_f2 xa = cacfe_mul2(cacfe_pk(fa[n], fa[n + 64]), wv), xb = cacfe_mul2(cacfe_pk(fb[n], fb[n + 64]), wv);
            const cacfe_f2 ua = cacfe_pk(fa[n + 2048], fa[n + 2112]), ub = cacfe_pk(fb[n + 2048], fb[n + 2112]);
            const cacfe_f2 ya = cacfe_sub2(ua, cacfe_mul2(ua, wv)), yb = cacfe_sub2(ub, cacfe_mul2(ub, wv));
            re[q] = cacfe_lo(xa);
            re[q + 1] = cacfe_hi(xa);
            im[q] = cacfe_lo(xb);
            im[q + 1] = cacfe_hi(xb);
            re[q + 32] = cacfe_lo(ya);
            re[q + 33] = cacfe_hi(ya);
            im[q + 32] = cacfe_lo(yb);
            im[q + 33] = cacfe_hi(yb);
          }
        } else
#endif
#pragma unroll
        for (int q = 0; q < 64; q += 2) {  // packed: the pair (q, q + 1) is also the input pair of cacfe_fft64x2
          const int n = 64 * q + t64;
          float2 w2;
          if (WINC) {
            w2.x = fmaf(kWinA[q], win_c, fmaf(kWinB[q], win_s, 0.5f));
            w2.y = fmaf(kWinA[q + 1], win_c, fmaf(kWinB[q + 1], win_s, 0.5f));
          } else {
            w2 = s_win2[(q >> 1) * 64 + t64];
          }
          const cacfe_f2 wv = cacfe_pk(w2.x, w2.y);
          const cacfe_f2 xa = cacfe_mul2(cacfe_pk(fa[n], fa[n + 64]), wv);
          const cacfe_f2 xb = cacfe_mul2(cacfe_pk(fb[n], fb[n + 64]), wv);
          re[q] = cacfe_lo(xa);
          re[q + 1] = cacfe_hi(xa);
          im[q] = cacfe_lo(xb);
          im[q + 1] = cacfe_hi(xb);
        }
        release_tile(i);
      }
      cacfe_fft64x2(re, im);
      if (ph == 0) {
        // ---- twiddle W4096^(n2 k1), then the transpose through shared memory, real parts first ------------------------
        // (measured alternative: a two-level table, W^(n2 k1) = W^(n2 8 a) W^(n2 b) with the four b-pairs in registers -- 11
        // LDS.128 instead of 32 per thread and pass, one more packed complex product per output pair, four FFT groups for the
        // registers: 10.15 ms against 9.88 ms per 4096 clips.  The extra FMA work costs what the saved wavefronts gain.)
#pragma unroll
        for (int k = 0; k < 64; k += 2) {  // outputs k, k + 1 leave cacfe_fft64x2 in one register pair
          const int s0 = k, s1 = k + 1;
          const float4 t = s_tw4[(k >> 1) * 64 + t64];  // (cos k, cos k+1, sin k, sin k+1) of -2 pi k n2 / 4096
          const cacfe_f2 tr = cacfe_pk(t.x, t.y), ti = cacfe_pk(t.z, t.w);
          const cacfe_f2 zr = cacfe_pk(re[s0], re[s1]), zi = cacfe_pk(im[s0], im[s1]);
          const cacfe_f2 yr = cacfe_sub2(cacfe_mul2(zr, tr), cacfe_mul2(zi, ti));
          const cacfe_f2 yi = cacfe_fma2(zr, ti, cacfe_mul2(zi, tr));
          re[s0] = cacfe_lo(yr);
          re[s1] = cacfe_hi(yr);
          im[s0] = cacfe_lo(yi);
          im[s1] = cacfe_hi(yi);
        }
#pragma unroll
        for (int k1 = 0; k1 < 64; ++k1) ex[k1 * kHalfStride + t64] = re[k1];
        group_barrier(1 + g, 64);
        {
          const float4* row = reinterpret_cast<const float4*>(ex + j * kHalfStride);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 v = row[q];
            re[4 * q] = v.x;
            re[4 * q + 1] = v.y;
            re[4 * q + 2] = v.z;
            re[4 * q + 3] = v.w;
          }
        }
        group_barrier(1 + g, 64);
        // re[] now holds stage-2 inputs while im[] still holds stage-1 outputs
#pragma unroll
        for (int k1 = 0; k1 < 64; ++k1) ex[k1 * kHalfStride + t64] = im[k1];
        group_barrier(1 + g, 64);
        {
          const float4* row = reinterpret_cast<const float4*>(ex + j * kHalfStride);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 v = row[q];
            im[4 * q] = v.x;
            im[4 * q + 1] = v.y;
            im[4 * q + 2] = v.z;
            im[4 * q + 3] = v.w;
          }
        }
        group_barrier(1 + g, 64);  // the exchange tile may now be overwritten with powers
      }
    }

    if (LAYOUT == LAYOUT_SPEC || LAYOUT == LAYOUT_SPECT) {
      // ---- spectrogram output (audiodataset.load_data, audiodataset.py:1302-1303): |X| (power 1) or |X|^2 of every bin
      // of an n_fft-point transform; bins of the 4096-point grid that are not multiples of `ratio` belong to no bin of
      // the shorter transform.
      const bool row0 = j == 0;
      const int n_out = mj.spec_bins;
      const int rsh = 31 - __clz(mj.spec_ratio), rmask = mj.spec_ratio - 1;   // 4096 / n_fft is a power of two (plan: n_fft 512..4096)
      // staged as [b][t][k] (k contiguous: the 64 threads of a group write 256-byte runs); spec_transpose_kernel turns it
      // into the stored [b][k][t].  A direct [k][t] store writes 8-byte pieces to 2049 different rows per pair (measured:
      // 11.8 ms instead of 2.6 ms per 1024 clips).
      float* obase = a.out + ((size_t)b_cur * a.n_frames + ta) * n_out;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float sr = __shfl_sync(kFullMask, re[63 - (q & 31)], plane);
        const float si = __shfl_sync(kFullMask, im[63 - (q & 31)], plane);
        if (q < 32) {
          const float pr = row0 ? re[(64 - q) & 63] : sr;
          const float pi = row0 ? im[(64 - q) & 63] : si;
          const float zr = re[q], zi = im[q];
          const float ar = zr + pr, ai = zi - pi, br = zr - pr, bi = zi + pi;
          float va = fmaf(ar, ar, ai * ai), vb = fmaf(br, br, bi * bi);   // 4 |XA|^2, 4 |XB|^2
          if (spec_magnitude) {
            va = 0.5f * sqrt_approx(va);
            vb = 0.5f * sqrt_approx(vb);
          } else {
            va *= 0.25f;
            vb *= 0.25f;
          }
          const int k = j + 64 * q;
          if (LAYOUT == LAYOUT_SPECT) {
            pbuf[k] = make_float2(va, vb);
          } else if ((k & rmask) == 0) {
            float* o = obase + (k >> rsh);
            o[0] = va;
            if (store_b) o[n_out] = vb;
          }
        } else if (row0) {
          // Nyquist bin k = 2048 = row 0, q = 32: Z[2048] pairs with itself
          const float zr = re[32], zi = im[32];
          float va = 4.0f * zr * zr, vb = 4.0f * zi * zi;
          if (spec_magnitude) {
            va = 0.5f * sqrt_approx(va);
            vb = 0.5f * sqrt_approx(vb);
          } else {
            va *= 0.25f;
            vb *= 0.25f;
          }
          if (LAYOUT == LAYOUT_SPECT) {
            pbuf[2048] = make_float2(va, vb);
          } else {
            float* o = obase + (2048 >> rsh);
            o[0] = va;
            if (store_b) o[n_out] = vb;
          }
        }
      }
      if (LAYOUT == LAYOUT_SPECT) spect_store();
      continue;
    }
    // ---- split the two frames, power: bin k = j + 64 q goes to pbuf[k] as (4 |XA|^2, 4 |XB|^2) -------------------------
    // (the 1/4 of  XA = (Z[k] + conj Z[N-k]) / 2  is folded into the mel weights, mel_jobs.h)
    {
      const bool row0 = j == 0;
      float2* prow = pbuf + j;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        // Z[N-k]: thread 64-j holds it in slot 63-q; row 0 pairs with its own slot (64-q)&63
        const float sr = __shfl_sync(kFullMask, re[63 - q], plane);
        const float si = __shfl_sync(kFullMask, im[63 - q], plane);
        const float pr = row0 ? re[(64 - q) & 63] : sr;
        const float pi = row0 ? im[(64 - q) & 63] : si;
        const float zr = re[q], zi = im[q];
        const float ar = zr + pr, ai = zi - pi;   // 2 XA
        const float br = zr - pr, bi = zi + pi;   // 2i XB
        prow[64 * q] = make_float2(fmaf(ar, ar, ai * ai), fmaf(br, br, bi * bi));
      }
    }
    group_barrier(1 + g, 64);
    if (magnitude) {  // magnitude (stored-spectrogram convention, tfdataset.py:1085-1088): rolled, off the hot path
      for (int k = t64; k < 64 * NQ; k += 64) {
        const float2 v = pbuf[k];
        pbuf[k] = make_float2(2.0f * sqrtf(v.x), 2.0f * sqrtf(v.y));  // 4 |X|, same weight scale as the power case
      }
      group_barrier(1 + g, 64);
    }

    // ---- banded mel projection (mel_jobs.h), both frames of the pair per thread, straight to global -------------------
    {
      const float4* p4 = reinterpret_cast<const float4*>(pbuf);
      const float4* wq = s_melw + t64;
#ifdef CACFE_K1_MEL_ROLLED   // A/B switch: one copy of the segment body (rolled over the segments), each band stored as it completes
      float* const o_btm = a.out + ((size_t)b_cur * a.n_frames + ta) * a.n_mels;
      float* const o_img = a.out + ((size_t)b_cur * a.n_mels * a.n_frames + ta) * a.channels;
#pragma unroll 1
      for (int sg = 0; sg < kMelMaxSeg; ++sg) {
#else
      float ra[kMelMaxSeg], rb[kMelMaxSeg];
      int dsc[kMelMaxSeg];
#pragma unroll
      for (int sg = 0; sg < kMelMaxSeg; ++sg) {
#endif
#ifdef CACFE_K1_MEL_ROLLED   // (a run-time index into the parameter array would copy it to local memory)
        const int nq = sg == 0 ? mj.nq[0] : (sg == 1 ? mj.nq[1] : mj.nq[2]);
#else
        const int nq = mj.nq[sg];                      // uniform
#endif
        const int d = s_desc[sg * 64 + t64];
        const float4* pp = p4 + ((d >> 8) & 0xffff);   // first 16-byte chunk: (A[k], B[k], A[k+1], B[k+1])
        float acc_a = 0.0f, acc_b = 0.0f;
        auto quad = [&](const float4 wv, const float4 p01, const float4 p23) {
          acc_a = fmaf(wv.x, p01.x, acc_a);
          acc_b = fmaf(wv.x, p01.y, acc_b);
          acc_a = fmaf(wv.y, p01.z, acc_a);
          acc_b = fmaf(wv.y, p01.w, acc_b);
          acc_a = fmaf(wv.z, p23.x, acc_a);
          acc_b = fmaf(wv.z, p23.y, acc_b);
          acc_a = fmaf(wv.w, p23.z, acc_a);
          acc_b = fmaf(wv.w, p23.w, acc_b);
        };
        int i = 0;
        // (measured alternative: the quad counts of the reference's bank (2 + 5 + 4) as a template parameter and this loop fully
        // unrolled -- 96 fewer SASS instructions, no pointer / counter arithmetic -- 10.11 ms against 9.68 ms per 4096 clips:
        // the scheduler hoists the 33 loads over the accumulate chains, spills more and the phase gets longer, not shorter.)
        {
#pragma unroll 1
          for (; i + 1 < nq; i += 2, wq += 128, pp += 4) {  // two quads per trip: six 16-byte loads in flight
            const float4 w0 = wq[0], a0 = pp[0], a1 = pp[1], w1 = wq[64], b0 = pp[2], b1 = pp[3];
            quad(w0, a0, a1);
            quad(w1, b0, b1);
          }
          if (i < nq) {
            quad(wq[0], pp[0], pp[1]);
            wq += 64;
          }
        }
        if (sg == mj.split_seg) {  // uniform: lanes 2i / 2i+1 hold the two halves of one band
          acc_a += __shfl_xor_sync(kFullMask, acc_a, 1);
          acc_b += __shfl_xor_sync(kFullMask, acc_b, 1);
        }
#ifdef CACFE_K1_MEL_ROLLED
        if ((d >> 25) & 1) {
          if (LAYOUT == LAYOUT_BTM) {
            o_btm[d & 0xff] = acc_a;
            if (store_b) o_btm[a.n_mels + (d & 0xff)] = acc_b;
          } else {
            float* o = o_img + (size_t)(d & 0xff) * a.n_frames * a.channels;
#pragma unroll 1
            for (int ch = 0; ch < a.channels; ++ch) {
              o[ch] = acc_a;
              if (store_b) o[a.channels + ch] = acc_b;
            }
          }
        }
      }
#else
        ra[sg] = acc_a;
        rb[sg] = acc_b;
        dsc[sg] = d;
      }
      if (LAYOUT == LAYOUT_BTM) {  // [b][t][m]: one coalesced row per frame
        float* o = a.out + ((size_t)b_cur * a.n_frames + ta) * a.n_mels;
#pragma unroll
        for (int sg = 0; sg < kMelMaxSeg; ++sg)
          if ((dsc[sg] >> 25) & 1) {
            o[dsc[sg] & 0xff] = ra[sg];
            if (store_b) o[a.n_mels + (dsc[sg] & 0xff)] = rb[sg];
          }
      } else {                     // [b][m][t][c]
        float* obase = a.out + ((size_t)b_cur * a.n_mels * a.n_frames + ta) * a.channels;
        const size_t m_stride = (size_t)a.n_frames * a.channels;
#pragma unroll
        for (int sg = 0; sg < kMelMaxSeg; ++sg)
          if ((dsc[sg] >> 25) & 1) {
            float* o = obase + (dsc[sg] & 0xff) * m_stride;
#pragma unroll 1
            for (int ch = 0; ch < a.channels; ++ch) {
              o[ch] = ra[sg];
              if (store_b) o[a.channels + ch] = rb[sg];
            }
          }
      }
#endif
    }
    // (the next trip's exchange stores come after release_tile's group barrier: the powers have been consumed by then)
  }
}

// [B][T][K] staging -> [B][K][T] (the layout audiowriter stores and tfdataset.read_tfrecord reshapes to, tfdataset.py:1083).
// 64 x 64 tiles through shared memory, both sides coalesced (256-byte runs), 16 loads in flight per thread.
// grid = (ceil(K/64), ceil(T/64), B), block = (32, 8).
constexpr int kTrTile = 64;
__global__ void __launch_bounds__(256) spec_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int K) {
  __shared__ float tile[kTrTile][kTrTile + 1];
  const size_t base = (size_t)blockIdx.z * T * K;
  const int k0 = blockIdx.x * kTrTile, t0 = blockIdx.y * kTrTile;
  const int nt = min(kTrTile, T - t0), nk = min(kTrTile, K - k0);
  const float* src = in + base + (size_t)t0 * K + k0;
#pragma unroll
  for (int r = threadIdx.y; r < kTrTile; r += 8) {
    if (r < nt) {
      if ((int)threadIdx.x < nk) tile[r][threadIdx.x] = ld_stream(src + (size_t)r * K + threadIdx.x);
      if ((int)threadIdx.x + 32 < nk) tile[r][threadIdx.x + 32] = ld_stream(src + (size_t)r * K + threadIdx.x + 32);
    }
  }
  __syncthreads();
  float* dst = out + base + (size_t)k0 * T + t0;
#pragma unroll
  for (int r = threadIdx.y; r < kTrTile; r += 8) {
    if (r < nk) {
      if ((int)threadIdx.x < nt) dst[(size_t)r * T + threadIdx.x] = tile[threadIdx.x][r];
      if ((int)threadIdx.x + 32 < nt) dst[(size_t)r * T + threadIdx.x + 32] = tile[threadIdx.x + 32][r];
    }
  }
}

// [B][T][M] (what the FFT kernel stores coalesced) -> the reference image [B][M][T][C] with the channel repeat of
// raw_to_mel (tfdataset.py:2052-2053).  Tiles of 64 frames x 32 bands through shared memory: 128-byte reads, and each output
// row piece is 64 * C contiguous floats (768 bytes for C = 3); 8 loads in flight per thread.
// grid = (ceil(M/32), ceil(T/64), B), block = (32, 8).
constexpr int kImgT = 64;
template <int C>   // 0: runtime channel count
__global__ void __launch_bounds__(256) btm_to_bmtc_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int M,
                                                          int channels) {
  __shared__ float tile[kImgT][33];
  const size_t b = blockIdx.z;
  const int m0 = blockIdx.x * 32, t0 = blockIdx.y * kImgT;
  const int nt = min(kImgT, T - t0);
  const int m_in = m0 + threadIdx.x;
#pragma unroll
  for (int r = threadIdx.y; r < kImgT; r += 8)
    if (r < nt && m_in < M) tile[r][threadIdx.x] = ld_stream(in + (b * T + t0 + r) * M + m_in);
  __syncthreads();
  const int cn = C > 0 ? C : channels;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int m = m0 + r;
    if (m >= M) continue;
    float* o = out + ((b * M + m) * T + t0) * cn;
    for (int i = threadIdx.x; i < nt * cn; i += 32) o[i] = tile[C > 0 ? i / C : i / cn][r];
  }
}

}  // namespace cacfe
