// K1 (streaming form): persistent fused frame / Hann / rFFT / power / mel kernel for sm_100a.
//
// Same arithmetic as stft_mel_kernel (k_frontend.cuh), re-organised after the first ncu capture
// (profiles/r01_k1_v1_*): that kernel spent 29 % of its warp time in long-scoreboard stalls (every CTA re-loads
// 48 KB of tables and its sample tile with nothing else resident on the SM), ran 8 warps per SM at 255 registers,
// and lost 29 % of its shared-memory wavefronts to bank conflicts in the mel epilogue.  Here:
//   * one persistent CTA per SM; tables are loaded once;
//   * sample tiles arrive by TMA bulk copy (cp.async.bulk + mbarrier complete_tx) into a 2-deep ring while the
//     FFT groups work on the previous tile; a buffer is re-armed by whichever group consumes its last frame pair;
//   * 6 independent 64-thread FFT groups (12 warps, <= 168 registers); the stage-1 -> stage-2 transpose goes
//     through shared memory one half (re, then im) at a time, which halves the exchange tile (17 KB / group);
//   * per-clip normalisation is folded into the spectrum: for a full frame FFT(w (s x + o)) = s FFT(w x) + o W,
//     and the periodic Hann transform W is non-zero only in bins 0 and +-1, so power scales by s^2 (applied to the
//     mel output) and bins 0, +-1 get an exact correction; the clip's mid-range value is removed first (one FMA
//     with the window) so the transform never sees a DC term larger than the signal.  Frames that touch zero padding (Q4: padding happens
//     after normalisation) take a masked time-domain path instead;
//   * the mel projection keeps both frames of a pair in one float2 power array, one thread per band, and stores
//     straight to global memory (coalesced over mel for the [B][T][M] layout).
#pragma once
#include "cacfe_common.cuh"
#include "frontend_core.cuh"
#include "k_frontend.cuh"

namespace cacfe {

constexpr int kSGroups = 6;
constexpr int kSThreads = kSGroups * 64;
constexpr int kSTileFrames = 16;
constexpr int kSPairs = kSTileFrames / 2;
constexpr int kHalfStride = 68;                    // floats per half-exchange row: 272 B = 2*128 + 16
constexpr int kHalfFloats = 64 * kHalfStride;      // 17408 B per group

struct TileInfo {
  float sc, of;        // normalisation x*sc + of (1, 0 when off)
  float out_scale;     // sc^2 (power 2) / sc (power 1) for spectrum-side normalisation; NaN for a constant clip
  float dc_ratio;      // residual offset / sc once the clip centre has been removed (see `center`)
  int b, t0, npairs, s_lo;
  float center;        // min + range/2, subtracted from full frames before the FFT so |x - center| <= range/2
  int pad[3];
};

struct SSmem {
  int tile_len, tile_pad, bw_in_smem;
  size_t off_win, off_tile, off_exch, off_bw, off_bstart, off_bofs, off_sync, total;
};

__host__ __device__ inline SSmem stream_smem_layout(int hop, int n_mels, int nnz, int bw_in_smem) {
  SSmem s;
  s.bw_in_smem = bw_in_smem;
  s.tile_len = kFft + hop * (kSTileFrames - 1);
  s.tile_pad = (s.tile_len + 3 + 4) & ~3;          // room for the copy length rounded up to 16 B
  size_t o = sizeof(float2) * 4096;
  s.off_win = o;    o += sizeof(float) * 2052;
  s.off_tile = o;   o += sizeof(float) * s.tile_pad * 2;
  s.off_exch = o;   o += sizeof(float) * kHalfFloats * kSGroups;
  s.off_bw = o;     o += bw_in_smem ? sizeof(float) * ((nnz + 3) & ~3) : 0;
  s.off_bstart = o; o += sizeof(int) * ((n_mels + 3) & ~3);
  s.off_bofs = o;   o += sizeof(int) * ((n_mels + 1 + 3) & ~3);
  s.off_sync = o;   o += 32 + 2 * sizeof(TileInfo);  // full[2] mbarriers, done[2] counters, TileInfo[2]
  s.total = o;
  return s;
}

// ---- mbarrier / bulk-copy wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// For waits that can last a large part of a tile (a group that runs ahead of the others): sleep between polls so the
// spinning warps do not take issue slots from the warps they are waiting for.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(800);
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int NQ>
__global__ void __launch_bounds__(kSThreads, 1) stft_mel_stream_kernel(const FrontendArgs a, const int total_tiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const SSmem L = stream_smem_layout(a.hop, a.n_mels, a.nnz, a.bw_in_smem);
  float2* s_tw = reinterpret_cast<float2*>(smem);
  float* s_win = reinterpret_cast<float*>(smem + L.off_win);
  float* s_tile = reinterpret_cast<float*>(smem + L.off_tile);
  float* s_exch = reinterpret_cast<float*>(smem + L.off_exch);
  const float* s_bw = a.bw_in_smem ? reinterpret_cast<const float*>(smem + L.off_bw) : a.band_w;
  int* s_bstart = reinterpret_cast<int*>(smem + L.off_bstart);
  int* s_bofs = reinterpret_cast<int*>(smem + L.off_bofs);
  uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + L.off_sync);          // [2]
  int* s_done = reinterpret_cast<int*>(smem + L.off_sync + 16);               // [2]
  TileInfo* s_info = reinterpret_cast<TileInfo*>(smem + L.off_sync + 32);     // [2] x 32 B

  const int tid = threadIdx.x;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // Arms buffer i&1 for this CTA's i-th tile: per-tile constants, then the bulk copy of the in-range samples.
  auto issue_tile = [&](int i) {
    const int s = i & 1;
    const int w = (int)blockIdx.x + i * (int)gridDim.x;
    const int b = w / a.tiles_per_clip;
    const int t0 = (w - b * a.tiles_per_clip) * kSTileFrames;
    const int s_lo = a.origin + a.hop * t0;
    TileInfo ti;
    ti.b = b;
    ti.t0 = t0;
    ti.s_lo = s_lo;
    ti.npairs = (min(kSTileFrames, a.n_frames - t0) + 1) >> 1;
    ti.sc = 1.0f;
    ti.of = 0.0f;
    ti.out_scale = 1.0f;
    ti.dc_ratio = 0.0f;
    ti.center = 0.0f;
    if (a.norm != nullptr) {
      const float2 rm = a.norm[b];                 // (max - min, min), folded by minmax_finalize_kernel
      const float mn = rm.y;
      const double inv = 1.0 / (double)rm.x;       // ((x-mn)/range + 1e-6 - 0.5) * 2  ==  x*sc + of
      ti.sc = (float)(2.0 * inv);
      ti.of = (float)((-(double)mn * inv + 0.000001 - 0.5) * 2.0);
      // full frames transform y = x - center; then x*sc + of = y*sc + o2 with o2 = of + sc*center (about 2e-6)
      ti.center = mn + 0.5f * rm.x;
      const double o2 = (-(double)mn * inv + 0.000001 - 0.5) * 2.0 + 2.0 * inv * (double)ti.center;
      ti.dc_ratio = (float)(o2 / (2.0 * inv));
      ti.out_scale = (a.power == 2) ? (float)(4.0 * inv * inv) : ti.sc;
      if (!(rm.x > 0.0f)) ti.out_scale = __int_as_float(0x7fc00000);  // Q1: constant clip -> NaN features
    }
    s_info[s] = ti;
    const int c0 = max(s_lo, 0);
    int c1 = min(s_lo + L.tile_len, a.n_samples);
    c1 = (c1 + 3) & ~3;  // n_samples % 4 == 0 on this path, so this never leaves the clip
    const uint32_t bytes = (uint32_t)(c1 - c0) * 4u;
    const uint32_t bar = smem_u32(&s_full[s]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this buffer vs the async write
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(s_tile + (size_t)s * L.tile_pad + (c0 - s_lo)), a.in + (size_t)b * a.n_samples + c0, bytes, bar);
  };

  if (tid == 0) {
    mbar_init(smem_u32(&s_full[0]), 1);
    mbar_init(smem_u32(&s_full[1]), 1);
    s_done[0] = 0;
    s_done[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- tables, once per CTA (L2 resident) ------------------------------------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(a.tw);
    float4* dst = reinterpret_cast<float4*>(s_tw);
    for (int i = tid; i < 2048; i += kSThreads) dst[i] = src[i];
    for (int i = tid; i < 2049; i += kSThreads) s_win[i] = a.win[i];
    if (a.bw_in_smem)
      for (int i = tid; i < a.nnz; i += kSThreads) reinterpret_cast<float*>(smem + L.off_bw)[i] = a.band_w[i];
    for (int i = tid; i < a.n_mels; i += kSThreads) s_bstart[i] = a.band_start[i];
    for (int i = tid; i <= a.n_mels; i += kSThreads) s_bofs[i] = a.band_ofs[i];
  }
  __syncthreads();
  if (tid == 0) {
    issue_tile(0);
    if (my_tiles > 1) issue_tile(1);
  }

  const int g = tid >> 6, t64 = tid & 63, lane = tid & 31;
  float* ex = s_exch + g * kHalfFloats;
  float2* pbuf = reinterpret_cast<float2*>(ex);  // (power A, power B) per bin; aliases the exchange tile
  const int j = stage2_row(t64);
  const bool self = (j == 0) || (j == 32);
  const int plane = self ? lane : (lane ^ 16);
  const int total_seq = my_tiles * kSPairs;

  for (int seq = g; seq < total_seq; seq += kSGroups) {
    const int i = seq >> 3, p = seq & 7, s = i & 1;
    mbar_wait(smem_u32(&s_full[s]), (uint32_t)((i >> 1) & 1));
    const TileInfo ti = s_info[s];
    const bool active = p < ti.npairs;
    float re[64], im[64];
    float out_scale = ti.out_scale;
    bool spectral = false;  // normalisation applied on the spectrum side (full frames only)

    if (active) {
      // ---- stage 1: window, 64-point FFT over n1, twiddle ------------------------------------------------------
      const float* fa = s_tile + (size_t)s * L.tile_pad + (2 * p) * a.hop;
      const float* fb = fa + a.hop;
      const int start_a = ti.s_lo + (2 * p) * a.hop;  // clip sample index of frame A, n = 0
      const int start_b = start_a + a.hop;
      const bool b_exists = ti.t0 + 2 * p + 1 < a.n_frames;
      const bool full = start_a >= 0 && start_b + kFft <= a.n_samples && b_exists;
      if (full) {
        spectral = a.norm != nullptr;
        const float neg_c = -ti.center;
#pragma unroll
        for (int q = 0; q < 64; ++q) {
          const int n = 64 * q + t64;
          const float w = s_win[q < 32 ? n : kFft - n];
          const float cw = neg_c * w;          // (x - center) * w in one rounding
          re[q] = fmaf(fa[n], w, cw);
          im[q] = fmaf(fb[n], w, cw);
        }
      } else {
        // frame touches the zero padding (or has no partner): normalise in the time domain, mask, no output scale
        out_scale = (ti.out_scale == ti.out_scale) ? 1.0f : ti.out_scale;
        const int lo_a = -start_a, hi_a = a.n_samples - start_a;
        const int lo_b = -start_b, hi_b = b_exists ? a.n_samples - start_b : lo_b;
#pragma unroll
        for (int q = 0; q < 64; ++q) {
          const int n = 64 * q + t64;
          const float w = s_win[q < 32 ? n : kFft - n];
          const float xa = (n >= lo_a && n < hi_a) ? fmaf(fa[n], ti.sc, ti.of) : 0.0f;
          const float xb = (n >= lo_b && n < hi_b) ? fmaf(fb[n], ti.sc, ti.of) : 0.0f;
          re[q] = xa * w;
          im[q] = xb * w;
        }
      }
      cacfe_fft64(re, im);
#pragma unroll
      for (int k1 = 0; k1 < 64; ++k1) {
        const int sl = CACFE_FFT64_SLOT(k1);
        const float2 t = s_tw[k1 * 64 + t64];
        const float yr = re[sl] * t.x - im[sl] * t.y;
        const float yi = re[sl] * t.y + im[sl] * t.x;
        re[sl] = yr;
        im[sl] = yi;
      }
      // half exchange, real parts: row k1, column n2
#pragma unroll
      for (int k1 = 0; k1 < 64; ++k1) ex[k1 * kHalfStride + t64] = re[CACFE_FFT64_SLOT(k1)];
      group_barrier(1 + g, 64);  // also: every thread of the group has finished reading the sample tile
    }
    // ---- the tile buffer is released by pair, the last release re-arms it with this CTA's tile i + 2 -----------
    if (t64 == 0) {
      __threadfence_block();
      const int old = atomicAdd(&s_done[s], 1);
      if (old == kSPairs - 1) {
        __threadfence_block();
        s_done[s] = 0;
        if (i + 2 < my_tiles) issue_tile(i + 2);
      }
    }
    if (!active) continue;

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
    // NOTE: re[] now holds stage-2 inputs in natural order while im[] still holds stage-1 outputs in slot order.
#pragma unroll
    for (int k1 = 0; k1 < 64; ++k1) ex[k1 * kHalfStride + t64] = im[CACFE_FFT64_SLOT(k1)];
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

    // ---- stage 2: 64-point FFT over n2, split the two frames, power -----------------------------------------------
    cacfe_fft64(re, im);
    if (a.bin_lo <= 1 && spectral) {
      // spectrum-side normalisation: bins 0 and +-1 also carry o/s * (1 + i) * W[k], W = {2048, -1024, -1024}
      const float d = ti.dc_ratio;
      if (j == 0) { re[CACFE_FFT64_SLOT(0)] += 2048.0f * d; im[CACFE_FFT64_SLOT(0)] += 2048.0f * d; }
      if (j == 1) { re[CACFE_FFT64_SLOT(0)] -= 1024.0f * d; im[CACFE_FFT64_SLOT(0)] -= 1024.0f * d; }
      if (j == 63) { re[CACFE_FFT64_SLOT(63)] -= 1024.0f * d; im[CACFE_FFT64_SLOT(63)] -= 1024.0f * d; }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float pr = __shfl_sync(kFullMask, re[CACFE_FFT64_SLOT(63 - q)], plane);
      float pi = __shfl_sync(kFullMask, im[CACFE_FFT64_SLOT(63 - q)], plane);
      if (j == 0) {
        pr = re[CACFE_FFT64_SLOT((64 - q) & 63)];
        pi = im[CACFE_FFT64_SLOT((64 - q) & 63)];
      }
      const int k = j + 64 * q;
      if (k >= a.bin_lo && k <= a.bin_hi) {
        float pa, pb;
        split_power(re[CACFE_FFT64_SLOT(q)], im[CACFE_FFT64_SLOT(q)], pr, pi, a.power, pa, pb);
        pbuf[k - a.bin_lo] = make_float2(pa, pb);
      }
    }
    group_barrier(1 + g, 64);

    // ---- banded mel projection, both frames of the pair per thread, straight to global ------------------------------
    const int ta = ti.t0 + 2 * p;
    const bool store_b = ta + 1 < a.n_frames;
    for (int m = t64; m < a.n_mels; m += 64) {
      const int o0 = s_bofs[m], o1 = s_bofs[m + 1];
      const float2* pp = pbuf + s_bstart[m] - o0;
      float acc_a = 0.0f, acc_b = 0.0f;
      for (int t = o0; t < o1; ++t) {
        const float wgt = s_bw[t];
        const float2 v = pp[t];
        acc_a = fmaf(wgt, v.x, acc_a);
        acc_b = fmaf(wgt, v.y, acc_b);
      }
      acc_a *= out_scale;
      acc_b *= out_scale;
      if (a.layout == LAYOUT_BTM) {
        float* o = a.out + ((size_t)ti.b * a.n_frames + ta) * a.n_mels + m;
        o[0] = acc_a;
        if (store_b) o[a.n_mels] = acc_b;
      } else {
        float* o = a.out + (((size_t)ti.b * a.n_mels + m) * a.n_frames + ta) * a.channels;
        for (int c = 0; c < a.channels; ++c) o[c] = acc_a;
        if (store_b)
          for (int c = 0; c < a.channels; ++c) o[a.channels + c] = acc_b;
      }
    }
    group_barrier(1 + g, 64);  // powers consumed before the next pair's exchange overwrites them
  }
}

}  // namespace cacfe
