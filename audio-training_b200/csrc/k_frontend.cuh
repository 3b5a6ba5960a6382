// K0 (per-clip min/max) and K1 (fused normalise / frame / Hann / rFFT / power / mel) for sm_100a.
//
// Replaces, per clip: normalize (tfdataset.py:1916-1934, predict_utils.py:153-160), tf.signal.stft
// 4096/281 pad_end (tfdataset.py:2026-2034) or the centred librosa.stft (predict_utils.py:194),
// pow/transpose/abs (tfdataset.py:2044-2046), the mel batch_dot (tfdataset.py:2049-2051,
// custommel.py:57-61) and the channel repeat (tfdataset.py:2052-2053).
#pragma once
#include "cacfe_common.cuh"
#include "frontend_core.cuh"

namespace cacfe {

// ------------------------------------------------------------------------------------------------
// K0: partial (min, max) of each row.  grid = (splits, rows), block = 256.  No atomics: the consumer
// folds the `splits` partials itself.  HBM bound: reads the clip once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_minmax_kernel(const float* __restrict__ in, long long n, int splits,
                                                         float2* __restrict__ partial) {
  __shared__ float scratch[64];
  const long long row = blockIdx.y;
  const float* x = in + row * n;
  float mn = INFINITY, mx = -INFINITY;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (n % 4 == 0);
  if (vec) {
    const long long n4 = n / 4;
    const long long per = (n4 + splits - 1) / splits;
    const long long lo = per * blockIdx.x, hi = min(n4, lo + per);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    long long i = lo + threadIdx.x;
    for (; i + 3 * 256 < hi; i += 4 * 256) {  // 4 independent 16 B loads in flight per thread
      const float4 a = ld_stream4(x4 + i), b = ld_stream4(x4 + i + 256), c = ld_stream4(x4 + i + 512),
                   d = ld_stream4(x4 + i + 768);
      mn = min_nan(mn, min_nan(min_nan(min_nan(a.x, a.y), min_nan(a.z, a.w)), min_nan(min_nan(b.x, b.y), min_nan(b.z, b.w))));
      mn = min_nan(mn, min_nan(min_nan(min_nan(c.x, c.y), min_nan(c.z, c.w)), min_nan(min_nan(d.x, d.y), min_nan(d.z, d.w))));
      mx = max_nan(mx, max_nan(max_nan(max_nan(a.x, a.y), max_nan(a.z, a.w)), max_nan(max_nan(b.x, b.y), max_nan(b.z, b.w))));
      mx = max_nan(mx, max_nan(max_nan(max_nan(c.x, c.y), max_nan(c.z, c.w)), max_nan(max_nan(d.x, d.y), max_nan(d.z, d.w))));
    }
    for (; i < hi; i += 256) {
      const float4 a = ld_stream4(x4 + i);
      mn = min_nan(mn, min_nan(min_nan(a.x, a.y), min_nan(a.z, a.w)));
      mx = max_nan(mx, max_nan(max_nan(a.x, a.y), max_nan(a.z, a.w)));
    }
  } else {
    const long long per = (n + splits - 1) / splits;
    const long long lo = per * blockIdx.x, hi = min(n, lo + per);
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
      const float v = x[i];
      mn = min_nan(mn, v);
      mx = max_nan(mx, v);
    }
  }
  block_minmax_nan(mn, mx, scratch);
  if (threadIdx.x == 0) partial[row * splits + blockIdx.x] = make_float2(mn, mx);
}

__device__ __forceinline__ void fold_partials(const float2* __restrict__ partial, long long row, int splits, float& mn,
                                              float& mx) {
  mn = INFINITY;
  mx = -INFINITY;
  for (int s = 0; s < splits; ++s) {
    const float2 p = partial[row * splits + s];
    mn = min_nan(mn, p.x);
    mx = max_nan(mx, p.y);
  }
}

// Standalone normalize (a1), the reference's operation order in f32 with a true division:
//   x -= min; x = x / max(x) + 1e-6; x -= 0.5; x *= 2.           grid = (blocks, rows)
__global__ void __launch_bounds__(256) row_normalize_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            long long n, int splits,
                                                            const float2* __restrict__ partial) {
  const long long row = blockIdx.y;
  float mn, mx;
  fold_partials(partial, row, splits, mn, mx);
  const float range = mx - mn;  // == max(x - mn): rounding is monotone
  const float* x = in + row * n;
  float* y = out + row * n;
  auto point = [&](float s) -> float {   // the reference's f32 order, one rounding per operation
    float v = s - mn;
    v = __fadd_rn(__fdiv_rn(v, range), 0.000001f);
    v = __fsub_rn(v, 0.5f);
    return __fmul_rn(v, 2.0f);
  };
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
#pragma unroll 4
    for (long long i = first; i < (n >> 2); i += stride) {
      const float4 v = ld_stream4(x4 + i);
      y4[i] = make_float4(point(v.x), point(v.y), point(v.z), point(v.w));
    }
  } else {
    for (long long i = first; i < n; i += stride) y[i] = point(x[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
enum : int { FRAME_TF_PAD_END = 0, FRAME_CENTER_ZERO = 1, FRAME_CENTER_REFLECT = 2, FRAME_NO_PAD = 3 };
enum : int { LAYOUT_BMTC = 0, LAYOUT_BTM = 1, LAYOUT_SPEC = 2, LAYOUT_SPECT = 3 };  // persistent kernel only: SPEC = spectrogram staged as
                                                                                    // [B][T][K], SPECT = written as the stored [B][K][T]

struct FrontendArgs {
  const float* in;         // [B][n_samples]
  float* out;              // [B][M][T][C] or [B][T][M]
  const float2* norm;      // [B] (max - min, min) per clip, or nullptr: no normalisation
  const float2* tw;        // [64][64]  W4096^(k1*n2)
  const float* win;        // [2049] periodic Hann, first half + centre
  const float* band_w;     // packed non-zero filterbank weights, band after band
  const int* band_start;   // [M]   first bin of band m, relative to bin_lo
  const int* band_ofs;     // [M+1] prefix offsets into band_w
  int bw_in_smem;
  int n_samples, hop, n_frames, n_mels, nnz;
  int origin;              // sample index of frame 0, element 0 (0, or -n_fft/2 for centred framing)
  int reflect;             // centred framing with reflect padding
  int power;               // 1 | 2
  int channels, layout;
  int bin_lo, bin_hi;      // inclusive band of bins with non-zero weights
  int tiles_per_clip;
};

constexpr int kGroups = 4;          // 64-thread FFT groups per CTA
constexpr int kTileFrames = 16;     // frames per CTA tile (8 frame pairs, 2 per group)
constexpr int kK1Threads = kGroups * 64;

struct K1Smem {
  int tile_len, tile_pad, out_stride, bw_in_smem;
  size_t off_win, off_tile, off_exch, off_out, off_bw, off_bstart, off_bofs, total;
};

// bw_in_smem = 0: the packed band weights stay in global memory (L1/L2) -- for banks too wide for the CTA's budget.
__host__ __device__ inline K1Smem k1_smem_layout(int hop, int n_mels, int nnz, int bw_in_smem) {
  K1Smem s;
  s.bw_in_smem = bw_in_smem;
  s.tile_len = kFft + hop * (kTileFrames - 1);
  s.tile_pad = (s.tile_len + 3) & ~3;
  s.out_stride = n_mels | 1;  // odd stride: transposed read-out is bank-conflict free
  size_t o = sizeof(float2) * 4096;
  s.off_win = o;   o += sizeof(float) * 2052;
  s.off_tile = o;  o += sizeof(float) * s.tile_pad;
  s.off_exch = o;  o += sizeof(float2) * kExchFloat2 * kGroups;
  s.off_out = o;   o += sizeof(float) * ((kTileFrames * s.out_stride + 3) & ~3);
  s.off_bw = o;    o += bw_in_smem ? sizeof(float) * ((nnz + 3) & ~3) : 0;
  s.off_bstart = o; o += sizeof(int) * ((n_mels + 3) & ~3);
  s.off_bofs = o;  o += sizeof(int) * ((n_mels + 1 + 3) & ~3);
  s.total = o;
  return s;
}

// NQ = number of 64-bin column groups the filterbank reaches: bins j + 64 q, q < NQ.
template <int NQ>
__global__ void __launch_bounds__(kK1Threads, 1) stft_mel_kernel(const FrontendArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const K1Smem L = k1_smem_layout(a.hop, a.n_mels, a.nnz, a.bw_in_smem);
  float2* s_tw = reinterpret_cast<float2*>(smem);
  float* s_win = reinterpret_cast<float*>(smem + L.off_win);
  float* s_tile = reinterpret_cast<float*>(smem + L.off_tile);
  float2* s_exch = reinterpret_cast<float2*>(smem + L.off_exch);
  float* s_out = reinterpret_cast<float*>(smem + L.off_out);
  const float* s_bw = a.bw_in_smem ? reinterpret_cast<const float*>(smem + L.off_bw) : a.band_w;
  int* s_bstart = reinterpret_cast<int*>(smem + L.off_bstart);
  int* s_bofs = reinterpret_cast<int*>(smem + L.off_bofs);

  const int tid = threadIdx.x;
  const int b = blockIdx.x / a.tiles_per_clip;
  const int t0 = (blockIdx.x - b * a.tiles_per_clip) * kTileFrames;
  const int frames_here = min(kTileFrames, a.n_frames - t0);

  // ---- tables (L2 resident) ---------------------------------------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(a.tw);
    float4* dst = reinterpret_cast<float4*>(s_tw);
    for (int i = tid; i < 2048; i += kK1Threads) dst[i] = src[i];
    for (int i = tid; i < 2049; i += kK1Threads) s_win[i] = a.win[i];
    if (a.bw_in_smem)
      for (int i = tid; i < a.nnz; i += kK1Threads) reinterpret_cast<float*>(smem + L.off_bw)[i] = a.band_w[i];
    for (int i = tid; i < a.n_mels; i += kK1Threads) s_bstart[i] = a.band_start[i];
    for (int i = tid; i <= a.n_mels; i += kK1Threads) s_bofs[i] = a.band_ofs[i];
  }
  // ---- sample tile: normalise on load (one FMA), zero pad after normalising (Q4) -----------------
  {
    float sc = 1.0f, of = 0.0f;
    if (a.norm != nullptr) {
      const float2 rm = a.norm[b];
      const float mn = rm.y;
      const double inv = 1.0 / (double)rm.x;   // ((x-mn)/range + 1e-6 - 0.5) * 2  ==  x*sc + of
      sc = (float)(2.0 * inv);
      of = (float)((-(double)mn * inv + 0.000001 - 0.5) * 2.0);
    }
    const float* x = a.in + (size_t)b * a.n_samples;
    const int s_lo = a.origin + a.hop * t0;
    for (int i = tid; i < L.tile_len; i += kK1Threads) {
      int s = s_lo + i;
      if (a.reflect) {  // numpy 'reflect': no edge repeat
        if (s < 0) s = -s;
        if (s >= a.n_samples) s = 2 * (a.n_samples - 1) - s;
      }
      float v = 0.0f;
      if (s >= 0 && s < a.n_samples) v = fmaf(ld_stream(x + s), sc, of);
      s_tile[i] = v;
    }
  }
  __syncthreads();

  // ---- FFT groups: each 64-thread group transforms frame pairs (2p, 2p+1) --------------------------
  const int g = tid >> 6, t64 = tid & 63, lane = tid & 31;
  float2* exch = s_exch + g * kExchFloat2;
  float* pbuf = reinterpret_cast<float*>(exch);  // aliases the exchange tile once stage 2 has read it
  const int nk = a.bin_hi - a.bin_lo + 1;
  const int j = stage2_row(t64);
  const bool self = (j == 0) || (j == 32);
  const int plane = self ? lane : (lane ^ 16);
  const int npairs = (frames_here + 1) >> 1;

  for (int p = g; p < npairs; p += kGroups) {
    stage1(s_tile + (2 * p) * a.hop, s_tile + (2 * p + 1) * a.hop, s_win, s_tw, t64, exch);
    group_barrier(1 + g, 64);
    float re[64], im[64];
    stage2_load(exch, j, re, im);
    group_barrier(1 + g, 64);  // every row is in registers: the tile may be overwritten with powers
    cacfe_fft64(re, im);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      // Z[N-k] for k = j + 64 q: thread 64-j holds it in slot 63-q; thread 0 in its own slot (64-q)&63.
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
        pbuf[k - a.bin_lo] = pa;
        pbuf[nk + k - a.bin_lo] = pb;
      }
    }
    group_barrier(1 + g, 64);
    // banded mel projection: 2 frames x n_mels dot products of 2..30 taps
    for (int e = t64; e < 2 * a.n_mels; e += 64) {
      const int f = e >= a.n_mels ? 1 : 0;
      const int m = e - f * a.n_mels;
      const int o0 = s_bofs[m], o1 = s_bofs[m + 1];
      const float* pp = pbuf + f * nk + s_bstart[m] - o0;
      float acc = 0.0f;
      for (int i = o0; i < o1; ++i) acc = fmaf(s_bw[i], pp[i], acc);
      s_out[(2 * p + f) * L.out_stride + m] = acc;
    }
    group_barrier(1 + g, 64);  // powers consumed before the next pair's stage 1 overwrites them
  }
  __syncthreads();

  // ---- write the tile out ------------------------------------------------------------------------------
  if (a.layout == LAYOUT_BTM) {
    float* o = a.out + ((size_t)b * a.n_frames + t0) * a.n_mels;
    const int total = frames_here * a.n_mels;
    for (int i = tid; i < total; i += kK1Threads) {
      const int f = i / a.n_mels, m = i - f * a.n_mels;
      o[i] = s_out[f * L.out_stride + m];
    }
  } else {
    const int run = frames_here * a.channels;  // contiguous floats per mel row
    const int total = a.n_mels * run;
    for (int i = tid; i < total; i += kK1Threads) {
      const int m = i / run, r = i - m * run;
      const int f = r / a.channels;
      a.out[(((size_t)b * a.n_mels + m) * a.n_frames + t0) * a.channels + r] = s_out[f * L.out_stride + m];
    }
  }
}

}  // namespace cacfe
