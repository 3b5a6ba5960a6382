// K1, round-2 form (fourth generation): the persistent fused normalise / frame / Hann / rFFT / power / mel kernel of
// k_frontend_v3.cuh without its in-place normalisation pass.
//
// What the round-1 capture of v3 showed (profiles/r01_k1_v5_source_regions.txt): 15 % of all warp samples sat in the
// tile hand-over -- the pass that normalises a freshly landed tile in place, the mbarrier every warp arrives on after
// it (a CTA-wide rendezvous once per tile: no FFT group could run more than one tile ahead of the slowest), and the
// polls around both -- for 11 % of the executed instructions.  None of that is arithmetic the result needs:
//   * the clip normalisation ((x - mn) * sc + of with sc = 2 / range, of = 2e-6 - 1; tfdataset.py:1916-1934) is affine, and the
//     DFT of a windowed frame that lies entirely inside the clip is linear in it:
//         DFT(w ((x - mn) sc + of)) = sc DFT(w (x - mn)) + of DFT(w),
//     where DFT(w) of the periodic Hann window is non-zero at bins 0 and +-1 only.  A filterbank that starts at bin >= 2
//     (9 for the reference's bank, tfdataset.py:47-56) never sees the second term, so full frames load  w (x - mn)  -- one
//     packed subtract more than before, the subtraction first so that clips whose DC dwarfs their range keep their bits --
//     and the epilogue multiplies the two or three mel sums a thread owns by sc^2 (power 2) or sc (power 1);
//   * frames that reach into the zero padding (Q4 of SURVEY appendix A: padding happens after normalisation; frames
//     498..512 of 513, and the first eight of the centred framing) take a second, rarely executed load path that applies
//     the affine map and the zeros per sample.
// What is left of the hand-over is the 2-deep TMA ring itself: a group waits for its tile to land, loads its two frames
// into registers, and the last group to do so re-arms the buffer.  Groups are free to drift up to two tiles apart.
//
// Serves n_fft = 4096, zero / end padding (not reflect), mel layouts, banks with bin_lo >= 2 when normalising; everything
// else stays on stft_mel_v3_kernel.  Arithmetic per frame is otherwise v3's (same DFT, same twiddles, same mel jobs).
#pragma once
#include "k_frontend_v3.cuh"

namespace cacfe {

struct V4Smem {
  int tile_len, tile_pad, mel_quads;
  size_t off_tile, off_exch, off_melw, off_desc, off_sync, total;
};

__host__ __device__ inline V4Smem v4_smem_layout(int hop, int mel_quads) {
  V4Smem s;
  s.mel_quads = mel_quads;
  s.tile_len = kFft + hop * (kVTileFrames - 1);
  s.tile_pad = (s.tile_len + 3 + 4) & ~3;          // room for the copy length rounded up to 16 B
  size_t o = sizeof(float4) * 2048;                 // stage twiddles
  s.off_tile = o;   o += sizeof(float) * s.tile_pad * 2;
  s.off_exch = o;   o += sizeof(float) * kHalfFloats * kVGroups;
  s.off_melw = o;   o += sizeof(float4) * 64 * (mel_quads > 0 ? mel_quads : 1);
  s.off_desc = o;   o += sizeof(int) * 64 * kMelMaxSeg;
  s.off_sync = o;   o += 64;                        // full mbarriers, done counters, normalisation pairs
  s.total = o;
  return s;
}

template <int NQ, int LAYOUT>
__global__ void __launch_bounds__(kVThreads, 1) stft_mel_v4_kernel(const FrontendArgs a, const MelArgs mj,
                                                                     const int total_tiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const V4Smem L = v4_smem_layout(a.hop, mj.total_quads);
  float4* s_tw4 = reinterpret_cast<float4*>(smem);   // [32 output pairs][64 n2]
  float* s_tile = reinterpret_cast<float*>(smem + L.off_tile);
  float* s_exch = reinterpret_cast<float*>(smem + L.off_exch);
  float4* s_melw = reinterpret_cast<float4*>(smem + L.off_melw);
  int* s_desc = reinterpret_cast<int*>(smem + L.off_desc);
  uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + L.off_sync);  // [2] tile landed
  int* s_done = reinterpret_cast<int*>(s_full + 2);                   // [2] groups that have read the tile
  float2* s_nrm = reinterpret_cast<float2*>(s_full + 4);              // [2][2] (range, min) of clips b & ~1, b | 1

  const int tid = threadIdx.x;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int g = tid >> 6, t64 = tid & 63, lane = tid & 31;

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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the buffer vs the async write
    mbar_expect_tx(bar, bytes + (a.norm != nullptr ? 16u : 0u));
    bulk_g2s(smem_u32(s_tile + (size_t)s * L.tile_pad + (c0 - s_lo)), a.in + (size_t)b * a.n_samples + c0, bytes, bar);
    if (a.norm != nullptr) bulk_g2s(smem_u32(s_nrm + 2 * s), a.norm + (b & ~1), 16u, bar);
  };

  // A group has read tile i into registers; the last of the groups re-arms the buffer with tile i + 2.
  auto release_tile = [&](int i) {
    group_barrier(1 + g, 64);
    if (t64 == 0) {
      const int s = i & 1;
      k1_jitter(3u * (unsigned)i + 1u);
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
      s_done[q] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- tables, once per CTA (L2 resident) ------------------------------------------------------------------
  for (int i = tid; i < 2048; i += kVThreads) s_tw4[i] = mj.tw4[i];
  for (int i = tid; i < 64 * mj.total_quads; i += kVThreads) s_melw[i] = mj.w[i];
  for (int i = tid; i < 64 * kMelMaxSeg; i += kVThreads) s_desc[i] = mj.desc[i];
  __syncthreads();
  if (tid == 0) {
    if (my_tiles > 0) issue_tile(0);
    if (my_tiles > 1) issue_tile(1);
  }
  // tile coordinates (clip, first frame) advance by gridDim.x tiles per trip: kept incrementally, no division in the loop
  const int step_b = (int)gridDim.x / a.tiles_per_clip, step_t = ((int)gridDim.x % a.tiles_per_clip) * kVTileFrames;
  const int wrap_t = a.tiles_per_clip * kVTileFrames;
  int b = (int)blockIdx.x / a.tiles_per_clip, t0 = ((int)blockIdx.x % a.tiles_per_clip) * kVTileFrames;

  float* ex = s_exch + g * kHalfFloats;
  float2* pbuf = reinterpret_cast<float2*>(ex);  // (power A, power B) per bin; aliases the exchange tile
  const int j = stage2_row(t64);
  const bool self = (j == 0) || (j == 32);
  const int plane = self ? lane : (lane ^ 16);
  // The groups run free: no CTA-wide barrier in the loop, only the ring hand-over.
  for (int i = 0; i < my_tiles; ++i) {
    const int s = i & 1;
    const float* tile = s_tile + (size_t)s * L.tile_pad;
    const int ta = t0 + 2 * g;
    const int b_cur = b, t0_cur = t0;
    {
      int b_next = b + step_b, t_next = t0 + step_t;
      if (t_next >= wrap_t) {
        t_next -= wrap_t;
        ++b_next;
      }
      b = b_next;
      t0 = t_next;
    }
    k1_jitter(3u * (unsigned)i);
    mbar_wait_relaxed(smem_u32(&s_full[s]), (uint32_t)((i >> 1) & 1));
    if (ta >= a.n_frames) {  // group-uniform
      release_tile(i);
      continue;
    }
    const bool store_b = ta + 1 < a.n_frames;

    // normalisation of this clip: x' = (x - mn) * sc + of  (the reference's order with one rounding less, as in v3)
    float mn = 0.0f, sc = 1.0f;
    if (a.norm != nullptr) {
      const float2 nrm = s_nrm[2 * s + (b_cur & 1)];
      mn = nrm.y;
      sc = 2.0f / nrm.x;      // range 0 -> inf: constant clips give NaN features (Q1)
    }
    // frames ta and ta + 1 read samples [p_lo, p_hi) of the clip; inside it, the linear form applies.  Ranges so small that
    // the unscaled power would underflow (range^2 * 1e-7 below the f32 normals) take the exact path as well.
    const int p_lo = a.origin + a.hop * ta;
    const int p_hi = p_lo + a.hop + kFft;
    bool fast = p_lo >= 0 && p_hi <= a.n_samples && (a.norm == nullptr || !(sc > 1.0e12f));
#ifndef CACFE_V4_SUBTRACT
    // The subtraction of the clip minimum is itself invisible to bins >= 2 (mn DFT(w) lives in bins 0 and +-1): it only keeps
    // the f32 products w x from drowning a small signal in a large DC.  Clips whose minimum is within 2x their range of zero load w x
    // directly (|x| <= 3 range: the rounding of w x stays at the level of the normalised path); the others take the exact path.
    if (a.norm != nullptr && !(fabsf(mn) * sc <= 4.0f)) fast = false;   // |mn| <= 2 range; NaN -> exact path (all NaN)
#endif
    float post = fast ? sc : 1.0f;

    float re[64], im[64];
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
      if (ph == 0) {
        // ---- stage 1 input: window; z[n] = w[n] (xA[n] + i xB[n]), thread n2 = t64 holds n = 64 q + n2 ----------------
        const float* fa = tile + (ta - t0_cur) * a.hop;
        const float* fb = fa + a.hop;
        // (cos, sin) of phi_t = 2 pi t / 4096 from the twiddle row of k = 1: (cos, -sin)(2 pi n2 / 4096)
        const float4 t1 = s_tw4[t64];
        const float win_c = t1.y, win_s = -t1.w;
        if (fast) {
#ifdef CACFE_V4_SUBTRACT
          const cacfe_f2 mn2 = cacfe_pk(mn, mn);
#endif
#pragma unroll
          for (int q = 0; q < 64; q += 2) {  // packed: the pair (q, q + 1) is also the input pair of cacfe_fft64x2
            const int n = 64 * q + t64;
            const cacfe_f2 wv = cacfe_pk(fmaf(kWinA[q], win_c, fmaf(kWinB[q], win_s, 0.5f)),
                                         fmaf(kWinA[q + 1], win_c, fmaf(kWinB[q + 1], win_s, 0.5f)));
#ifdef CACFE_V4_SUBTRACT
            const cacfe_f2 xa = cacfe_mul2(cacfe_sub2(cacfe_pk(fa[n], fa[n + 64]), mn2), wv);
            const cacfe_f2 xb = cacfe_mul2(cacfe_sub2(cacfe_pk(fb[n], fb[n + 64]), mn2), wv);
#else
            const cacfe_f2 xa = cacfe_mul2(cacfe_pk(fa[n], fa[n + 64]), wv);
            const cacfe_f2 xb = cacfe_mul2(cacfe_pk(fb[n], fb[n + 64]), wv);
#endif
            re[q] = cacfe_lo(xa);
            re[q + 1] = cacfe_hi(xa);
            im[q] = cacfe_lo(xb);
            im[q + 1] = cacfe_hi(xb);
          }
        } else {
          // a frame of the pair reaches into the padding (or the clip's range is degenerate): exact affine map per sample,
          // zeros outside the clip.  The tile holds stale data there (the bulk copy covers the clip only): select, not multiply.
          const float of = a.norm != nullptr ? -0.999998f : 0.0f;
          const int pa0 = p_lo + t64, pb0 = pa0 + a.hop;
#pragma unroll
          for (int q = 0; q < 64; ++q) {
            const int n = 64 * q + t64;
            const float w = fmaf(kWinA[q], win_c, fmaf(kWinB[q], win_s, 0.5f));
            const float va = fmaf(fa[n] - mn, sc, of), vb = fmaf(fb[n] - mn, sc, of);
            const bool ina = (unsigned)(pa0 + 64 * q) < (unsigned)a.n_samples;
            const bool inb = (unsigned)(pb0 + 64 * q) < (unsigned)a.n_samples;
            re[q] = ina ? va * w : 0.0f;
            im[q] = inb ? vb * w : 0.0f;
          }
        }
        release_tile(i);
      }
      cacfe_fft64x2(re, im);
      if (ph == 0) {
        // ---- twiddle W4096^(n2 k1), then the transpose through shared memory, real parts first ------------------------
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
    if (a.power == 1) {  // magnitude (stored-spectrogram convention, tfdataset.py:1085-1088): rolled, off the hot path
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
      float ra[kMelMaxSeg], rb[kMelMaxSeg];
      int dsc[kMelMaxSeg];
#pragma unroll
      for (int sg = 0; sg < kMelMaxSeg; ++sg) {
        const int nq = mj.nq[sg];                      // uniform
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
        int i2 = 0;
#pragma unroll 1
        for (; i2 + 1 < nq; i2 += 2, wq += 128, pp += 4) {  // two quads per trip: six 16-byte loads in flight
          const float4 w0 = wq[0], a0 = pp[0], a1 = pp[1], w1 = wq[64], b0 = pp[2], b1 = pp[3];
          quad(w0, a0, a1);
          quad(w1, b0, b1);
        }
        if (i2 < nq) {
          quad(wq[0], pp[0], pp[1]);
          wq += 64;
        }
        if (sg == mj.split_seg) {  // uniform: lanes 2i / 2i+1 hold the two halves of one band
          acc_a += __shfl_xor_sync(kFullMask, acc_a, 1);
          acc_b += __shfl_xor_sync(kFullMask, acc_b, 1);
        }
        // the linear form's scale: sc^2 for power, sc for magnitude (post == 1 on the exact path).  Two multiplies rather
        // than one by sc^2: the intermediate stays in range whenever the result does.
        acc_a *= post;
        acc_b *= post;
        if (a.power != 1) {
          acc_a *= post;
          acc_b *= post;
        }
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
    }
    // (the next trip's exchange stores come after release_tile's group barrier: the powers have been consumed by then)
  }
}

}  // namespace cacfe
