// K3/K4: PCEN (EMA smoother + gain/bias/root compression) and the min-max epilogue, sm_100a.
//
// Replaces tfpcen.py:8-39 (ExponentialMovingAverage: tf.scan over time), :89-95 (PCEN.call) and
// :105-110 (tensor-global normalize_minmax).
//
// Data is viewed as x[outer][T][inner] with the recurrence along T:
//   reference contract [B, T, F]        -> outer = B,       inner = F   (lanes across F: coalesced rows)
//   image layout       [B, M, T, C] (*) -> outer = B * M,   inner = C   (* our rank-4 extension)
// Each thread owns one (outer, inner) lane and walks T sequentially -- the reference's exact EMA order --
// with kPcenUnroll (32) independent loads in flight.  HBM bound: 1 read (+1 write) of 4 B per element and pass.
//
// The tensor-global min/max (Q14) is a grid-wide dependency, so the op is two passes over the *input*:
//   pass REDUCE : PCEN in registers, block (min, max) partials only            (reads 4 B / element)
//   pass APPLY  : PCEN again, rescale with the folded (min, max), store         (reads 4 B, writes 4 B)
// Both passes run the same instruction sequence, so the extremes map to exactly -1 and +1.
#pragma once
#include "cacfe_common.cuh"

namespace cacfe {

struct PcenArgs {
  const float* in;
  float* out;
  int T, inner, rows_per_clip;   // rows_per_clip = (outer / B) * inner
  float w, one_minus_w;          // clip(smooth, 0, 1)
  float gain;                    // min(gain, 1)
  float bias, inv_root;          // 1 / max(root, 1)
  float bias_pow;                // bias ** inv_root
  float eps;
  int root_is_2;
  float2* partial;               // [B][gridDim.x]           (REDUCE)
  const float2* extremes;        // [1] or [B] (min, max)    (APPLY)
  int per_clip_extremes;
  const float* init;             // ema_kernel: initial state [outer][inner], or nullptr = inputs[:, 0, :]
  int zero;                      // always 0, and only the host knows: pcen_kernel's load fence
};

enum : int { PCEN_REDUCE = 0, PCEN_APPLY = 1, PCEN_RAW = 2 };

// MUFU wrappers in their flush-to-zero form: one instruction each.  The default forms (`__log2f`, `exp2f`, `rsqrtf` without
// -ftz) wrap the same MUFU in a compare and two predicated multiplies that rescale subnormal arguments / results -- nine
// instructions per element of a pass that issue, not HBM, holds back (APPLY: 36 instructions per element, issue slots 68 %).
// For normal arguments and results the values are the same bit for bit.  Subnormals: eps + M subnormal needs M = -eps to
// 1e-38 (impossible for eps = 1e-6: the sum of two floats of opposite sign is a multiple of 9e-14); a smoother gain below
// 1e-38 needs M > 1e38; a subnormal y is mapped to 0 by the select in pcen_root (error <= 1.1e-19).
__device__ __forceinline__ float lg2_ftz(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// y = x / (eps + M)^gain + bias,  with  x / (eps + M)^gain == x * 2^(-gain * log2(eps + M))      (MUFU lg2 / ex2)
__device__ __forceinline__ float pcen_y(float x, float m, const PcenArgs& a) {
  const float smooth = ex2_ftz(-a.gain * lg2_ftz(a.eps + m));
  return fmaf(x, smooth, a.bias);
}
// y^(1/root) - bias^(1/root).  root 2 (the layer's initial value): y * rsqrt(y) -- one MUFU and one multiply, <= 2 ulp --
// instead of the IEEE sqrt sequence.  ROOT2: the caller knows at compile time that root is 2 (no branch per element).
template <bool ROOT2 = false>
__device__ __forceinline__ float pcen_root(float y, const PcenArgs& a) {
  const float r = (ROOT2 || a.root_is_2) ? (y >= 1.17549435e-38f ? y * rsqrt_ftz(y) : 0.0f) : ex2_ftz(a.inv_root * lg2_ftz(y));
  return r - a.bias_pow;
}
template <bool ROOT2 = false>
__device__ __forceinline__ float pcen_point(float x, float m, const PcenArgs& a) { return pcen_root<ROOT2>(pcen_y(x, m, a), a); }

// Loads in flight per lane.  Measured on [4096, 513, 160], both passes of the tensor-global scope (tools/probe_pcen_waves.py,
// results bit-identical): 8 -> 0.811 ms, 12 -> 0.821, 16 -> 0.778, 24 -> 0.776, 32 -> 0.752, 48 -> 0.868, 64 -> 0.895.  The passes are latency-bound
// (capping the resident blocks per SM to make the rounds of a launch come out even only slows them down: 0.81 -> 1.21 ms from
// 12 down to 4 blocks), so bytes in flight per lane are what pays.  ema_kernel keeps its own double-buffered batch of 8.
#ifndef CACFE_PCEN_UNROLL
#define CACFE_PCEN_UNROLL 32
#endif
constexpr int kPcenUnroll = CACFE_PCEN_UNROLL;
constexpr int kEmaUnroll = 8;

// ROOT2: the layer's initial root (2) resolved at compile time -- no branch per element.
#ifndef CACFE_PCEN_FENCE       // A/B switch (tools/probe_pcen_hot.py)
#define CACFE_PCEN_FENCE 1
#endif
template <int MODE, bool ROOT2 = false>
__global__ void __launch_bounds__(256) pcen_kernel(const PcenArgs a) {
  __shared__ float scratch[64];
  const int inner = a.inner;
  // REDUCE walks the clips from the last to the first and APPLY from the first to the last: what the producer wrote last (and
  // what REDUCE read last) is still in the 126 MB L2 when the next pass starts there
  const int clip = MODE == PCEN_REDUCE ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;  // row inside the clip
  const bool live = r < a.rows_per_clip;
  float mn = INFINITY, mx = -INFINITY;
  if (live) {
    const int o = r / inner, i = r - o * inner;
    const size_t base = ((size_t)clip * (a.rows_per_clip / inner) + o) * a.T * inner + i;
    const float* x = a.in + base;
    float* y = a.out + base;
    float scale = 1.0f, shift = 0.0f;
    if (MODE == PCEN_APPLY) {
      const float2 e = a.extremes[a.per_clip_extremes ? clip : 0];  // (range, min)
      // 2 * ((v - min) / range) - 1 as one FMA with 2 / range: one division per thread instead of one per element
      // (the pass is MUFU-bound; the result moves by <= 1 ulp of a value in [-1, 1])
      scale = 2.0f / e.x;
      if (fmaf(e.x, scale, -2.0f) < 0.0f) scale = __uint_as_float(__float_as_uint(scale) + 1u);  // exact: range * scale >= 2, the maximum
      shift = e.y;                                                                  // reaches 1 and is clamped to it
    }
    float m = x[0];  // initial state = inputs[:, 0, :]  (tfpcen.py:92)
    auto point = [&](float v, float* dst) {
      m = __fadd_rn(__fmul_rn(a.w, v), __fmul_rn(a.one_minus_w, m));  // unfused, the reference's f32 order
      if (MODE == PCEN_REDUCE) {
        // the root is monotone: the extremes of p are the roots of the extremes of y (pcen_extremes_kernel takes them with
        // the same instruction sequence APPLY uses) -- two MUFU per element instead of three in the pass they bound
        const float yy = pcen_y(v, m, a);
        mn = min_nan(mn, yy);   // NaN-propagating like tf.reduce_min / reduce_max (tfpcen.py:108-109)
        mx = max_nan(mx, yy);
      } else {
        float p = pcen_point<ROOT2>(v, m, a);
        // clamp to [-1, 1] with NaN passing through: a constant tensor (range 0 -> 0 * inf) gives NaN as the reference's 0 / 0 does
        if (MODE == PCEN_APPLY) p = max_nan(min_nan(fmaf(p - shift, scale, -1.0f), 1.0f), -1.0f);
        *dst = p;
      }
    };
    // (prefetching the next batch into a second register set -- what ema_kernel does -- costs this MUFU-bound kernel
    // occupancy: measured 0.48 -> 0.54 ms on [2048, 513, 160])
    int t = 0;
    for (; t + kPcenUnroll <= a.T; t += kPcenUnroll) {
      float v[kPcenUnroll];
#pragma unroll
      for (int u = 0; u < kPcenUnroll; ++u) v[u] = ld_stream(x + (size_t)(t + u) * inner);
      // load fence: the first value is made to depend on the last load (OR with `last & 0`, the zero read from the constant
      // bank).  Without it ptxas sinks the loads below the arithmetic to save registers (32 instead of 48 - 60) and about nine
      // of the 32 stay in flight: measured on [4096, 513, 160], REDUCE + APPLY, 0.731 ms without and 0.713 ms with the fence.
      // (Compile-time row stride, immediate offsets: four instructions fewer per element, and ptxas hoists the one load the
      // fence names; with a fence over all 32 loads 0.710 ms -- once the loads are in flight the instruction count is not what
      // bounds the pass.)
      if (CACFE_PCEN_FENCE) v[0] = __int_as_float(__float_as_int(v[0]) | (__float_as_int(v[kPcenUnroll - 1]) & a.zero));
#pragma unroll
      for (int u = 0; u < kPcenUnroll; ++u) point(v[u], y + (size_t)(t + u) * inner);
    }
    for (; t < a.T; ++t) point(ld_stream(x + (size_t)t * inner), y + (size_t)t * inner);
  }
  if (MODE == PCEN_REDUCE) {
    block_minmax_nan(mn, mx, scratch);
    if (threadIdx.x == 0) a.partial[(size_t)clip * gridDim.x + blockIdx.x] = make_float2(mn, mx);
  }
}

// Fold block partials into (range, min) per scope entry.  grid = entries, block = 256.
// NANP: NaN-propagating fold (the clip normalisation: numpy semantics, a NaN sample makes the clip's range NaN).
template <bool NANP = false>
__global__ void __launch_bounds__(256) minmax_finalize_kernel(const float2* __restrict__ partial, int per_entry,
                                                              float2* __restrict__ extremes) {
  __shared__ float scratch[64];
  const float2* p = partial + (size_t)blockIdx.x * per_entry;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < per_entry; i += blockDim.x) {
    mn = NANP ? min_nan(mn, p[i].x) : fminf(mn, p[i].x);
    mx = NANP ? max_nan(mx, p[i].y) : fmaxf(mx, p[i].y);
  }
  if (NANP) block_minmax_nan(mn, mx, scratch);
  else block_minmax(mn, mx, scratch);
  if (threadIdx.x == 0) extremes[blockIdx.x] = make_float2(mx - mn, mn);  // (range, min)
}

// Block partials of y (pcen_kernel<REDUCE>) -> (range, min) of p per scope entry, or (min, max) of p when `raw`.
__global__ void __launch_bounds__(256) pcen_extremes_kernel(const float2* __restrict__ partial, int per_entry, float2* __restrict__ extremes,
                                                            const PcenArgs a, int raw) {
  __shared__ float scratch[64];
  const float2* p = partial + (size_t)blockIdx.x * per_entry;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < per_entry; i += blockDim.x) {
    mn = min_nan(mn, p[i].x);
    mx = max_nan(mx, p[i].y);
  }
  block_minmax_nan(mn, mx, scratch);
  if (threadIdx.x == 0) {
    const float pmn = pcen_root(mn, a), pmx = pcen_root(mx, a);
    extremes[blockIdx.x] = raw ? make_float2(pmn, pmx) : make_float2(pmx - pmn, pmn);
  }
}

// EMA alone (tfpcen.ExponentialMovingAverage.call): same walk, no compression.
__global__ void __launch_bounds__(256) ema_kernel(const PcenArgs a) {
  const int clip = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.rows_per_clip) return;
  const int o = r / a.inner, i = r - o * a.inner;
  const size_t base = ((size_t)clip * (a.rows_per_clip / a.inner) + o) * a.T * a.inner + i;
  const float* x = a.in + base;
  float* y = a.out + base;
  // tf.scan's initializer (tfpcen.py:36-38): any state; the reference's only caller passes inputs[:, 0, :] (tfpcen.py:92)
  float m = a.init != nullptr ? a.init[((size_t)clip * (a.rows_per_clip / a.inner) + o) * a.inner + i] : x[0];
  int t = 0;
#ifdef CACFE_EMA_DOUBLE_BUFFER   // A/B switch: the round-1 form, a double-buffered batch of 8 (16 loads in flight)
  float v[kEmaUnroll], nv[kEmaUnroll];
  if (a.T >= kEmaUnroll) {
#pragma unroll
    for (int u = 0; u < kEmaUnroll; ++u) v[u] = ld_stream(x + (size_t)u * a.inner);
  }
  for (; t + kEmaUnroll <= a.T; t += kEmaUnroll) {   // next batch of loads in flight while this one is folded and stored
    if (t + 2 * kEmaUnroll <= a.T) {
#pragma unroll
      for (int u = 0; u < kEmaUnroll; ++u) nv[u] = ld_stream(x + (size_t)(t + kEmaUnroll + u) * a.inner);
    }
#pragma unroll
    for (int u = 0; u < kEmaUnroll; ++u) {
      m = __fadd_rn(__fmul_rn(a.w, v[u]), __fmul_rn(a.one_minus_w, m));  // w*x + (1-w)*a, unfused like TF
      y[(size_t)(t + u) * a.inner] = m;
    }
#pragma unroll
    for (int u = 0; u < kEmaUnroll; ++u) v[u] = nv[u];
  }
#else
  // batches of 32 loads held ahead of the arithmetic by the same fence as in pcen_kernel
  for (; t + kPcenUnroll <= a.T; t += kPcenUnroll) {
    float v[kPcenUnroll];
#pragma unroll
    for (int u = 0; u < kPcenUnroll; ++u) v[u] = ld_stream(x + (size_t)(t + u) * a.inner);
    v[0] = __int_as_float(__float_as_int(v[0]) | (__float_as_int(v[kPcenUnroll - 1]) & a.zero));
#pragma unroll
    for (int u = 0; u < kPcenUnroll; ++u) {
      m = __fadd_rn(__fmul_rn(a.w, v[u]), __fmul_rn(a.one_minus_w, m));  // w*x + (1-w)*a, unfused like TF
      y[(size_t)(t + u) * a.inner] = m;
    }
  }
#endif
  for (; t < a.T; ++t) {
    m = __fadd_rn(__fmul_rn(a.w, x[(size_t)t * a.inner]), __fmul_rn(a.one_minus_w, m));
    y[(size_t)t * a.inner] = m;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// K3b: PCEN for time-contiguous data  x[rows][T][C]  with a small inner size C (the image layout [B][M][T][C] that
// raw_to_mel / get_spect produce and audiomodel.py:793 hands to the layer): the EMA as a warp-level PARALLEL SCAN.
//
// One warp per row.  The row (T * C floats, contiguous) is staged in shared memory with coalesced loads; every lane owns
// a contiguous run of time steps and
//   1. folds its run into the affine map  M_out = a * M_in + b   (a = (1-w)^n, b = the run's EMA from a zero state),
//   2. the 32 maps are combined with a shuffle scan  (a1, b1) o (a2, b2) = (a1 a2, a2 b1 + b2)  so that every lane
//      learns the smoother's state at the start of its run (the sequence starts from M = x[0], tfpcen.py:92),
//   3. walks its run again from that state with the reference's own update  w x + (1-w) M  and applies the gain /
//      bias / root compression, in place in shared memory; the row goes back to HBM with coalesced stores.
// The smoother's rounding order differs from the sequential tf.scan (survey probe: <= 0.05 of the tolerance budget).
// The per-lane kernel above walks these rows one 4-byte load per lane and sector (measured 0.7-1.7 TB/s); this one moves
// whole rows.  Same REDUCE / APPLY / RAW protocol for the min-max scopes.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kScanWarps = 8;
constexpr int kScanMaxRow = 2048;   // floats per row (T * C) the staging holds
constexpr int kScanMaxC = 4;

template <int MODE>
__global__ void __launch_bounds__(kScanWarps * 32) pcen_scan_kernel(const PcenArgs a, const long long rows,
                                                                     const long long rows_per_clip_) {
  extern __shared__ __align__(16) float s_rows[];   // [kScanWarps][T * C]
  __shared__ float scratch[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.inner, T = a.T, n = T * C;
  const long long row = (long long)blockIdx.x * kScanWarps + warp;
  float mn = INFINITY, mx = -INFINITY;
  if (row < rows) {
    float* s = s_rows + (size_t)warp * n;
    const float* x = a.in + (size_t)row * n;
    for (int i = lane; i < n; i += 32) s[i] = ld_stream(x + i);
    __syncwarp();
    float scale = 1.0f, shift = 0.0f;
    if (MODE == PCEN_APPLY) {
      const float2 e = a.extremes[a.per_clip_extremes ? (int)(row / rows_per_clip_) : 0];
      scale = 2.0f / e.x;
      if (fmaf(e.x, scale, -2.0f) < 0.0f) scale = __uint_as_float(__float_as_uint(scale) + 1u);
      shift = e.y;
    }
    const int per = (T + 31) / 32;                      // time steps per lane
    const int t_lo = min(lane * per, T), t_hi = min(t_lo + per, T);
    for (int c = 0; c < C; ++c) {
      const float x0 = s[c];   // initial state, read before any lane writes its results back in place (step 3)
      // 1. this lane's run as an affine map of the incoming state
      float fa = 1.0f, fb = 0.0f;
      for (int t = t_lo; t < t_hi; ++t) {
        fb = __fadd_rn(__fmul_rn(a.w, s[t * C + c]), __fmul_rn(a.one_minus_w, fb));
        fa *= a.one_minus_w;
      }
      // 2. inclusive scan of the maps over the lanes (earlier map applied first)
      float ia = fa, ib = fb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float pa = __shfl_up_sync(kFullMask, ia, o), pb = __shfl_up_sync(kFullMask, ib, o);
        if (lane >= o) {
          ib = fmaf(ia, pb, ib);
          ia *= pa;
        }
      }
      // exclusive: the composition of all earlier lanes, applied to the initial state x[0]
      float ea = __shfl_up_sync(kFullMask, ia, 1), eb = __shfl_up_sync(kFullMask, ib, 1);
      if (lane == 0) {
        ea = 1.0f;
        eb = 0.0f;
      }
      float m = fmaf(ea, x0, eb);
      __syncwarp();   // every lane has read its inputs of steps 1-2 before lane 0 overwrites s[c]
      // 3. the run again, from the right state, with the compression
      for (int t = t_lo; t < t_hi; ++t) {
        const float v = s[t * C + c];
        m = __fadd_rn(__fmul_rn(a.w, v), __fmul_rn(a.one_minus_w, m));
        float p = pcen_point(v, m, a);
        if (MODE == PCEN_REDUCE) {
          mn = min_nan(mn, p);
          mx = max_nan(mx, p);
        } else {
          if (MODE == PCEN_APPLY) p = max_nan(min_nan(fmaf(p - shift, scale, -1.0f), 1.0f), -1.0f);
          s[t * C + c] = p;
        }
      }
      __syncwarp();   // lane l's first sample of channel c + 1 may be ... (runs are disjoint; kept for the in-place writes)
    }
    if (MODE != PCEN_REDUCE) {
      __syncwarp();
      float* y = a.out + (size_t)row * n;
      for (int i = lane; i < n; i += 32) y[i] = s[i];
    }
  }
  if (MODE == PCEN_REDUCE) {
    block_minmax_nan(mn, mx, scratch);
    if (threadIdx.x == 0) a.partial[blockIdx.x] = make_float2(mn, mx);
  }
}

}  // namespace cacfe
