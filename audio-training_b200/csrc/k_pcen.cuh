// K3/K4: PCEN (EMA smoother + gain/bias/root compression) and the min-max epilogue, sm_100a.
//
// Replaces tfpcen.py:8-39 (ExponentialMovingAverage: tf.scan over time), :89-95 (PCEN.call) and
// :105-110 (tensor-global normalize_minmax).
//
// Data is viewed as x[outer][T][inner] with the recurrence along T:
//   reference contract [B, T, F]        -> outer = B,       inner = F   (lanes across F: coalesced rows)
//   image layout       [B, M, T, C] (*) -> outer = B * M,   inner = C   (* our rank-4 extension)
// Each thread owns one (outer, inner) lane and walks T sequentially -- the reference's exact EMA order --
// with kUnroll independent loads in flight.  HBM bound: 1 read (+1 write) of 4 B per element and pass.
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
};

enum : int { PCEN_REDUCE = 0, PCEN_APPLY = 1, PCEN_RAW = 2 };

__device__ __forceinline__ float pcen_point(float x, float m, const PcenArgs& a) {
  // x / (eps + M)^gain  ==  x * 2^(-gain * log2(eps + M))      (MUFU lg2 / ex2)
  const float smooth = exp2f(-a.gain * __log2f(a.eps + m));
  const float y = fmaf(x, smooth, a.bias);
  // root 2 (the layer's initial value): y * rsqrt(y) -- one MUFU and one multiply, <= 2 ulp -- instead of the IEEE sqrt
  // sequence; the pass is MUFU / issue bound, not HBM bound, until these are trimmed
  const float r = a.root_is_2 ? (y > 0.0f ? y * rsqrtf(y) : 0.0f) : exp2f(a.inv_root * __log2f(y));
  return r - a.bias_pow;
}

constexpr int kPcenUnroll = 8;

template <int MODE>
__global__ void __launch_bounds__(256) pcen_kernel(const PcenArgs a) {
  __shared__ float scratch[64];
  const int clip = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;  // row inside the clip
  const bool live = r < a.rows_per_clip;
  float mn = INFINITY, mx = -INFINITY;
  if (live) {
    const int o = r / a.inner, i = r - o * a.inner;
    const size_t base = ((size_t)clip * (a.rows_per_clip / a.inner) + o) * a.T * a.inner + i;
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
    int t = 0;
    for (; t + kPcenUnroll <= a.T; t += kPcenUnroll) {
      float v[kPcenUnroll];
#pragma unroll
      for (int u = 0; u < kPcenUnroll; ++u) v[u] = ld_stream(x + (size_t)(t + u) * a.inner);
#pragma unroll
      for (int u = 0; u < kPcenUnroll; ++u) {
        m = __fadd_rn(__fmul_rn(a.w, v[u]), __fmul_rn(a.one_minus_w, m));  // unfused, the reference's f32 order
        float p = pcen_point(v[u], m, a);
        if (MODE == PCEN_REDUCE) {
          mn = fminf(mn, p);
          mx = fmaxf(mx, p);
        } else {
          if (MODE == PCEN_APPLY) p = fminf(fmaf(p - shift, scale, -1.0f), 1.0f);
          y[(size_t)(t + u) * a.inner] = p;
        }
      }
    }
    for (; t < a.T; ++t) {
      const float v = ld_stream(x + (size_t)t * a.inner);
      m = __fadd_rn(__fmul_rn(a.w, v), __fmul_rn(a.one_minus_w, m));
      float p = pcen_point(v, m, a);
      if (MODE == PCEN_REDUCE) {
        mn = fminf(mn, p);
        mx = fmaxf(mx, p);
      } else {
        if (MODE == PCEN_APPLY) p = fminf(fmaf(p - shift, scale, -1.0f), 1.0f);
        y[(size_t)t * a.inner] = p;
      }
    }
  }
  if (MODE == PCEN_REDUCE) {
    block_minmax(mn, mx, scratch);
    if (threadIdx.x == 0) a.partial[(size_t)clip * gridDim.x + blockIdx.x] = make_float2(mn, mx);
  }
}

// Fold block partials into (range, min) per scope entry.  grid = entries, block = 256.
__global__ void __launch_bounds__(256) minmax_finalize_kernel(const float2* __restrict__ partial, int per_entry,
                                                              float2* __restrict__ extremes) {
  __shared__ float scratch[64];
  const float2* p = partial + (size_t)blockIdx.x * per_entry;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < per_entry; i += blockDim.x) {
    mn = fminf(mn, p[i].x);
    mx = fmaxf(mx, p[i].y);
  }
  block_minmax(mn, mx, scratch);
  if (threadIdx.x == 0) extremes[blockIdx.x] = make_float2(mx - mn, mn);  // (range, min)
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
  float m = x[0];
  int t = 0;
  for (; t + kPcenUnroll <= a.T; t += kPcenUnroll) {
    float v[kPcenUnroll];
#pragma unroll
    for (int u = 0; u < kPcenUnroll; ++u) v[u] = ld_stream(x + (size_t)(t + u) * a.inner);
#pragma unroll
    for (int u = 0; u < kPcenUnroll; ++u) {
      m = __fadd_rn(__fmul_rn(a.w, v[u]), __fmul_rn(a.one_minus_w, m));  // w*x + (1-w)*a, unfused like TF
      y[(size_t)(t + u) * a.inner] = m;
    }
  }
  for (; t < a.T; ++t) {
    m = __fadd_rn(__fmul_rn(a.w, x[(size_t)t * a.inner]), __fmul_rn(a.one_minus_w, m));
    y[(size_t)t * a.inner] = m;
  }
}

}  // namespace cacfe
