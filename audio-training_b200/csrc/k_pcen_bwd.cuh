// PCEN backward: gradients of tfpcen.PCEN.call (tfpcen.py:89-99, normalize_minmax :105-110) with
// respect to the input and to the four scalars the call uses (gain, bias, root, EMA smooth), sm_100a.
//
// Forward, per (outer, inner) lane along t (k_pcen.cuh):
//   M_t = w x_t + (1 - w) M_{t-1},  M_{-1} = x_0            w = clip(smooth, 0, 1)
//   s_t = (eps + M_t)^-g                                     g = min(gain, 1)
//   y_t = x_t s_t + b,   p_t = y_t^(1/r) - b^(1/r)           r = max(root, 1)
//   out = 2 (p - mn) / (mx - mn) - 1                         mn, mx over the scope (tensor / clip / none)
// Backward is what TensorFlow's autodiff does with that graph: the min-max contributes 2 / R to every element plus the
// sums that flow through reduce_min / reduce_max to the elements that attain them (ties share equally); the smoother's
// adjoint is the reverse-time recurrence  l_t = dM_t + (1 - w) l_{t+1}.
//
// Three launches (two when norm_scope = NONE), same thread <-> lane map as pcen_kernel:
//   (extremes: pcen_kernel<REDUCE> + pcen_extremes_kernel in raw (mn, mx) form -- the range alone does not give mx back exactly)
//   pcen_bwd_reduce_kernel   recomputes p (same instruction sequence as the forward: extremes compare bit for bit) and
//                            accumulates per block  sum G, sum G p, #(p == mn), #(p == mx)          reads x, G
//   pcen_bwd_fold_kernel     folds the partials into  2/R, dL/dmx / n_max, dL/dmn / n_min           tiny
//   pcen_bwd_kernel          reverse sweep.  M_t cannot be run backwards stably (division by 1 - w per step), so a
//                            first forward walk keeps the state entering every 16-step segment in shared memory; each
//                            segment is then replayed forward into registers and differentiated backwards.  Scalar
//                            gradients: per-thread float sums, block-reduced in double.            reads x (x2), G; writes dx
// T <= kBwdSeg * kBwdMaxSeg.
#pragma once
#include "k_pcen.cuh"

namespace cacfe {

constexpr int kScopeClip = 1, kScopeNone = 2;   // cacfe_norm_scope (include/cacfe.h)
constexpr int kBwdSeg = 16;         // time steps per replayed segment (the unrolled reverse loop: 16 x ~150 instructions; 32
                                    // steps did not fit the instruction cache: 1.8 of 9 stall cycles per issue were fetches)
constexpr int kBwdMaxSeg = 128;     // T <= 2048
constexpr int kBwdMaxThreads = 192; // checkpoints: n_seg * blockDim floats of dynamic shared memory

struct PcenBwdArgs {
  PcenArgs f;                  // forward constants; f.in = x, f.out unused
  const float* g;              // dL/dout
  float* dx;
  float gain_raw, root_raw, smooth_raw;   // unclipped: a clipped parameter gets no gradient
  int scope;                   // cacfe_norm_scope: 0 tensor, 1 clip, 2 none
  const float2* extremes;      // (mn, mx) per entry                       (reduce, fold, backward)
  double* partial;             // [entries... blocks][4]                    (reduce / backward)
  const float4* fold;          // per entry (2/R, dmx / n_max, dmn / n_min, 0)
};

__device__ __forceinline__ double block_sum_d(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < nwarp; ++i) s += scratch[i];
  return s;   // valid in thread 0
}

__global__ void __launch_bounds__(256) pcen_bwd_reduce_kernel(const PcenBwdArgs a) {
  __shared__ double scratch[8];
  const PcenArgs& f = a.f;
  const int clip = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float s0 = 0.0f, s1 = 0.0f, n_mn = 0.0f, n_mx = 0.0f;
  if (r < f.rows_per_clip) {
    const int o = r / f.inner, i = r - o * f.inner;
    const size_t base = ((size_t)clip * (f.rows_per_clip / f.inner) + o) * f.T * f.inner + i;
    const float2 e = a.extremes[a.scope == kScopeClip ? clip : 0];
    const float* x = f.in + base;
    const float* g = a.g + base;
    float m = x[0];
    auto take = [&](float v, float gv) {
      m = __fadd_rn(__fmul_rn(f.w, v), __fmul_rn(f.one_minus_w, m));
      const float p = pcen_point(v, m, f);
      s0 += gv;
      s1 = fmaf(gv, p - e.x, s1);
      n_mn += p == e.x ? 1.0f : 0.0f;
      n_mx += p == e.y ? 1.0f : 0.0f;
    };
    int t = 0;
    for (; t + 8 <= f.T; t += 8) {   // 16 loads in flight per thread
      float v[8], gv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = ld_stream(x + (size_t)(t + u) * f.inner);
        gv[u] = ld_stream(g + (size_t)(t + u) * f.inner);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) take(v[u], gv[u]);
    }
    for (; t < f.T; ++t) take(ld_stream(x + (size_t)t * f.inner), ld_stream(g + (size_t)t * f.inner));
  }
  double* out = a.partial + ((size_t)clip * gridDim.x + blockIdx.x) * 4;
  const double t0 = block_sum_d(s0, scratch), t1 = block_sum_d(s1, scratch), t2 = block_sum_d(n_mn, scratch),
               t3 = block_sum_d(n_mx, scratch);
  if (threadIdx.x == 0) {
    out[0] = t0;
    out[1] = t1;
    out[2] = t2;
    out[3] = t3;
  }
}

// out_t = 2 u_t / R - 1, u = p - mn, R = mx - mn:   dout/dmx = -2 u / R^2,   dout/dmn = -2 / R + 2 u / R^2
__global__ void pcen_bwd_fold_kernel(const double* __restrict__ partial, int per_entry, const float2* __restrict__ extremes,
                                     float4* __restrict__ fold) {
  if (threadIdx.x != 0) return;
  const double* p = partial + (size_t)blockIdx.x * per_entry * 4;
  double s0 = 0, s1 = 0, n_mn = 0, n_mx = 0;
  for (int i = 0; i < per_entry; ++i) {
    s0 += p[4 * i];
    s1 += p[4 * i + 1];
    n_mn += p[4 * i + 2];
    n_mx += p[4 * i + 3];
  }
  const float2 e = extremes[blockIdx.x];
  const double R = (double)(e.y - e.x);   // the forward's f32 range
  const double dmx = -2.0 * s1 / (R * R), dmn = -2.0 * s0 / R + 2.0 * s1 / (R * R);
  fold[blockIdx.x] = make_float4((float)(2.0 / R), (float)(dmx / fmax(n_mx, 1.0)), (float)(dmn / fmax(n_mn, 1.0)), 0.0f);
}

__global__ void __launch_bounds__(kBwdMaxThreads) pcen_bwd_kernel(const PcenBwdArgs a) {
  extern __shared__ float ckpt[];     // [n_seg][blockDim.x]
  __shared__ double scratch[8];
  const PcenArgs& f = a.f;
  const int clip = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float d_gain = 0.0f, d_bias = 0.0f, d_root = 0.0f, d_smooth = 0.0f;
  if (r < f.rows_per_clip) {
    const int o = r / f.inner, i = r - o * f.inner;
    const size_t base = ((size_t)clip * (f.rows_per_clip / f.inner) + o) * f.T * f.inner + i;
    const float* x = f.in + base;
    const float* g = a.g + base;
    float* dx = a.dx + base;
    float two_over_r = 1.0f, dmx = 0.0f, dmn = 0.0f, mn = 0.0f, mx = 0.0f;
    if (a.scope != kScopeNone) {
      const int ent = a.scope == kScopeClip ? clip : 0;
      const float4 fo = a.fold[ent];
      const float2 e = a.extremes[ent];
      two_over_r = fo.x;
      dmx = fo.y;
      dmn = fo.z;
      mn = e.x;
      mx = e.y;
    }
    const float ln2 = 0.6931471805599453f;
    const float ln_b = __logf(f.bias);
    const float dpdb = f.inv_root * f.bias_pow / f.bias;   // d(b^(1/r))/db
    const int n_seg = (f.T + kBwdSeg - 1) / kBwdSeg;
    // ---- forward walk: the smoother state entering every segment ----------------------------------------------------
    {
      float m = x[0];
      for (int s = 0; s < n_seg; ++s) {
        ckpt[s * blockDim.x + threadIdx.x] = m;
        float v[kBwdSeg];   // the segment's loads are all issued before the (sequential) fold
#pragma unroll
        for (int u = 0; u < kBwdSeg; ++u) v[u] = s * kBwdSeg + u < f.T ? x[(size_t)(s * kBwdSeg + u) * f.inner] : 0.0f;
#pragma unroll
        for (int u = 0; u < kBwdSeg; ++u) m = __fadd_rn(__fmul_rn(f.w, v[u]), __fmul_rn(f.one_minus_w, m));   // steps past T only follow the last checkpoint
      }
    }
    // ---- reverse sweep, one segment at a time -----------------------------------------------------------------------
    float lam = 0.0f;   // l_{t+1}
    for (int s = n_seg - 1; s >= 0; --s) {
      const int t_lo = s * kBwdSeg;
      const float m_in = ckpt[s * blockDim.x + threadIdx.x];
      // x and dL/dout of the whole segment are loaded up front: inside the reverse loop every load would queue behind the
      // previous step's dx store (the compiler must assume they alias) and expose one memory latency per time step
      float xv[kBwdSeg], mv[kBwdSeg], gvv[kBwdSeg];
#pragma unroll
      for (int u = 0; u < kBwdSeg; ++u) {
        const bool in = t_lo + u < f.T;
        xv[u] = in ? x[(size_t)(t_lo + u) * f.inner] : 0.0f;
        gvv[u] = in ? ld_stream(g + (size_t)(t_lo + u) * f.inner) : 0.0f;
      }
      {
        float m = m_in;
#pragma unroll
        for (int u = 0; u < kBwdSeg; ++u) {
          m = __fadd_rn(__fmul_rn(f.w, xv[u]), __fmul_rn(f.one_minus_w, m));
          mv[u] = m;
        }
      }
#pragma unroll
      for (int u = kBwdSeg - 1; u >= 0; --u) {
        const int t = t_lo + u;
        if (t < f.T) {
          const float gv = gvv[u];
          const float xt = xv[u], m = mv[u];
          const float p = pcen_point(xt, m, f);           // the forward's value, bit for bit
          float dp = gv * two_over_r;
          if (a.scope != kScopeNone) dp += (p == mx ? dmx : 0.0f) + (p == mn ? dmn : 0.0f);
          // local derivatives
          const float base_m = f.eps + m;
          const float lg = __log2f(base_m);
          const float sm = exp2f(-f.gain * lg);            // s_t
          const float y = fmaf(xt, sm, f.bias);
          const float pw = p + f.bias_pow;                 // y^(1/r)
          const float dpdy = __fdividef(f.inv_root * pw, y);
          const float dy = dp * dpdy;
          d_bias += dy - dp * dpdb;
          d_root += dp * (ln_b * f.bias_pow - ln2 * __log2f(y) * pw) * f.inv_root * f.inv_root;
          const float ds = dy * xt;
          d_gain -= ds * ln2 * lg * sm;
          const float dm = __fdividef(-f.gain * ds * sm, base_m);
          lam = fmaf(f.one_minus_w, lam, dm);              // l_t
          const float m_prev = u > 0 ? mv[u - 1] : m_in;
          d_smooth = fmaf(lam, xt - m_prev, d_smooth);
          float dxt = fmaf(dy, sm, f.w * lam);
          if (t == 0) dxt = fmaf(f.one_minus_w, lam, dxt); // M_{-1} = x_0
          dx[(size_t)t * f.inner] = dxt;
        }
      }
    }
  }
  double* out = a.partial + ((size_t)clip * gridDim.x + blockIdx.x) * 4;
  const double t0 = block_sum_d(d_gain, scratch), t1 = block_sum_d(d_bias, scratch), t2 = block_sum_d(d_root, scratch),
               t3 = block_sum_d(d_smooth, scratch);
  if (threadIdx.x == 0) {
    out[0] = t0;
    out[1] = t1;
    out[2] = t2;
    out[3] = t3;
  }
}

// Sum the block partials in a fixed order; clipped parameters get no gradient (tf.minimum / tf.maximum / clip_by_value
// pass the gradient only while the parameter is the active argument).
__global__ void pcen_bwd_params_kernel(const double* __restrict__ partial, int n, float gain_raw, float root_raw, float smooth_raw,
                                       float* __restrict__ grad_params) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) s[k] += partial[4 * (size_t)i + k];
  grad_params[0] = gain_raw <= 1.0f ? (float)s[0] : 0.0f;
  grad_params[1] = (float)s[1];
  grad_params[2] = root_raw >= 1.0f ? (float)s[2] : 0.0f;
  grad_params[3] = (smooth_raw >= 0.0f && smooth_raw <= 1.0f) ? (float)s[3] : 0.0f;
}

}  // namespace cacfe
