// Point-wise compression / normalisation epilogues with a tensor- or clip-wide statistic, sm_100a.
//
//   MINMAX       normalize_minmax  tfpcen.py:105-110, tfdataset.py:1897-1902    2*((x-min)/(max-min)) - 1
//   POWER_TO_DB  power_to_db       tfdataset.py:1906-1913 (== librosa.power_to_db(ref=np.max), predict_utils.py:216)
//   STD          normalize_std     tfdataset.py:1883-1893                        (x-mean)/(std + 1e-7)
//   MAG_POW      MagTransform      badwinner2.py:32-49                           x ** sigmoid(a)   (no statistic)
//   MEAN_SUB     get_spect(mean_sub=True)  predict_utils.py:233-236              x - mean   (one entry per mel row)
//
// The tensor is a flat run of `per_entry` floats per scope entry (1 entry = whole tensor, or B = per clip).
// Pass 1 (stats_kernel) writes block partials, stats_finalize_kernel folds them, pass 2 applies.  HBM bound.
#pragma once
#include "cacfe_common.cuh"

namespace cacfe {

enum : int { COMPRESS_MAG_POW = 0, COMPRESS_POWER_TO_DB = 1, COMPRESS_MINMAX = 2, COMPRESS_STD = 3, COMPRESS_MEAN_SUB = 4 };

struct Stats {  // one per scope entry
  float mn, mx;
  double sum, sumsq;
};

// grid = (blocks_per_entry, entries), block = 256
__global__ void __launch_bounds__(256) stats_kernel(const float* __restrict__ in, long long per_entry,
                                                    Stats* __restrict__ partial) {
  __shared__ float scratch[64];
  __shared__ double dscratch[16];
  const float* x = in + (size_t)blockIdx.y * per_entry;
  float mn = INFINITY, mx = -INFINITY;
  double s = 0.0, ss = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  auto take = [&](float v) {
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
    s += (double)v;
    ss += (double)v * (double)v;
  };
  if ((per_entry & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {  // 16-byte loads, four in flight per thread
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const long long n4 = per_entry >> 2;
#pragma unroll 4
    for (long long i = first; i < n4; i += stride) {
      const float4 v = ld_stream4(x4 + i);
      take(v.x);
      take(v.y);
      take(v.z);
      take(v.w);
    }
  } else {
#pragma unroll 4
    for (long long i = first; i < per_entry; i += stride) take(ld_stream(x + i));
  }
  block_minmax(mn, mx, scratch);
  s = warp_sum(s);
  ss = warp_sum(ss);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    dscratch[warp] = s;
    dscratch[8 + warp] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += dscratch[w];
      b += dscratch[8 + w];
    }
    Stats st;
    st.mn = mn;
    st.mx = mx;
    st.sum = a;
    st.sumsq = b;
    partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = st;
  }
}

// grid = entries, block = 32
__global__ void stats_finalize_kernel(const Stats* __restrict__ partial, int per_entry_blocks, Stats* __restrict__ out) {
  const Stats* p = partial + (size_t)blockIdx.x * per_entry_blocks;
  float mn = INFINITY, mx = -INFINITY;
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < per_entry_blocks; i += 32) {
    mn = fminf(mn, p[i].mn);
    mx = fmaxf(mx, p[i].mx);
    s += p[i].sum;
    ss += p[i].sumsq;
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (threadIdx.x == 0) {
    Stats st;
    st.mn = mn;
    st.mx = mx;
    st.sum = s;
    st.sumsq = ss;
    out[blockIdx.x] = st;
  }
}

// grid = (blocks_per_entry, entries), block = 256
template <int MODE>
__global__ void __launch_bounds__(256) compress_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       long long per_entry, float param,
                                                       const Stats* __restrict__ stats) {
  const float* x = in + (size_t)blockIdx.y * per_entry;
  float* y = out + (size_t)blockIdx.y * per_entry;
  float c0 = 0.0f, c1 = 0.0f;
  if (MODE == COMPRESS_MINMAX) {
    const Stats st = stats[blockIdx.y];
    c0 = st.mn;
    c1 = st.mx - st.mn;
  } else if (MODE == COMPRESS_POWER_TO_DB) {
    const Stats st = stats[blockIdx.y];
    c0 = 10.0f * log10f(fmaxf(1e-10f, st.mx));  // 10 log10(max(amin, ref)); max of the dB image is exactly 0
  } else if (MODE == COMPRESS_STD) {
    const Stats st = stats[blockIdx.y];
    const double mean = st.sum / (double)per_entry;
    const double var = fmax(st.sumsq / (double)per_entry - mean * mean, 0.0);
    c0 = (float)mean;
    c1 = (float)sqrt(var) + 1e-7f;  // keras.backend.epsilon()
  } else if (MODE == COMPRESS_MEAN_SUB) {
    c0 = (float)(stats[blockIdx.y].sum / (double)per_entry);   // the correctly rounded f32 mean (tf.reduce_mean sums in f32)
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  auto point = [&](float v) -> float {
    if (MODE == COMPRESS_MAG_POW) return exp2f(param * log2f(v));  // v ** param for v >= 0 (0 -> 0, like tf.pow)
    if (MODE == COMPRESS_POWER_TO_DB) return fmaxf(10.0f * log10f(fmaxf(1e-10f, v)) - c0, -80.0f);
    if (MODE == COMPRESS_MINMAX) return 2.0f * ((v - c0) / c1) - 1.0f;
    if (MODE == COMPRESS_MEAN_SUB) return v - c0;
    return (v - c0) / c1;
  };
  if ((per_entry & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    const long long n4 = per_entry >> 2;
#pragma unroll 4
    for (long long i = first; i < n4; i += stride) {
      const float4 v = ld_stream4(x4 + i);
      y4[i] = make_float4(point(v.x), point(v.y), point(v.z), point(v.w));
    }
  } else {
#pragma unroll 4
    for (long long i = first; i < per_entry; i += stride) y[i] = point(ld_stream(x + i));
  }
}

// tfdataset.mix_up (tfdataset.py:948): images_one * l + images_two * (1 - l) with one l per batch entry, the reference's f32
// operation order (three roundings).  grid = (blocks, B), block = 256.
__global__ void __launch_bounds__(256) mix_up_kernel(const float* __restrict__ one, const float* __restrict__ two,
                                                     const float* __restrict__ lam, float* __restrict__ out, long long per_entry) {
  const float l = lam[blockIdx.y], r = __fsub_rn(1.0f, l);
  const float* a = one + (size_t)blockIdx.y * per_entry;
  const float* b = two + (size_t)blockIdx.y * per_entry;
  float* y = out + (size_t)blockIdx.y * per_entry;
  auto mix = [&](float u, float v) { return __fadd_rn(__fmul_rn(u, l), __fmul_rn(v, r)); };
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if ((per_entry & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
#pragma unroll 4
    for (long long i = first; i < (per_entry >> 2); i += stride) {
      const float4 u = ld_stream4(reinterpret_cast<const float4*>(a) + i), v = ld_stream4(reinterpret_cast<const float4*>(b) + i);
      reinterpret_cast<float4*>(y)[i] = make_float4(mix(u.x, v.x), mix(u.y, v.y), mix(u.z, v.z), mix(u.w, v.w));
    }
  } else {
    for (long long i = first; i < per_entry; i += stride) y[i] = mix(a[i], b[i]);
  }
}

// 16-bit PCM -> float32, the conversion libsndfile / soundfile / librosa.load apply to a 16-bit file: s / 32768 (exact in
// FP32).  16-byte loads of eight samples, two 16-byte stores; a scalar tail for lengths that are not a multiple of 8.
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const short* __restrict__ in, float* __restrict__ out, long long n) {
  const long long n8 = (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) ? n >> 3 : 0;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const int4* in8 = reinterpret_cast<const int4*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  constexpr float k = 1.0f / 32768.0f;
  for (long long i = first; i < n8; i += stride) {
    const int4 v = __ldcs(in8 + i);
    auto lo = [](int w) { return (float)(short)(w & 0xffff) * k; };
    auto hi = [](int w) { return (float)(w >> 16) * k; };
    out4[2 * i] = make_float4(lo(v.x), hi(v.x), lo(v.y), hi(v.y));       // read again right away by K0 / K1: plain stores
    out4[2 * i + 1] = make_float4(lo(v.z), hi(v.z), lo(v.w), hi(v.w));
  }
  for (long long i = (n8 << 3) + first; i < n; i += stride) out[i] = (float)in[i] * k;
}

}  // namespace cacfe
