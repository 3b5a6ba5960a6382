// identifytracks.signal_noise (identifytracks.py:51-143) on the device, from the magnitude spectrogram to the list of
// connected components.  Integer / byte work: every stage is bit-exact against numpy / OpenCV on the
// same spectrogram.
//
//   a_max = max(S);  S' = S / a_max                                        (:79-80)
//   row_medians = median(S', axis=1), column_medians = median(S', axis=0)   (:81-82)   select_median_kernel
//   signal = (S' > 2 column_medians) & (S' > 3 row_medians)                 (:91)      signal_mask_kernel
//   open 4x4, dilate (height x width), erode (height/10 x width)            (:94-101)  morph_pass_kernel (separable)
//   connectedComponentsWithStats, 8-connectivity                            (:106)     ccl_* kernels (union-find)
//
// Medians: exact order statistics by radix select on the IEEE bit pattern (4 passes of 8 bits over an L2-resident row);
// dividing by a positive constant is monotone, so the k-th smallest of S' is (k-th smallest of S) / a_max, the same f32
// division numpy performs on that element; for an even count numpy averages the two middle values in f32.
// OpenCV conventions (checked against cv2 4.13): rectangular structuring element, anchor = size / 2, pixels outside the
// image never win (erode: +inf, dilate: -inf); an EMPTY kernel -- np.ones((0, width)), what `height // 10` gives for
// height 6 -- means the default 3 x 3 rectangle; component labels are numbered by the first 2 x 2 block of the component
// in block-raster order (the Spaghetti / BBDT scan), which is what `order_key` reproduces for the stable sort by x (:111).
#pragma once
#include "cacfe_common.cuh"
#include "k_compress.cuh"

namespace cacfe {

// ---- exact median of each contiguous row ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_key(float v) {   // monotone map float -> uint32 (handles negatives)
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest key (0-based) of x[0..n): 4 radix passes.  Block-wide; result valid in every thread.
__device__ uint32_t block_select(const float* __restrict__ x, int n, int k, uint32_t* hist /* [256] */, uint32_t* bcast /* [2] */) {
  uint32_t prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = float_key(x[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t acc = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] > (uint32_t)k) break;
        acc += hist[b];
      }
      bcast[0] = (uint32_t)b;
      bcast[1] = acc;
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 255u << shift;
    k -= (int)bcast[1];
    __syncthreads();
  }
  return prefix;
}

// grid = rows, block = 256.  out[row] = np.median(x[row] / amax) in f32 arithmetic.
__global__ void __launch_bounds__(256) select_median_kernel(const float* __restrict__ in, int n, const Stats* __restrict__ stats,
                                                            float* __restrict__ out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bcast[2];
  __shared__ uint32_t red[8];
  const float* x = in + (size_t)blockIdx.x * n;
  const float amax = stats[0].mx;
  const int k_hi = n / 2;
  const uint32_t key_hi = block_select(x, n, k_hi, hist, bcast);
  const float v_hi = key_float(key_hi);
  float med = __fdiv_rn(v_hi, amax);
  if ((n & 1) == 0) {
    // lower middle = largest element below position k_hi in sorted order: v_hi itself if it occurs more than once at or below
    // that position, else the largest key smaller than key_hi
    uint32_t below = 0, best = 0;   // count of keys < key_hi, and the largest of them
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = float_key(x[i]);
      if (key < key_hi) {
        ++below;
        best = max(best, key);
      }
    }
    // block reduce (sum, max)
    for (int o = 16; o > 0; o >>= 1) {
      below += __shfl_xor_sync(kFullMask, below, o);
      best = max(best, __shfl_xor_sync(kFullMask, best, o));
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      hist[threadIdx.x >> 5] = below;
      red[threadIdx.x >> 5] = best;
    }
    __syncthreads();
    uint32_t tot = 0, bst = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      tot += hist[w];
      bst = max(bst, red[w]);
    }
    const float v_lo = (tot == (uint32_t)k_hi) ? key_float(bst) : v_hi;   // exactly k_hi keys below: position k_hi - 1 holds the largest of them
    // numpy: mean of the two middle elements of the (already divided) f32 array
    med = __fmul_rn(__fadd_rn(__fdiv_rn(v_lo, amax), med), 0.5f);
  }
  if (threadIdx.x == 0) out[blockIdx.x] = med;
}

// signal[k][t] = (S' > 2 cm[t]) & (S' > 3 rm[k]),  S' = S / a_max in f32
__global__ void __launch_bounds__(256) signal_mask_kernel(const float* __restrict__ spec, int K, int T, const Stats* __restrict__ stats,
                                                          const float* __restrict__ rm, const float* __restrict__ cm,
                                                          unsigned char* __restrict__ mask) {
  const long long n = (long long)K * T;
  const float amax = stats[0].mx;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i / T), t = (int)(i - (long long)k * T);
    const float s = __fdiv_rn(spec[i], amax);
    mask[i] = (s > __fmul_rn(2.0f, cm[t]) && s > __fmul_rn(3.0f, rm[k])) ? 1 : 0;
  }
}

// One separable pass of a rectangular erode (IS_MAX = false) / dilate (true): window [-(anchor), size - anchor) along x
// (horizontal) or y.  Pixels outside the image are ignored.
template <bool IS_MAX>
__global__ void __launch_bounds__(256) morph_pass_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int H, int W,
                                                         int size, int anchor, int horizontal) {
  const long long n = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (long long)y * W);
    unsigned char r = IS_MAX ? 0 : 255;
    if (horizontal) {
      const int lo = max(0, x - anchor), hi = min(W, x - anchor + size);
      const unsigned char* p = in + (size_t)y * W;
      for (int j = lo; j < hi; ++j) r = IS_MAX ? max(r, p[j]) : min(r, p[j]);
    } else {
      const int lo = max(0, y - anchor), hi = min(H, y - anchor + size);
      for (int j = lo; j < hi; ++j) {
        const unsigned char v = in[(size_t)j * W + x];
        r = IS_MAX ? max(r, v) : min(r, v);
      }
    }
    out[i] = r;
  }
}

// ---- connected components, 8-connectivity: union-find on the pixel grid -------------------------------------------
struct Component {   // cv2 stats row + ordering key
  int min_x, min_y, max_x, max_y, area, order_key;
};

__device__ __forceinline__ int ccl_find(const int* L, int a) {
  int p = L[a];
  while (p != a) {
    a = p;
    p = L[a];
  }
  return a;
}
__device__ __forceinline__ void ccl_union(int* L, int a, int b) {
  while (true) {
    a = ccl_find(L, a);
    b = ccl_find(L, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&L[a], b);   // a > b: hang the larger root under the smaller
    if (old == a) return;
    a = old;
  }
}

// Initial labels: every foreground pixel points at the first pixel of its horizontal run inside the warp's 32 consecutive
// pixels (ballot + bit scan).  A run that continues from the previous warp is joined by one union in ccl_merge_kernel.
// Both kernels walk the image with the same warp <-> pixel map (grid-stride in multiples of the block size).
__device__ __forceinline__ bool ccl_left_link(const unsigned char* mask, long long i, long long n, int W, bool& fg, int& x, int& y) {
  fg = i < n && mask[i] != 0;
  y = (int)(i / W);
  x = (int)(i - (long long)y * W);
  return fg && x > 0 && mask[i - 1] != 0;
}
__global__ void __launch_bounds__(256) ccl_init_kernel(const unsigned char* __restrict__ mask, int* __restrict__ L, int H, int W) {
  const long long n = (long long)H * W, span = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n; i0 += span) {
    const long long i = i0 + threadIdx.x;
    bool fg;
    int x, y;
    const bool link = ccl_left_link(mask, i, n, W, fg, x, y);
    const unsigned links = __ballot_sync(kFullMask, link);
    if (i < n) {
      const unsigned starts = ~links & ((2u << lane) - 1u);           // lanes <= this one that begin a run (or are background)
      const int first = starts ? 31 - __clz(starts) : 0;
      L[i] = fg ? (int)(i - (lane - first)) : -1;
    }
  }
}
// Unions across rows, pruned to the places where a run meets something new in the row above: with u / ul / ur the pixels
// above, above-left and above-right,  (i, u) unless the left neighbour already joined ul (ul and u are neighbours in their
// own row);  if u is background: (i, ul) unless the left neighbour -- whose u that is -- exists, and (i, ur).
__global__ void __launch_bounds__(256) ccl_merge_kernel(const unsigned char* __restrict__ mask, int* __restrict__ L, int H, int W) {
  const long long n = (long long)H * W, span = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n; i0 += span) {
    const long long i = i0 + threadIdx.x;
    bool fg;
    int x, y;
    const bool left = ccl_left_link(mask, i, n, W, fg, x, y);
    if (!fg) continue;
    if (lane == 0 && left) ccl_union(L, (int)i, (int)i - 1);          // the run continues from the previous warp
    if (y == 0) continue;
    const long long up = i - W;
    const bool u = mask[up] != 0, ul = x > 0 && mask[up - 1] != 0, ur = x + 1 < W && mask[up + 1] != 0;
    if (u) {
      if (!(left && ul)) ccl_union(L, (int)i, (int)up);
    } else {
      if (ul && !left) ccl_union(L, (int)i, (int)up - 1);
      if (ur) ccl_union(L, (int)i, (int)up + 1);
    }
  }
}
// flatten, and give every root a slot in the component list (slot_of[root] = index; counter[0] = number of components)
__global__ void __launch_bounds__(256) ccl_flatten_kernel(int* __restrict__ L, long long n, int* __restrict__ slot_of, int* __restrict__ counter,
                                                          int max_components) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (L[i] < 0) continue;
    const int r = ccl_find(L, (int)i);
    L[i] = r;   // path compression: any ancestor is a valid parent for concurrent readers
    if (r == (int)i) {
      const int s = atomicAdd(counter, 1);
      slot_of[i] = s < max_components ? s : -1;
    }
  }
}
__global__ void __launch_bounds__(256) ccl_stats_kernel(const int* __restrict__ L, int H, int W, const int* __restrict__ slot_of,
                                                        Component* __restrict__ comps) {
  const long long n = (long long)H * W;
  const long long span = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n; i0 += span) {   // whole warps stay in the loop together
    const long long i = i0 + threadIdx.x;
    int s = -1, x = 0, y = 0;
    if (i < n && L[i] >= 0) {
      s = slot_of[ccl_find(L, (int)i)];
      y = (int)(i / W);
      x = (int)(i - (long long)y * W);
    }
    // lanes of a warp that belong to the same component and the same image row fold into one update
    const unsigned peers = __match_any_sync(kFullMask, ((unsigned long long)(unsigned)s << 32) | (unsigned)y);
    if (s < 0) continue;
    const int lane = threadIdx.x & 31;
    const int first = __ffs(peers) - 1, last = 31 - __clz(peers);
    const int x_hi = __shfl_sync(peers, x, last);
    if (lane == first) {
      Component* c = comps + s;
      atomicMin(&c->min_x, x);
      atomicMax(&c->max_x, x_hi);
      atomicMin(&c->min_y, y);
      atomicMax(&c->max_y, y);
      atomicAdd(&c->area, __popc(peers));
      atomicMin(&c->order_key, (y >> 1) * ((W + 1) >> 1) + (x >> 1));
    }
  }
}
__global__ void ccl_clear_kernel(Component* __restrict__ comps, int max_components, int* __restrict__ counter) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max_components; i += gridDim.x * blockDim.x) {
    Component c;
    c.min_x = c.min_y = c.order_key = 0x7fffffff;
    c.max_x = c.max_y = -1;
    c.area = 0;
    comps[i] = c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) counter[0] = 0;
}

}  // namespace cacfe
