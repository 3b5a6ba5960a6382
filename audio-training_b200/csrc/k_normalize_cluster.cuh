// Standalone normalize (a1: tfdataset.normalize, tfdataset.py:1916-1934; predict_utils.normalize_data :153-160) with the
// clip held ON CHIP between the two passes: a thread-block cluster of 8 CTAs owns one clip (576 KB for 3 s @ 48 kHz =
// 72 KB per CTA), so HBM sees one read and one write per sample instead of two reads (min/max, then the rescale) and
// one write.
//   * every CTA brings its eighth of the clip into shared memory with TMA bulk copies (one mbarrier, 72 KB in flight per
//     CTA, three CTAs resident per SM so loads, the shared-memory passes and stores of different clips overlap);
//   * local min / max from shared memory, published in the CTA's shared memory; barrier.cluster; every CTA reads the eight
//     pairs through distributed shared memory (mapa / ld.shared::cluster via cooperative_groups);
//   * the reference's four f32 operations per sample (true division), shared memory -> global, 16-byte stores.
// Same arithmetic, operation for operation, as row_normalize_kernel: results are bit-identical.  Clips whose eighth does
// not fit (n > 8 * 50 K samples), n % 32 != 0 or unaligned rows keep the two-kernel path.
#pragma once
#include <cooperative_groups.h>

#include "cacfe_async.cuh"
#include "cacfe_common.cuh"

namespace cacfe {

constexpr int kNcCluster = 8;
constexpr int kNcThreads = 256;
constexpr int kNcMaxBytesPerCta = 200 * 1024;

__global__ void __cluster_dims__(kNcCluster, 1, 1) __launch_bounds__(kNcThreads)
    row_normalize_cluster_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int per_cta) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(128) unsigned char smem[];
  float4* s4 = reinterpret_cast<float4*>(smem);
  __shared__ float2 s_mm;
  __shared__ float2 s_all;
  __shared__ uint64_t s_bar;
  __shared__ float scratch[64];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const long long row = blockIdx.x / kNcCluster;
  const float* x = in + row * n + (long long)rank * per_cta;
  float* y = out + row * n + (long long)rank * per_cta;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)per_cta * 4u, bar = smem_u32(&s_bar);
    mbar_expect_tx(bar, bytes);
    const uint32_t piece = 16 * 1024;
    for (uint32_t o = 0; o < bytes; o += piece)
      bulk_g2s(smem_u32(smem + o), reinterpret_cast<const unsigned char*>(x) + o, min(piece, bytes - o), bar);
  }
  mbar_wait(smem_u32(&s_bar), 0);
  const int n4 = per_cta >> 2;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = tid; i < n4; i += kNcThreads) {
    const float4 v = s4[i];
    mn = min_nan(mn, min_nan(min_nan(v.x, v.y), min_nan(v.z, v.w)));
    mx = max_nan(mx, max_nan(max_nan(v.x, v.y), max_nan(v.z, v.w)));
  }
  block_minmax_nan(mn, mx, scratch);
  if (tid == 0) s_mm = make_float2(mn, mx);
  cluster.sync();
  if (tid < 32) {
    float a = INFINITY, b = -INFINITY;
    if (tid < kNcCluster) {
      const float2 v = *cluster.map_shared_rank(&s_mm, tid);
      a = v.x;
      b = v.y;
    }
    a = warp_min_nan(a);
    b = warp_max_nan(b);
    if (tid == 0) s_all = make_float2(a, b);
  }
  cluster.sync();   // nobody leaves (or overwrites s_mm) while a peer may still be reading it
  mn = s_all.x;
  const float range = s_all.y - mn;   // == max(x - mn): rounding is monotone
  auto point = [&](float s) -> float {   // the reference's f32 order, one rounding per operation
    float v = s - mn;
    v = __fadd_rn(__fdiv_rn(v, range), 0.000001f);
    v = __fsub_rn(v, 0.5f);
    return __fmul_rn(v, 2.0f);
  };
  float4* y4 = reinterpret_cast<float4*>(y);
  for (int i = tid; i < n4; i += kNcThreads) {
    const float4 v = s4[i];
    y4[i] = make_float4(point(v.x), point(v.y), point(v.z), point(v.w));
  }
}

}  // namespace cacfe
