// Path C on the 5th-generation tensor cores: stored spectrogram [B][K][T] -> mel, as a banded 3xTF32 GEMM (sm_100a).
//
// Replaces tf.tensordot(MEL_WEIGHTS, spectrogram(2049x513), 1) (tfdataset.py:1082-1090) -- the north star's "mel
// projection as a dense contraction on tensor cores (TF32 / FP32 accumulate)".  A single TF32 pass misses the 1e-4
// tolerance by 5-7x, so both operands are split  x = hi + lo  (hi = x with the low 13
// mantissa bits cleared, lo = x - hi, exact) and three MMAs accumulate  hi*hi + lo*hi + hi*lo  in FP32.
//
// One CTA computes one output tile D[t = 128 frames][m = n_mels] of one clip:
//   A = S^T tile: 128 frames x 32 bins per K chunk.  The stored rows have a pitch of T*4 = 2052 bytes -- not a multiple
//     of 16, so neither TMA tensor maps nor bulk copies can address them: eight producer warps load with coalesced LDG
//     (lanes along t), square if power == 2, split and write the hi / lo operands -- transposed -- straight into the
//     UMMA canonical K-major no-swizzle layout [k / 4][t / 8][t % 8][k % 4] (8 x 16-byte core matrices), four bins per
//     16-byte store.  (An M-major A needs no transpose but a no-swizzle MN-major TF32 operand multiplies as zero on
//     sm_100a -- probed with tools/tc_probe.cu -- so the transpose is done by the producers.)
//   B (K-major) = filterbank chunk.  The bank is banded: a 32-bin chunk touches 16..64 adjacent mel bands, so the plan
//     stores per chunk only those rows (hi and lo, already in canonical layout) and the MMA runs with N = that count
//     into the matching TMEM columns: 5.6x less L2 traffic and tensor work than the dense product.  One
//     cp.async.bulk per chunk (TMA, mbarrier complete_tx);
//   D = 128 lanes x n_mels columns of TMEM (FP32), zeroed with tcgen05.st, read back with tcgen05.ld by the four
//     epilogue warps and stored coalesced along t.
// Warp roles: 0-7 producers (0-3 also epilogue), 8 MMA issuer (one lane), 9 filterbank loader (one lane).
// HBM bound: 930 rows x 513 x 4 B in, 160 x 513 x 4 B out per clip.
// Measured alternatives (384 clips): this kernel (A in shared memory, 32-bin chunks, 2 stages, 2 CTAs/SM) 0.370 ms; one
// CTA/SM with 4 stages 0.584 ms; A operand written to TENSOR MEMORY by the producers (tcgen05.st, tcgen05.mma with
// [a_tmem]; 16-bin chunks, 3 stages, 256 TMEM columns, 2 CTAs/SM -- in the history: git show 49ad234:tools/k_melspec_tc_tmemA.cuh.txt) 0.467 ms.
// Time per chunk scales with the chunk's bytes in all of them: the limit is DRAM efficiency on 412-byte row pieces at a
// 2052-byte pitch (the banded FP32 kernel tops out at 1.6 TB/s on the same pattern), not the tensor pipe or the hand-over.
#pragma once
#include "cacfe_common.cuh"
#include "cacfe_async.cuh"  // mbarrier / bulk-copy wrappers

namespace cacfe {

struct MelTcChunk {
  int k0;     // first spectrogram bin of the chunk (32 bins)
  int n0;     // first mel band (multiple of 16) = first TMEM column
  int nc;     // number of bands (multiple of 16, <= kTcMaxN)
  int w_ofs;  // float offset of the chunk's packed weights: hi block [8][nc/8][8][4], then the lo block
};

constexpr int kTcM = 128;          // frames per tile (UMMA M)
constexpr int kTcK = 32;           // bins per chunk
constexpr int kTcMaxN = 64;        // bands per chunk the plan accepts
constexpr int kTcStages = 2;          // 2 x 48 KB: two CTAs per SM (measured: 0.37 ms / 384 clips; one CTA with 4 stages: 0.58 ms)
constexpr int kTcProducers = 256;
constexpr int kTcThreads = kTcProducers + 64;
constexpr int kTcSboA = 128;                       // bytes between 8-frame groups of A (core matrices are contiguous)
constexpr int kTcLboA = (kTcM / 8) * kTcSboA;      // bytes between K-adjacent core matrices (4 bins) of A = 2048
constexpr int kTcABytes = (kTcK / 4) * kTcLboA;    // one operand (hi or lo) of a stage = 16384
constexpr int kTcBBytes = kTcMaxN * kTcK * 4;      // hi or lo weights of a stage = 8192
constexpr int kTcStageBytes = 2 * kTcABytes + 2 * kTcBBytes;
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 256;

struct MelTcArgs {
  const float* spec;          // [B][n_bins][T]
  float* out;                 // [B][M][T][C] or [B][T][M]
  const float* wpk;           // packed weights, see MelTcChunk
  const MelTcChunk* chunks;
  int n_chunks, n_bins, T, n_mels, power, channels, layout, tiles_per_clip;
  int tile_frames;            // frames per tile (<= 128): T split evenly over the tiles (513 -> 5 x 103)
};

// ---- tcgen05 wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46),
  // version = 1 [46,48), base offset 0, layout type SWIZZLE_NONE = 0 [61,64)
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  // InstrDescriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), A and B K-major (bits 15, 16 = 0),
  // N >> 3 at [17,23), M >> 4 at [24,29)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__global__ void __launch_bounds__(kTcThreads, 2) melspec_tc_kernel(const MelTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
  uint64_t* bar_ready = bars;                    // [S] producers wrote A (256 arrivals)
  uint64_t* bar_wfull = bars + kTcStages;        // [S] weights landed (tx)
  uint64_t* bar_empty = bars + 2 * kTcStages;    // [S] MMAs of the stage retired (tcgen05.commit)
  uint64_t* bar_done = bars + 3 * kTcStages;     // accumulator complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 3 * kTcStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / a.tiles_per_clip;
  const int t0 = (blockIdx.x - b * a.tiles_per_clip) * a.tile_frames;
  const int t_end = min(t0 + a.tile_frames, a.T);
  const int tmem_cols = a.n_mels <= 32 ? 32 : a.n_mels <= 64 ? 64 : a.n_mels <= 128 ? 128 : 256;

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(smem_u32(&bar_ready[s]), kTcProducers);
      mbar_init(smem_u32(&bar_wfull[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM allocation: one warp, power-of-two columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // zero the accumulator: chunks write disjoint-but-overlapping column ranges, all with accumulate = 1
  if (warp < 4) {
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < a.n_mels; c += 16)
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + c),
          "r"(0)
          : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kTcProducers / 32) {
    // ---- producers: spectrogram chunk -> hi / lo operands in canonical layout ---------------------------------------
    const int t = tid & (kTcM - 1), kh = tid >> 7;       // frame inside the tile; this thread takes bins kh*16 .. +15
    const bool t_ok = t0 + t < t_end;
    const float* col = a.spec + (size_t)b * a.n_bins * a.T + t0 + t;
    const uint32_t a_off = (uint32_t)((t >> 3) * kTcSboA + (t & 7) * 16);
    auto load_chunk = [&](int c, float (&v)[16]) {
      const int k0 = c < a.n_chunks ? a.chunks[c].k0 + kh * 16 : a.n_bins;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = k0 + i;
        v[i] = (t_ok && k < a.n_bins) ? ld_stream(col + (size_t)k * a.T) : 0.0f;
      }
    };
    float v[16], vn[16], vnn[16];
    load_chunk(0, v);
    load_chunk(1, vn);
    for (int c = 0; c < a.n_chunks; ++c) {
      const int s = c % kTcStages, use = c / kTcStages;
      load_chunk(c + 2, vnn);  // two chunks of loads stay in flight while this one is split and stored
      if (use > 0) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)((use - 1) & 1));
      unsigned char* sa = stage_base + (size_t)s * kTcStageBytes;
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {  // four bins = one 16-byte row of a core matrix
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float x = v[4 * g4 + i];
          if (a.power == 2) x *= x;
          hi[i] = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
          lo[i] = x - hi[i];
        }
        const uint32_t off = (uint32_t)((kh * 4 + g4) * kTcLboA) + a_off;
        *reinterpret_cast<float4*>(sa + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(sa + kTcABytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores -> visible to the tensor core
      mbar_arrive(smem_u32(&bar_ready[s]));
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = vn[i];
        vn[i] = vnn[i];
      }
    }
  } else if (warp == 9) {
    // ---- filterbank loader ------------------------------------------------------------------------------------------
    if (lane == 0) {
      for (int c = 0; c < a.n_chunks; ++c) {
        const int s = c % kTcStages, use = c / kTcStages;
        if (use > 0) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)((use - 1) & 1));
        const MelTcChunk ch = a.chunks[c];
        const uint32_t bytes = (uint32_t)(2 * ch.nc * kTcK * 4);
        unsigned char* sb = stage_base + (size_t)s * kTcStageBytes + 2 * kTcABytes;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(smem_u32(&bar_wfull[s]), bytes);
        bulk_g2s(smem_u32(sb), a.wpk + ch.w_ofs, bytes, smem_u32(&bar_wfull[s]));
      }
    }
  } else if (warp == 8) {
    // ---- MMA issuer: one thread --------------------------------------------------------------------------------------
    if (lane == 0) {
      for (int c = 0; c < a.n_chunks; ++c) {
        const int s = c % kTcStages, use = c / kTcStages;
        const MelTcChunk ch = a.chunks[c];
        mbar_wait(smem_u32(&bar_wfull[s]), (uint32_t)(use & 1));
        mbar_wait(smem_u32(&bar_ready[s]), (uint32_t)(use & 1));
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_base + (size_t)s * kTcStageBytes);
        const uint32_t sb = sa + 2 * kTcABytes;
        const uint32_t lbo_b = (uint32_t)(ch.nc / 8) * 128u;     // between K-adjacent core matrices of B
        const uint32_t b_lo = (uint32_t)(ch.nc * kTcK * 4);      // the lo block follows the hi block
        const uint32_t idesc = umma_idesc_tf32(ch.nc);
        const uint32_t d = tmem + (uint32_t)ch.n0;
#pragma unroll
        for (int j = 0; j < kTcK / 8; ++j) {  // K = 8 per instruction: two K-adjacent core matrices of A and of B
          const uint64_t a_hi = umma_smem_desc(sa + j * 2 * kTcLboA, kTcLboA, kTcSboA);
          const uint64_t a_lo = umma_smem_desc(sa + kTcABytes + j * 2 * kTcLboA, kTcLboA, kTcSboA);
          const uint64_t b_hi = umma_smem_desc(sb + j * 2 * lbo_b, lbo_b, 128);
          const uint64_t b_lw = umma_smem_desc(sb + b_lo + j * 2 * lbo_b, lbo_b, 128);
          umma_tf32(d, a_hi, b_hi, idesc, true);
          umma_tf32(d, a_lo, b_hi, idesc, true);
          umma_tf32(d, a_hi, b_lw, idesc, true);
        }
        umma_commit(smem_u32(&bar_empty[s]));   // frees the stage when these MMAs have read it
      }
      umma_commit(smem_u32(bar_done));
    }
  }

  // ---- epilogue: TMEM -> registers -> global (warps 0-3: warp w owns TMEM lanes 32 w .. 32 w + 31) ---------------------
  if (warp < 4) {
    mbar_wait(smem_u32(bar_done), 0);
    tc_fence_after();
    const int t = t0 + 32 * warp + lane;
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < a.n_mels; c += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr + c)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (t < t_end) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int m = c + i;
          const float val = __uint_as_float(r[i]);
          if (a.layout == 1) {
            a.out[((size_t)b * a.T + t) * a.n_mels + m] = val;
          } else {
            float* o = a.out + (((size_t)b * a.n_mels + m) * a.T + t) * a.channels;
            for (int chn = 0; chn < a.channels; ++chn) o[chn] = val;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

}  // namespace cacfe
