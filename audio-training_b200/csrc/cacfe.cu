// cacfe C ABI (include/cacfe.h) over the sm_100a kernels in this directory.
// Host side: plan (tables, filterbank in band form), argument checks, launches.  No CPU compute path.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cacfe.h"
#include "k_compress.cuh"
#include "k_frontend.cuh"
#include "k_frontend_v3.cuh"
#include "k_melspec.cuh"
#include "k_normalize_cluster.cuh"
#include "k_melspec_stream.cuh"
#include "k_melspec_tc.cuh"
#include "k_pcen.cuh"
#include "k_pcen_bwd.cuh"
#include "k_sosfilt.cuh"
#include "k_signal.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess) return fail(CACFE_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));    \
  } while (0)

inline size_t align256(size_t n) { return (n + 255) & ~size_t(255); }

constexpr int kMaxSplits = 16;

std::mutex g_upload_gate[16];  // one per device: see cacfe_hostpipe_run

}  // namespace

struct cacfe_plan {
  cacfe_config cfg;
  int device = 0, sm_count = 0;
  int n_frames = 0, n_bins = 0, bin_lo = 0, bin_hi = -1, nnz = 0, origin = 0, nq = 0;
  std::vector<float> bank;  // host [n_mels][n_bins]
  float2* d_tw = nullptr;
  float* d_win = nullptr;
  float* d_band_w = nullptr;
  int* d_band_start = nullptr;
  int* d_band_ofs = nullptr;
  cacfe::K1Smem k1;
  cacfe::VSmem kv;
  cacfe::MelJobs jobs;
  float4* d_mel_w = nullptr;
  float4* d_tw4 = nullptr;
  float* d_win_full = nullptr;
  int nq_v3 = 0;     // 64-bin column groups of the 4096-point grid the (strided) bank reaches
  // tcgen05 stored-spectrogram path (CACFE_MEL_TC_3XTF32)
  bool tc_ok = false;
  std::vector<cacfe::MelTcChunk> tc_chunks;
  float* d_tc_w = nullptr;
  cacfe::MelTcChunk* d_tc_chunks = nullptr;
  int* d_mel_desc = nullptr;
  // streaming stored-spectrogram path (k_melspec_stream.cuh): per-row records and the 1 / 2 / 4-way band segmentations
  bool ms_ok = false;
  int ms_rows = 0;
  float4* d_ms_rows = nullptr;
  cacfe::MelStreamSeg ms_segs[3][cacfe::kMsMaxSegs];
  size_t smem_optin = 0, smem_per_sm = 0;
  bool v3_ok = false;
  bool force_generic = false;  // tests: run the non-streaming kernel on configurations that allow both
  bool no_hot = false;         // tests: the streaming kernel without its HOT instantiations (cacfe_plan_force_generic(plan, 2))
  bool frontend_ok = false;
  std::atomic<long long> launches{0};
  // optional CUDA-event bracket around every K1 launch (bench.py's live per-kernel timing)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> k1_events;
};

extern "C" {

int cacfe_version(void) { return CACFE_VERSION; }
const char* cacfe_last_error(void) { return g_err.c_str(); }

// custommel.py:6-54 restated in double precision with numpy's evaluation order:
//   linspace: start + i*step, last point pinned to `stop`;  rfftfreq: k * (1 / (n_fft * (1/sr))).
int cacfe_mel_filterbank(int sr, int n_mels, double fmin, double fmax, int n_fft, double break_freq, float* out) {
  if (sr <= 0 || n_mels <= 0 || n_fft <= 0 || break_freq <= 0 || out == nullptr)
    return fail(CACFE_EINVAL, "mel_filterbank: bad arguments");
  const int n_bins = 1 + n_fft / 2;
  const int n_pts = n_mels + 2;
  auto hz_to_mel = [&](double f) { return 2595.0 * std::log10(1.0 + f / break_freq); };
  const double lo = hz_to_mel(fmin), hi = hz_to_mel(fmax);
  std::vector<double> edge(n_pts);
  const double step = (hi - lo) / (double)(n_pts - 1);
  for (int i = 0; i < n_pts; ++i) {
    const double mel = (i == n_pts - 1) ? hi : (double)i * step + lo;
    edge[i] = break_freq * (std::pow(10.0, mel / 2595.0) - 1.0);
  }
  const double val = 1.0 / ((double)n_fft * (1.0 / (double)sr));
  for (int m = 0; m < n_mels; ++m) {
    const double d0 = edge[m + 1] - edge[m], d1 = edge[m + 2] - edge[m + 1];
    const double enorm = 2.0 / (edge[m + 2] - edge[m]);
    for (int k = 0; k < n_bins; ++k) {
      const double f = (double)k * val;
      const double lower = -(edge[m] - f) / d0;
      const double upper = (edge[m + 2] - f) / d1;
      const float w = (float)std::fmax(0.0, std::fmin(lower, upper));  // stored into the f32 array (custommel.py:37)
      out[(size_t)m * n_bins + k] = (float)((double)w * enorm);         // f32 row *= f64 enorm (custommel.py:42)
    }
  }
  return CACFE_OK;
}

int cacfe_num_frames(int n_samples, int n_fft, int hop, int framing) {
  if (hop <= 0 || n_fft <= 0 || n_samples < 0) return fail(CACFE_EINVAL, "num_frames: bad arguments");
  switch (framing) {
    case CACFE_FRAME_TF_PAD_END: return (n_samples + hop - 1) / hop;
    case CACFE_FRAME_CENTER_ZERO:
    case CACFE_FRAME_CENTER_REFLECT: return 1 + n_samples / hop;
    case CACFE_FRAME_NO_PAD: return n_samples < n_fft ? 0 : 1 + (n_samples - n_fft) / hop;
    default: return fail(CACFE_EINVAL, "num_frames: unknown framing %d", framing);
  }
}

int cacfe_plan_create(const cacfe_config* cfg, int device, cacfe_plan** out) {
  if (cfg == nullptr || out == nullptr) return fail(CACFE_EINVAL, "plan_create: null argument");
  *out = nullptr;
  if (cfg->n_fft < 8 || cfg->n_fft > (1 << 20) || (cfg->n_fft & 1))
    return fail(CACFE_EINVAL, "plan_create: n_fft=%d out of range", cfg->n_fft);
  if (cfg->hop < 1 || cfg->hop > 2048) return fail(CACFE_EINVAL, "plan_create: hop=%d out of range [1, 2048]", cfg->hop);
  if (cfg->n_samples < 1) return fail(CACFE_EINVAL, "plan_create: n_samples must be positive");
  if (cfg->power != 1 && cfg->power != 2) return fail(CACFE_EINVAL, "plan_create: power must be 1 or 2");
  if (cfg->n_mels < 1 || cfg->n_mels > 1024) return fail(CACFE_EINVAL, "plan_create: n_mels out of range");
  if (cfg->channels < 1 || cfg->channels > 16) return fail(CACFE_EINVAL, "plan_create: channels out of range");
  if (cfg->out_layout != CACFE_LAYOUT_BMTC && cfg->out_layout != CACFE_LAYOUT_BTM)
    return fail(CACFE_EINVAL, "plan_create: unknown out_layout");
  if (cfg->out_layout == CACFE_LAYOUT_BTM && cfg->channels != 1)
    return fail(CACFE_EINVAL, "plan_create: layout BTM has no channel axis (channels must be 1)");
  if (cfg->framing < 0 || cfg->framing > 3) return fail(CACFE_EINVAL, "plan_create: unknown framing");
  if (cfg->framing == CACFE_FRAME_CENTER_REFLECT && cfg->n_samples <= cfg->n_fft / 2)
    return fail(CACFE_EINVAL, "plan_create: reflect padding needs n_samples > n_fft/2");
  if (cfg->mel_impl != CACFE_MEL_BANDED_FP32 && cfg->mel_impl != CACFE_MEL_TC_3XTF32)
    return fail(CACFE_EINVAL, "plan_create: unknown mel_impl %d", cfg->mel_impl);

  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(CACFE_ECUDA, "plan_create: no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= count) return fail(CACFE_EDEVICE, "plan_create: device %d of %d", device, count);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(CACFE_EDEVICE, "plan_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));

  cacfe_plan* p = new cacfe_plan();
  p->cfg = *cfg;
  p->cfg.filterbank = nullptr;
  p->device = device;
  p->sm_count = prop.multiProcessorCount;
  p->n_bins = 1 + cfg->n_fft / 2;
  p->n_frames = cacfe_num_frames(cfg->n_samples, cfg->n_fft, cfg->hop, cfg->framing);
  p->origin = (cfg->framing == CACFE_FRAME_CENTER_ZERO || cfg->framing == CACFE_FRAME_CENTER_REFLECT) ? -cfg->n_fft / 2 : 0;
  if (p->n_frames < 1) {
    delete p;
    return fail(CACFE_ESHAPE, "plan_create: clip too short for one frame");
  }
  p->bank.resize((size_t)cfg->n_mels * p->n_bins);
  if (cfg->filterbank != nullptr) {
    std::memcpy(p->bank.data(), cfg->filterbank, p->bank.size() * sizeof(float));
  } else {
    int rc = cacfe_mel_filterbank(cfg->sr, cfg->n_mels, cfg->fmin, cfg->fmax, cfg->n_fft, cfg->break_freq, p->bank.data());
    if (rc != CACFE_OK) {
      delete p;
      return rc;
    }
  }
  // band form: per mel the run [first non-zero, last non-zero]
  std::vector<float> bw;
  std::vector<int> bfirst(cfg->n_mels, 0), bofs(cfg->n_mels + 1, 0);
  int lo = p->n_bins, hi = -1;
  for (int m = 0; m < cfg->n_mels; ++m) {
    const float* row = p->bank.data() + (size_t)m * p->n_bins;
    int a = -1, b = -1;
    for (int k = 0; k < p->n_bins; ++k)
      if (row[k] != 0.0f) {
        if (a < 0) a = k;
        b = k;
      }
    bofs[m] = (int)bw.size();
    if (a >= 0) {
      bfirst[m] = a;
      for (int k = a; k <= b; ++k) bw.push_back(row[k]);
      lo = a < lo ? a : lo;
      hi = b > hi ? b : hi;
    }
  }
  bofs[cfg->n_mels] = (int)bw.size();
  if (hi < 0) {  // an all-zero bank is legal (every feature is 0): keep one dummy bin
    lo = hi = 0;
  }
  for (int m = 0; m < cfg->n_mels; ++m) bfirst[m] = (bofs[m + 1] > bofs[m]) ? bfirst[m] - lo : 0;
  p->bin_lo = lo;
  p->bin_hi = hi;
  p->nnz = (int)bw.size();
  p->nq = hi / 64 + 1;
  if (bw.empty()) bw.push_back(0.0f);

  // tensor-core form for the stored-spectrogram path: 32-bin chunks, each with the 16-aligned run of bands it touches,
  // weights split hi (TF32: low 13 mantissa bits cleared) + lo and packed in the UMMA canonical K-major layout
  // [k / 4][n / 8][n % 8][k % 4] (k_melspec_tc.cuh)
  std::vector<float> tc_w;
  p->tc_ok = cfg->n_mels % 16 == 0 && cfg->n_mels <= 256;
  for (int k0 = lo & ~3; p->tc_ok && k0 <= hi; k0 += cacfe::kTcK) {
    int m_lo = cfg->n_mels, m_hi = -1;
    for (int m = 0; m < cfg->n_mels; ++m)
      for (int k = k0; k < k0 + cacfe::kTcK && k < p->n_bins; ++k)
        if (p->bank[(size_t)m * p->n_bins + k] != 0.0f) {
          m_lo = std::min(m_lo, m);
          m_hi = std::max(m_hi, m);
        }
    if (m_hi < 0) continue;  // no band touches these bins
    cacfe::MelTcChunk ch;
    ch.k0 = k0;
    ch.n0 = m_lo & ~15;
    ch.nc = ((m_hi + 16) & ~15) - ch.n0;
    ch.w_ofs = (int)tc_w.size();
    if (ch.nc > cacfe::kTcMaxN) {
      p->tc_ok = false;
      break;
    }
    const size_t block = (size_t)ch.nc * cacfe::kTcK;
    tc_w.resize(tc_w.size() + 2 * block, 0.0f);
    for (int n = 0; n < ch.nc; ++n)
      for (int kk = 0; kk < cacfe::kTcK; ++kk) {
        const int k = k0 + kk, m = ch.n0 + n;
        const float w = (k < p->n_bins && m < cfg->n_mels) ? p->bank[(size_t)m * p->n_bins + k] : 0.0f;
        uint32_t bits;
        std::memcpy(&bits, &w, 4);
        bits &= 0xffffe000u;
        float w_hi;
        std::memcpy(&w_hi, &bits, 4);
        const size_t idx = (size_t)(kk / 4) * (ch.nc / 8) * 32 + (size_t)(n / 8) * 32 + (n % 8) * 4 + (kk % 4);
        tc_w[ch.w_ofs + idx] = w_hi;
        tc_w[ch.w_ofs + block + idx] = w - w_hi;
      }
    p->tc_chunks.push_back(ch);
  }
  if (p->tc_chunks.empty()) p->tc_ok = false;
  if (cfg->mel_impl == CACFE_MEL_TC_3XTF32 && !p->tc_ok) {
    delete p;
    return fail(CACFE_EINVAL, "plan_create: the tensor-core mel path needs n_mels %% 16 == 0, n_mels <= 256 and at most %d "
                "bands per 32-bin chunk", cacfe::kTcMaxN);
  }

  // streaming form of the stored-spectrogram path: every bin feeds at most two adjacent bands (lo, lo + 1), lo never
  // decreases along k.  Anything else keeps the per-column kernel.
  std::vector<float4> ms_rows;
  p->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  p->smem_per_sm = (size_t)prop.sharedMemPerMultiprocessor;
  p->ms_ok = hi >= lo && p->nnz > 0;
  {
    int cur = -1;
    for (int k = lo; k <= hi && p->ms_ok; ++k) {
      int m_a = -1, m_b = -1, cnt = 0;
      for (int m = 0; m < cfg->n_mels; ++m)
        if (p->bank[(size_t)m * p->n_bins + k] != 0.0f) {
          if (cnt == 0) m_a = m;
          m_b = m;
          ++cnt;
        }
      int L;
      if (cnt == 0) {
        L = cur < 0 ? 0 : cur;
      } else if (cnt == 1) {
        L = std::max(std::max(cur, m_a - 1), 0);
        if (m_a < L) p->ms_ok = false;
      } else {
        L = m_a;
        if (cnt > 2 || m_b != m_a + 1 || L < cur) p->ms_ok = false;
      }
      if (!p->ms_ok) break;
      cur = L;
      float4 r;
      r.x = p->bank[(size_t)L * p->n_bins + k];
      r.y = L + 1 < cfg->n_mels ? p->bank[(size_t)(L + 1) * p->n_bins + k] : 0.0f;
      std::memcpy(&r.z, &L, 4);
      r.w = 0.0f;
      ms_rows.push_back(r);
    }
    p->ms_rows = (int)ms_rows.size();
  }
  if (p->ms_ok) {
    // first / last row of every band (absolute bins); cuts at the band whose first row passes the i-th share of the rows
    auto first_row = [&](int m) { return bfirst[m] + lo; };
    auto n_rows_of = [&](int m) { return bofs[m + 1] - bofs[m]; };
    for (int v = 0; v < 3; ++v) {
      const int ns = 1 << v;
      int cut[cacfe::kMsMaxSegs + 1];
      cut[0] = 0;
      cut[ns] = cfg->n_mels;
      for (int i = 1; i < ns; ++i) {
        const int target = lo + (int)((long long)(hi - lo + 1) * i / ns);
        int m = cut[i - 1];
        while (m < cfg->n_mels && (n_rows_of(m) == 0 || first_row(m) < target)) ++m;
        cut[i] = std::max(m, cut[i - 1]);
      }
      for (int i = 0; i < ns; ++i) {
        cacfe::MelStreamSeg sg;
        sg.m0 = cut[i];
        sg.m1 = cut[i + 1];
        int r0 = p->n_bins, r1 = 0;
        for (int m = sg.m0; m < sg.m1; ++m)
          if (n_rows_of(m) > 0) {
            r0 = std::min(r0, first_row(m));
            r1 = std::max(r1, first_row(m) + n_rows_of(m));
          }
        if (r1 <= r0) r0 = r1 = lo;
        sg.row0 = r0;
        sg.row1 = r1;
        p->ms_segs[v][i] = sg;
      }
    }
  }

  // The fused raw->mel kernel exists for n_fft = 4096 (the reference's only shipped configuration); other sizes
  // get a plan for the spectrogram / PCEN / compression entry points and cacfe_frontend refuses them.
  p->k1 = cacfe::k1_smem_layout(cfg->hop, cfg->n_mels, p->nnz, 1);
  if (p->k1.total > (size_t)prop.sharedMemPerBlockOptin) p->k1 = cacfe::k1_smem_layout(cfg->hop, cfg->n_mels, p->nnz, 0);
  p->frontend_ok = cfg->n_fft == cacfe::kFft && p->k1.total <= (size_t)prop.sharedMemPerBlockOptin;

  // persistent kernel: 4096-point transform; shorter n_fft = 4096 / r use bins r k of it (bank placed with stride r)
  const bool fft_divides = cfg->n_fft >= 512 && cfg->n_fft <= cacfe::kFft && cacfe::kFft % cfg->n_fft == 0;
  const int ratio = fft_divides ? cacfe::kFft / cfg->n_fft : 1;
  std::vector<float> bank4096;
  const float* bank_v3 = p->bank.data();
  if (fft_divides && ratio > 1) {
    bank4096.assign((size_t)cfg->n_mels * (cacfe::kFft / 2 + 1), 0.0f);
    for (int m = 0; m < cfg->n_mels; ++m)
      for (int k = 0; k < p->n_bins; ++k) bank4096[(size_t)m * (cacfe::kFft / 2 + 1) + (size_t)ratio * k] = p->bank[(size_t)m * p->n_bins + k];
    bank_v3 = bank4096.data();
  }
  p->nq_v3 = (ratio * hi) / 64 + 1;
  p->jobs = cacfe::build_mel_jobs(bank_v3, cfg->n_mels, cacfe::kFft / 2 + 1, 32 * (p->nq_v3 <= 15 ? 15 : 33));
  p->kv = cacfe::v3_smem_layout(cfg->hop, p->jobs.total_quads);
  bool reflect_fits = true;
  if (cfg->framing == CACFE_FRAME_CENTER_REFLECT)
    for (int t0 = 0; t0 < p->n_frames; t0 += cacfe::kVTileFrames) {
      const long long s_lo = (long long)p->origin + (long long)cfg->hop * t0;
      const int t1 = std::min(p->n_frames, t0 + cacfe::kVTileFrames);
      const long long p_max = s_lo + (long long)cfg->hop * (t1 - 1 - t0) + cfg->n_fft - 1;  // last sample a frame reads
      if (p_max >= cfg->n_samples && 2LL * (cfg->n_samples - 1) - p_max < std::max(s_lo, 0LL)) reflect_fits = false;
      if (s_lo < 0 && -s_lo >= s_lo + p->kv.tile_len) reflect_fits = false;
    }
  p->v3_ok = fft_divides && p->jobs.ok && p->kv.total <= (size_t)prop.sharedMemPerBlockOptin &&
             p->kv.tile_pad <= cacfe::kNormIters * cacfe::kVThreads * 4 &&
             cfg->n_samples % 4 == 0 && reflect_fits;

  // tables
  std::vector<float2> tw(4096);
  for (int k1 = 0; k1 < 64; ++k1)
    for (int n2 = 0; n2 < 64; ++n2) {
      const double ang = -2.0 * M_PI * (double)((k1 * n2) % 4096) / 4096.0;
      tw[k1 * 64 + n2] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  std::vector<float> tw4(32 * 64 * 4);  // persistent kernel: twiddles of output pair (k, k+1), k even, for thread n2
  for (int pr = 0; pr < 32; ++pr)
    for (int n2 = 0; n2 < 64; ++n2)
      for (int h = 0; h < 2; ++h) {
        const double ang = -2.0 * M_PI * (double)(((2 * pr + h) * n2) % 4096) / 4096.0;
        tw4[(pr * 64 + n2) * 4 + h] = (float)std::cos(ang);
        tw4[(pr * 64 + n2) * 4 + 2 + h] = (float)std::sin(ang);
      }
  std::vector<float> win_full(cacfe::kFft, 0.0f);   // periodic Hann of length n_fft (tf.signal.hann_window / scipy fftbins)
  if (fft_divides)
    for (int n = 0; n < cfg->n_fft; ++n) win_full[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)cfg->n_fft));
  std::vector<float> win(2049);
  for (int n = 0; n <= 2048; ++n) win[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / 4096.0));

  auto upload = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = upload((void**)&p->d_tw, tw.data(), tw.size() * sizeof(float2));
  if (e == cudaSuccess) e = upload((void**)&p->d_win, win.data(), win.size() * sizeof(float));
  if (e == cudaSuccess) e = upload((void**)&p->d_band_w, bw.data(), bw.size() * sizeof(float));
  if (e == cudaSuccess) e = upload((void**)&p->d_band_start, bfirst.data(), bfirst.size() * sizeof(int));
  if (e == cudaSuccess) e = upload((void**)&p->d_band_ofs, bofs.data(), bofs.size() * sizeof(int));
  if (e == cudaSuccess && p->frontend_ok)
    e = cudaFuncSetAttribute(cacfe::stft_mel_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
  if (e == cudaSuccess && p->frontend_ok)
    e = cudaFuncSetAttribute(cacfe::stft_mel_kernel<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
  if (e == cudaSuccess) {
    const int scan_smem = cacfe::kScanWarps * cacfe::kScanMaxRow * (int)sizeof(float);
    e = cudaFuncSetAttribute(cacfe::pcen_scan_kernel<cacfe::PCEN_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, scan_smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(cacfe::pcen_scan_kernel<cacfe::PCEN_REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, scan_smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(cacfe::pcen_scan_kernel<cacfe::PCEN_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, scan_smem);
  }
  if (e == cudaSuccess && p->tc_ok) e = upload((void**)&p->d_tc_w, tc_w.data(), tc_w.size() * sizeof(float));
  if (e == cudaSuccess && p->tc_ok)
    e = upload((void**)&p->d_tc_chunks, p->tc_chunks.data(), p->tc_chunks.size() * sizeof(cacfe::MelTcChunk));
  if (e == cudaSuccess && p->tc_ok)
    e = cudaFuncSetAttribute(cacfe::melspec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cacfe::kTcSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(cacfe::pcen_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             cacfe::kBwdMaxSeg * cacfe::kBwdMaxThreads * (int)sizeof(float));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(cacfe::row_normalize_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cacfe::kNcMaxBytesPerCta);
  if (e == cudaSuccess && p->ms_ok) e = upload((void**)&p->d_ms_rows, ms_rows.data(), ms_rows.size() * sizeof(float4));
  if (e == cudaSuccess && p->ms_ok) {
    const void* kernels[24] = {
#define CACFE_MS_K(C_) (const void*)cacfe::melspec_stream_kernel<C_, false, 1>, (const void*)cacfe::melspec_stream_kernel<C_, false, 3>, \
                       (const void*)cacfe::melspec_stream_kernel<C_, false, 0>, (const void*)cacfe::melspec_stream_kernel<C_, true, 1>,  \
                       (const void*)cacfe::melspec_stream_kernel<C_, true, 3>,  (const void*)cacfe::melspec_stream_kernel<C_, true, 0>
        CACFE_MS_K(1), CACFE_MS_K(2), CACFE_MS_K(3), CACFE_MS_K(4)
#undef CACFE_MS_K
    };
    for (int q = 0; q < 24 && e == cudaSuccess; ++q)
      e = cudaFuncSetAttribute(kernels[q], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
  }
  if (e == cudaSuccess && p->v3_ok) e = upload((void**)&p->d_mel_w, p->jobs.w.data(), p->jobs.w.size() * sizeof(float));
  if (e == cudaSuccess && p->v3_ok) e = upload((void**)&p->d_win_full, win_full.data(), win_full.size() * sizeof(float));
  if (e == cudaSuccess && p->v3_ok) e = upload((void**)&p->d_tw4, tw4.data(), tw4.size() * sizeof(float));
  if (e == cudaSuccess && p->v3_ok) e = upload((void**)&p->d_mel_desc, p->jobs.desc.data(), p->jobs.desc.size() * sizeof(int));
  if (e == cudaSuccess && p->v3_ok) {
    const void* kernels[13] = {
#define CACFE_V3_K(NQ_, LAYOUT_) (const void*)cacfe::stft_mel_v3_kernel<NQ_, LAYOUT_, false>, (const void*)cacfe::stft_mel_v3_kernel<NQ_, LAYOUT_, true>
        CACFE_V3_K(15, cacfe::LAYOUT_BTM), CACFE_V3_K(15, cacfe::LAYOUT_BMTC),
        CACFE_V3_K(33, cacfe::LAYOUT_BTM), CACFE_V3_K(33, cacfe::LAYOUT_BMTC),
        (const void*)cacfe::stft_mel_v3_kernel<15, cacfe::LAYOUT_BTM, true, 1>,
        (const void*)cacfe::stft_mel_v3_kernel<15, cacfe::LAYOUT_BTM, true, 2>,
        CACFE_V3_K(33, cacfe::LAYOUT_SPECT), (const void*)cacfe::stft_mel_v3_kernel<33, cacfe::LAYOUT_SPECT, true, 1>
#undef CACFE_V3_K
    };
    for (int q = 0; q < 13 && e == cudaSuccess; ++q)
      // the attribute belongs to the function, not to the plan: always the device maximum, so that a plan created later
      // with a smaller layout cannot shrink it under an earlier plan
      e = cudaFuncSetAttribute(kernels[q], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
  }
  if (e != cudaSuccess) {
    cacfe_plan_destroy(p);
    return fail(CACFE_ECUDA, "plan_create: %s", cudaGetErrorString(e));
  }
  *out = p;
  return CACFE_OK;
}

void cacfe_plan_destroy(cacfe_plan* p) {
  if (p == nullptr) return;
  cudaSetDevice(p->device);
  cudaFree(p->d_tw);
  cudaFree(p->d_win);
  cudaFree(p->d_band_w);
  cudaFree(p->d_band_start);
  cudaFree(p->d_band_ofs);
  cudaFree(p->d_mel_w);
  cudaFree(p->d_tc_w);
  cudaFree(p->d_tc_chunks);
  cudaFree(p->d_tw4);
  cudaFree(p->d_win_full);
  cudaFree(p->d_mel_desc);
  cudaFree(p->d_ms_rows);
  delete p;
}

int cacfe_plan_num_frames(const cacfe_plan* p) { return p ? p->n_frames : fail(CACFE_EINVAL, "null plan"); }
int cacfe_plan_num_bins(const cacfe_plan* p) { return p ? p->n_bins : fail(CACFE_EINVAL, "null plan"); }
long long cacfe_plan_launch_count(const cacfe_plan* p) { return p ? p->launches.load() : 0; }

int cacfe_plan_profile(cacfe_plan* p, int enable) {
  if (!p) return fail(CACFE_EINVAL, "plan_profile: null plan");
  p->profile = enable != 0;
  return CACFE_OK;
}

int cacfe_plan_profile_read(cacfe_plan* p, double* k1_ms, long long* k1_launches) {
  if (!p || !k1_ms || !k1_launches) return fail(CACFE_EINVAL, "plan_profile_read: null argument");
  double total = 0.0;
  for (auto& ev : p->k1_events) {
    float ms = 0.0f;
    CUDA_TRY(cudaEventSynchronize(ev.second));
    CUDA_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
    total += ms;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  *k1_ms = total;
  *k1_launches = (long long)p->k1_events.size();
  p->k1_events.clear();
  return CACFE_OK;
}

int cacfe_plan_force_generic(cacfe_plan* p, int enable) {
  if (!p) return fail(CACFE_EINVAL, "plan_force_generic: null plan");
  p->force_generic = enable == 1;
  p->no_hot = enable == 2;
  return CACFE_OK;
}

int cacfe_plan_filterbank(const cacfe_plan* p, float* out) {
  if (!p || !out) return fail(CACFE_EINVAL, "plan_filterbank: null argument");
  std::memcpy(out, p->bank.data(), p->bank.size() * sizeof(float));
  return CACFE_OK;
}

int cacfe_plan_bin_range(const cacfe_plan* p, int* lo, int* hi) {
  if (!p || !lo || !hi) return fail(CACFE_EINVAL, "plan_bin_range: null argument");
  *lo = p->bin_lo;
  *hi = p->bin_hi;
  return CACFE_OK;
}

}  // extern "C"

namespace {

int pick_splits(const cacfe_plan* p, long long rows) {
  long long s = (2LL * p->sm_count + rows - 1) / rows;
  if (s < 1) s = 1;
  if (s > kMaxSplits) s = kMaxSplits;
  return (int)s;
}

struct PcenGrid {
  int block, gx;
};
PcenGrid pcen_grid(long long rows_per_clip) {
  PcenGrid g;
  long long blk = ((rows_per_clip + 31) / 32) * 32;
  g.block = (int)(blk > 256 ? 256 : blk);
  g.gx = (int)((rows_per_clip + g.block - 1) / g.block);
  return g;
}

// One pass of the lane-per-row PCEN kernel.  Root 2 (the layer's initial value, tfpcen.py:78-87) takes the instantiation with
// the root at compile time; cacfe_plan_force_generic(plan, 2) turns that off (tests: the two must agree bit for bit).
template <int MODE>
void pcen_pass(const cacfe_plan* p, dim3 grid, int block, int dyn, cudaStream_t st, const cacfe::PcenArgs& a) {
  if (a.root_is_2 && !(p && p->no_hot)) cacfe::pcen_kernel<MODE, true><<<grid, block, dyn, st>>>(a);
  else cacfe::pcen_kernel<MODE><<<grid, block, dyn, st>>>(a);
}

size_t pcen_ws_bytes(int B, long long rows_per_clip) {
  const PcenGrid g = pcen_grid(rows_per_clip);
  const size_t lanes = align256((size_t)B * g.gx * sizeof(float2));
  // the warp-scan kernel writes one partial per 8 rows; rows <= rows_per_clip per clip whatever the inner size is
  const size_t scan = align256((size_t)B * ((rows_per_clip + cacfe::kScanWarps - 1) / cacfe::kScanWarps) * sizeof(float2));
  return (lanes > scan ? lanes : scan) + align256((size_t)B * sizeof(float2));
}

int compress_blocks(long long entries, long long per_entry) {
  long long want = (per_entry + 4095) / 4096;
  long long cap = 4096 / (entries < 1 ? 1 : entries);
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

size_t frontend_ws_bytes(int B) {
  return align256((size_t)B * kMaxSplits * sizeof(float2)) + align256((size_t)B * sizeof(float2));
}

int check_launch(cacfe_plan* p, const char* what, int n_launches = 1) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CACFE_ECUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  p->launches.fetch_add(n_launches);
  return CACFE_OK;
}

int fill_pcen_args(const cacfe_pcen_params* q, cacfe::PcenArgs& a) {
  if (!(q->eps > 0.0f)) return fail(CACFE_EINVAL, "pcen: eps must be positive");
  const float w = fminf(fmaxf(q->smooth, 0.0f), 1.0f);        // tf.clip_by_value(smooth, 0, 1)   tfpcen.py:35
  const float gain = fminf(q->gain, 1.0f);                     // tfpcen.py:90
  const float root = fmaxf(q->root, 1.0f);                     // tfpcen.py:91
  a.zero = 0;
  a.w = w;
  a.one_minus_w = 1.0f - w;
  a.gain = gain;
  a.bias = q->bias;
  a.inv_root = 1.0f / root;
  a.bias_pow = powf(q->bias, a.inv_root);
  a.eps = q->eps;
  a.root_is_2 = (root == 2.0f);
  return CACFE_OK;
}

}  // namespace

extern "C" {

size_t cacfe_pcen_workspace_bytes(int B, long long outer_per_clip, int inner) {
  return pcen_ws_bytes(B, outer_per_clip * inner);
}

size_t cacfe_compress_workspace_bytes(long long entries, long long per_entry) {
  return align256((size_t)entries * compress_blocks(entries, per_entry) * sizeof(cacfe::Stats)) +
         align256((size_t)entries * sizeof(cacfe::Stats));
}

size_t cacfe_workspace_bytes(const cacfe_plan* p, int B) {
  if (!p || B < 1) return 0;
  const size_t feat = (size_t)p->n_frames * p->cfg.n_mels;
  size_t need = frontend_ws_bytes(B);
  need += align256((size_t)B * feat * sizeof(float));                   // mel intermediate of frontend_pcen
  size_t pc = pcen_ws_bytes(B, p->cfg.n_mels);
  const size_t pc_img = pcen_ws_bytes(B, (long long)p->cfg.n_mels * p->cfg.channels);
  if (pc_img > pc) pc = pc_img;
  size_t cp = cacfe_compress_workspace_bytes(B, (long long)feat * p->cfg.channels);
  const size_t cp1 = cacfe_compress_workspace_bytes(1, (long long)B * feat * p->cfg.channels);
  if (cp1 > cp) cp = cp1;
  return need + (pc > cp ? pc : cp);
}

int cacfe_normalize(cacfe_plan* p, const float* in, float* out, long long rows, long long n, void* ws, void* stream) {
  if (!p || !in || !out || !ws) return fail(CACFE_EINVAL, "normalize: null argument");
  if (rows < 1 || n < 1 || rows > 65535) return fail(CACFE_ESHAPE, "normalize: rows=%lld n=%lld", rows, n);
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  // a cluster of 8 CTAs holds the clip in (distributed) shared memory between the min/max and the rescale: one HBM read
  const long long per_cta = n / cacfe::kNcCluster;
  if (!p->force_generic && n % (4 * cacfe::kNcCluster) == 0 && per_cta * 4 <= cacfe::kNcMaxBytesPerCta &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 && rows * cacfe::kNcCluster <= 2147483647LL) {
    cacfe::row_normalize_cluster_kernel<<<(unsigned)(rows * cacfe::kNcCluster), cacfe::kNcThreads, (size_t)per_cta * 4, st>>>(
        in, out, n, (int)per_cta);
    return check_launch(p, "normalize (cluster)", 1);
  }
  int splits = pick_splits(p, rows);
  if (rows * splits * (long long)sizeof(float2) > (long long)frontend_ws_bytes((int)rows)) splits = 1;
  float2* partial = (float2*)ws;
  cacfe::row_minmax_kernel<<<dim3(splits, (unsigned)rows), 256, 0, st>>>(in, n, splits, partial);
  long long gx = (n + 256 * 8 - 1) / (256 * 8);
  if (gx > 1024) gx = 1024;
  cacfe::row_normalize_kernel<<<dim3((unsigned)gx, (unsigned)rows), 256, 0, st>>>(in, out, n, splits, partial);
  return check_launch(p, "normalize", 2);
}

static int launch_frontend(cacfe_plan* p, const float* raw, float* feat, int B, int layout, int channels, void* ws,
                           cudaStream_t st, float* staging = nullptr) {
  if (!p->frontend_ok && !p->v3_ok)
    return fail(CACFE_EINVAL, "frontend: n_fft=%d hop=%d n_mels=%d is not served by the fused kernels: they take n_fft 4096 "
                "(512..2048 through the persistent kernel, which needs n_samples %% 4 == 0 and a bank of at most 192 bands) and a "
                "frame tile of 4096 + 11 hop samples that fits shared memory next to the tables (hop <= ~430 at 192 bands)",
                p->cfg.n_fft, p->cfg.hop, p->cfg.n_mels);
  cacfe::FrontendArgs a;
  int launches = 1;
  a.norm = nullptr;
  if (p->cfg.normalize) {
    const int splits = pick_splits(p, B);
    float2* partial = (float2*)ws;
    float2* norm = (float2*)((char*)ws + align256((size_t)B * kMaxSplits * sizeof(float2)));
    cacfe::row_minmax_kernel<<<dim3(splits, B), 256, 0, st>>>(raw, p->cfg.n_samples, splits, partial);
    cacfe::minmax_finalize_kernel<true><<<B, 32, 0, st>>>(partial, splits, norm);  // -> (max - min, min) per clip
    a.norm = norm;
    launches = 3;
  }
  a.in = raw;
  a.out = feat;
  a.tw = p->d_tw;
  a.win = p->d_win;
  a.band_w = p->d_band_w;
  a.band_start = p->d_band_start;
  a.band_ofs = p->d_band_ofs;
  a.n_samples = p->cfg.n_samples;
  a.hop = p->cfg.hop;
  a.n_frames = p->n_frames;
  a.n_mels = p->cfg.n_mels;
  a.nnz = p->nnz;
  a.origin = p->origin;
  a.reflect = p->cfg.framing == CACFE_FRAME_CENTER_REFLECT;
  a.power = p->cfg.power;
  a.channels = channels;
  a.layout = layout;
  a.bin_lo = p->bin_lo;
  a.bin_hi = p->bin_hi;
  a.tiles_per_clip = (p->n_frames + cacfe::kTileFrames - 1) / cacfe::kTileFrames;
  const long long grid = (long long)B * a.tiles_per_clip;
  if (grid > 2147483647LL) return fail(CACFE_ESHAPE, "frontend: batch too large for one launch");
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (p->profile) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
  }
  const bool aligned = (reinterpret_cast<uintptr_t>(raw) & 15) == 0;
  const bool v3 = p->v3_ok && !p->force_generic && aligned &&
                  (reinterpret_cast<uintptr_t>(a.norm) & 15) == 0;  // the kernel bulk-copies 16-byte pairs of it
  if (v3) {
    cacfe::MelArgs mj;
    mj.w = p->d_mel_w;
    mj.tw4 = p->d_tw4;
    mj.win = p->d_win_full;
    mj.desc = p->d_mel_desc;
    for (int sgm = 0; sgm < cacfe::kMelMaxSeg; ++sgm) mj.nq[sgm] = p->jobs.nq[sgm];
    mj.split_seg = p->jobs.split_seg;
    mj.total_quads = p->jobs.total_quads;
    mj.tile_len = p->kv.tile_len;
    mj.tile_pad = p->kv.tile_pad;
    a.bw_in_smem = 0;
    a.tiles_per_clip = (p->n_frames + cacfe::kVTileFrames - 1) / cacfe::kVTileFrames;
    const long long tiles = (long long)B * a.tiles_per_clip;
    const unsigned ctas = (unsigned)(tiles < p->sm_count ? tiles : p->sm_count);  // one persistent CTA per SM
    // the image layout [B][M][T][C] is written through a [B][T][M] staging buffer when the caller's workspace has one:
    // the FFT kernel stores 640-byte rows instead of 8 * C-byte pieces, and a tiled transpose does the rest
    // (raw_to_mel, B = 2048: 6.35 ms direct)
    const bool via_staging = layout == CACFE_LAYOUT_BMTC && staging != nullptr;
    if (via_staging) {
      a.out = staging;
      a.layout = cacfe::LAYOUT_BTM;
      a.channels = 1;
    }
    const bool btm = layout == CACFE_LAYOUT_BTM || via_staging;
    mj.spec_ratio = cacfe::kFft / p->cfg.n_fft;
    mj.spec_bins = p->n_bins;
    // n_fft = 4096: Hann window computed per thread (WINC); shorter transforms read the zero-padded window table
    const bool winc = p->cfg.n_fft == cacfe::kFft;
#define CACFE_V3_LAUNCH(NQ_, LAYOUT_)                                                                                             \
  do {                                                                                                                            \
    if (winc) cacfe::stft_mel_v3_kernel<NQ_, LAYOUT_, true><<<ctas, cacfe::kVThreads, p->kv.total, st>>>(a, mj, (int)tiles);    \
    else cacfe::stft_mel_v3_kernel<NQ_, LAYOUT_, false><<<ctas, cacfe::kVThreads, p->kv.total, st>>>(a, mj, (int)tiles);        \
  } while (0)
    const bool hot = !p->no_hot;
    if (hot && layout == cacfe::LAYOUT_SPECT && winc && a.norm != nullptr && !a.reflect && a.power == 1)   // audiodataset.load_data's configuration
      cacfe::stft_mel_v3_kernel<33, cacfe::LAYOUT_SPECT, true, 1><<<ctas, cacfe::kVThreads, p->kv.total, st>>>(a, mj, (int)tiles);
    else if (layout == cacfe::LAYOUT_SPECT)
      CACFE_V3_LAUNCH(33, cacfe::LAYOUT_SPECT);
    else if (hot && p->nq_v3 <= 15 && btm && winc && a.norm != nullptr && !a.reflect && a.power == 2)   // the benchmarked configuration
      cacfe::stft_mel_v3_kernel<15, cacfe::LAYOUT_BTM, true, 1><<<ctas, cacfe::kVThreads, p->kv.total, st>>>(a, mj, (int)tiles);
    else if (hot && p->nq_v3 <= 15 && btm && winc && a.norm == nullptr && !a.reflect && a.power == 2)   // raw_to_mel / get_spect on normalised clips
      cacfe::stft_mel_v3_kernel<15, cacfe::LAYOUT_BTM, true, 2><<<ctas, cacfe::kVThreads, p->kv.total, st>>>(a, mj, (int)tiles);
    else if (p->nq_v3 <= 15 && btm)
      CACFE_V3_LAUNCH(15, cacfe::LAYOUT_BTM);
    else if (p->nq_v3 <= 15)
      CACFE_V3_LAUNCH(15, cacfe::LAYOUT_BMTC);
    else if (btm)
      CACFE_V3_LAUNCH(33, cacfe::LAYOUT_BTM);
    else
      CACFE_V3_LAUNCH(33, cacfe::LAYOUT_BMTC);
#undef CACFE_V3_LAUNCH
  } else {
    if (layout == cacfe::LAYOUT_SPECT)
      return fail(CACFE_EINVAL, "stft: the spectrogram output needs the persistent kernel (16-byte aligned input, n_samples %% 4 == 0)");
    if (!p->frontend_ok)
      return fail(CACFE_EINVAL, "frontend: n_fft=%d needs the persistent kernel (16-byte aligned input, not forced generic)",
                  p->cfg.n_fft);
    a.bw_in_smem = p->k1.bw_in_smem;
    if (p->nq <= 15)
      cacfe::stft_mel_kernel<15><<<(unsigned)grid, cacfe::kK1Threads, p->k1.total, st>>>(a);
    else
      cacfe::stft_mel_kernel<33><<<(unsigned)grid, cacfe::kK1Threads, p->k1.total, st>>>(a);
  }
  if (p->profile) {
    cudaEventRecord(ev1, st);
    p->k1_events.emplace_back(ev0, ev1);
  }
  if (v3 && layout == CACFE_LAYOUT_BMTC && staging != nullptr) {
    dim3 grid((p->cfg.n_mels + 31) / 32, (p->n_frames + cacfe::kImgT - 1) / cacfe::kImgT, B);
    if (channels == 1)
      cacfe::btm_to_bmtc_kernel<1><<<grid, dim3(32, 8), 0, st>>>(staging, feat, p->n_frames, p->cfg.n_mels, channels);
    else if (channels == 3)
      cacfe::btm_to_bmtc_kernel<3><<<grid, dim3(32, 8), 0, st>>>(staging, feat, p->n_frames, p->cfg.n_mels, channels);
    else
      cacfe::btm_to_bmtc_kernel<0><<<grid, dim3(32, 8), 0, st>>>(staging, feat, p->n_frames, p->cfg.n_mels, channels);
    ++launches;
  }
  return check_launch(p, "frontend", launches);
}

int cacfe_frontend(cacfe_plan* p, const float* raw, float* feat, int B, void* ws, void* stream) {
  if (!p || !raw || !feat) return fail(CACFE_EINVAL, "frontend: null argument");
  if (p->cfg.normalize && !ws) return fail(CACFE_EINVAL, "frontend: workspace required when normalize=1");
  if (B < 1 || B > 65535) return fail(CACFE_ESHAPE, "frontend: B=%d out of range [1, 65535]", B);
  if ((reinterpret_cast<uintptr_t>(raw) & 3) || (reinterpret_cast<uintptr_t>(feat) & 3))
    return fail(CACFE_EALIGN, "frontend: buffers must be 4-byte aligned");
  CUDA_TRY(cudaSetDevice(p->device));
  float* staging = ws ? (float*)((char*)ws + frontend_ws_bytes(B)) : nullptr;  // the mel region of cacfe_workspace_bytes
  return launch_frontend(p, raw, feat, B, p->cfg.out_layout, p->cfg.channels, ws, (cudaStream_t)stream, staging);
}

size_t cacfe_stft_workspace_bytes(const cacfe_plan* p, int B) {   // the min / max partials of the normalisation; no staging buffer
  if (!p || B < 1) return 0;
  return frontend_ws_bytes(B);
}

int cacfe_stft_stats(cacfe_plan* p, const float* raw, float* spec, float* range_min, int B, void* ws, void* stream) {
  if (!p || !raw || !spec || !ws) return fail(CACFE_EINVAL, "stft: null argument");
  if (B < 1) return fail(CACFE_ESHAPE, "stft: B=%d", B);
  if (range_min && !p->cfg.normalize) return fail(CACFE_EINVAL, "stft_stats: the per-clip statistics exist only when normalize=1");
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t per_clip = (size_t)p->n_bins * p->n_frames;
  // the fused kernel writes the stored [b][k][t] layout itself (LAYOUT_SPECT): no staging buffer, no second pass.  (Round 1
  // staged [b][t][k] rows and transposed them: 5.19 ms per 1024 clips against 4.20 ms now.)
  constexpr int kChunk = 16384;   // keeps the tile count inside the kernel's int arithmetic
  for (int b0 = 0; b0 < B; b0 += kChunk) {
    const int nb = B - b0 < kChunk ? B - b0 : kChunk;
    int rc = launch_frontend(p, raw + (size_t)b0 * p->cfg.n_samples, spec + (size_t)b0 * per_clip, nb, cacfe::LAYOUT_SPECT, 1, ws, st);
    if (rc != CACFE_OK) return rc;
    if (range_min)  // the (max - min, min) pairs K0 left in the workspace for this chunk (launch_frontend's layout)
      CUDA_TRY(cudaMemcpyAsync(range_min + 2 * (size_t)b0, (char*)ws + align256((size_t)nb * kMaxSplits * sizeof(float2)),
                               (size_t)nb * sizeof(float2), cudaMemcpyDeviceToDevice, st));
  }
  return CACFE_OK;
}

int cacfe_stft(cacfe_plan* p, const float* raw, float* spec, int B, void* ws, void* stream) {
  return cacfe_stft_stats(p, raw, spec, nullptr, B, ws, stream);
}

int cacfe_mel_from_spectrogram(cacfe_plan* p, const float* spec, float* feat, int B, int T, void* stream) {
  if (!p || !spec || !feat) return fail(CACFE_EINVAL, "mel_from_spectrogram: null argument");
  if (B < 1 || B > 65535 || T < 1) return fail(CACFE_ESHAPE, "mel_from_spectrogram: B=%d T=%d", B, T);
  CUDA_TRY(cudaSetDevice(p->device));
  if (p->cfg.mel_impl == CACFE_MEL_TC_3XTF32) {  // tcgen05 banded 3xTF32 GEMM
    cacfe::MelTcArgs t;
    t.spec = spec;
    t.out = feat;
    t.wpk = p->d_tc_w;
    t.chunks = p->d_tc_chunks;
    t.n_chunks = (int)p->tc_chunks.size();
    t.n_bins = p->n_bins;
    t.T = T;
    t.n_mels = p->cfg.n_mels;
    t.power = p->cfg.power;
    t.channels = p->cfg.channels;
    t.layout = p->cfg.out_layout;
    t.tiles_per_clip = (T + cacfe::kTcM - 1) / cacfe::kTcM;
    t.tile_frames = (T + t.tiles_per_clip - 1) / t.tiles_per_clip;
    const long long grid = (long long)B * t.tiles_per_clip;
    cacfe::melspec_tc_kernel<<<(unsigned)grid, cacfe::kTcThreads, cacfe::kTcSmemBytes, (cudaStream_t)stream>>>(t);
    return check_launch(p, "mel_from_spectrogram (tensor core)");
  }
  if (p->ms_ok && !p->force_generic && p->cfg.out_layout == CACFE_LAYOUT_BMTC &&
      T <= cacfe::kMsConsumers * cacfe::kMsMaxCols) {  // streaming kernel: whole-row TMA chunks
    // two CTAs per SM (each reserves 1 KB) when that leaves four stages, else one
    cacfe::MelStreamGeom g = cacfe::melstream_geometry(T, p->ms_rows, std::min(p->smem_optin, p->smem_per_sm / 2 - 1024));
    int ctas_per_sm = 2;
    if (g.stages < 4) {
      g = cacfe::melstream_geometry(T, p->ms_rows, p->smem_optin);
      ctas_per_sm = 1;
    }
    if (g.stages >= 3) {
      cacfe::MelStreamArgs s{};
      s.spec = spec;
      s.out = feat;
      s.rows = p->d_ms_rows;
      // enough items for ~6 per CTA, so that the static round-robin ends evenly
      const long long want = 6LL * ctas_per_sm * p->sm_count;
      const int v = (long long)B >= want ? 0 : ((long long)B * 2 >= want ? 1 : 2);
      s.n_segs = 1 << v;
      for (int i = 0; i < s.n_segs; ++i) s.segs[i] = p->ms_segs[v][i];
      s.n_items = B * s.n_segs;
      s.n_rows = p->ms_rows;
      s.bin_lo = p->bin_lo;
      s.n_bins = p->n_bins;
      s.T = T;
      s.n_mels = p->cfg.n_mels;
      s.channels = p->cfg.channels;
      s.stages = g.stages;
      s.stage_floats = g.stage_floats;
      s.total_bytes = (size_t)B * p->n_bins * T * sizeof(float);
      int grid = std::min(s.n_items, ctas_per_sm * p->sm_count);
      grid -= grid % s.n_segs;   // the kernel's segment rotation needs whole clips per grid stride (n_items is a multiple already)
      const int cols = (T + cacfe::kMsConsumers - 1) / cacfe::kMsConsumers;
      const bool p2 = p->cfg.power == 2;
      cudaStream_t st = (cudaStream_t)stream;
#define CACFE_MS_LAUNCH2(C_, P_, CH_) cacfe::melspec_stream_kernel<C_, P_, CH_><<<grid, cacfe::kMsThreads, g.smem_bytes, st>>>(s)
#define CACFE_MS_LAUNCH(C_)                                                                                       \
  do {                                                                                                            \
    const int chs = s.channels == 1 ? 1 : (s.channels == 3 ? 3 : 0);                                              \
    if (p2) {                                                                                                     \
      if (chs == 1) CACFE_MS_LAUNCH2(C_, true, 1); else if (chs == 3) CACFE_MS_LAUNCH2(C_, true, 3); else CACFE_MS_LAUNCH2(C_, true, 0);     \
    } else {                                                                                                      \
      if (chs == 1) CACFE_MS_LAUNCH2(C_, false, 1); else if (chs == 3) CACFE_MS_LAUNCH2(C_, false, 3); else CACFE_MS_LAUNCH2(C_, false, 0);  \
    }                                                                                                             \
  } while (0)
      switch (cols) {
        case 1: CACFE_MS_LAUNCH(1); break;
        case 2: CACFE_MS_LAUNCH(2); break;
        case 3: CACFE_MS_LAUNCH(3); break;
        default: CACFE_MS_LAUNCH(4); break;
      }
#undef CACFE_MS_LAUNCH2
#undef CACFE_MS_LAUNCH
      return check_launch(p, "mel_from_spectrogram (stream)");
    }
  }
  cacfe::MelSpecArgs a;
  a.spec = spec;
  a.out = feat;
  a.band_w = p->d_band_w;
  a.band_start = p->d_band_start;
  a.band_ofs = p->d_band_ofs;
  a.n_bins = p->n_bins;
  a.T = T;
  a.n_mels = p->cfg.n_mels;
  a.nnz = p->nnz;
  a.bin_lo = p->bin_lo;
  a.power = p->cfg.power;
  a.channels = p->cfg.channels;
  a.layout = p->cfg.out_layout;
  const size_t smem = sizeof(float) * ((p->nnz + 3) & ~3) + sizeof(int) * (2 * p->cfg.n_mels + 1);
  if (smem > 48 * 1024) return fail(CACFE_EINVAL, "mel_from_spectrogram: filterbank too dense for the banded kernel");
  dim3 grid((T + cacfe::kMelSpecThreads - 1) / cacfe::kMelSpecThreads, B);
  cacfe::melspec_banded_kernel<<<grid, cacfe::kMelSpecThreads, smem, (cudaStream_t)stream>>>(a);
  return check_launch(p, "mel_from_spectrogram");
}

int cacfe_ema_init(cacfe_plan* p, float smooth, const float* in, const float* init, float* out, int B, long long outer_per_clip,
                   int T, int inner, void* stream) {
  if (!p || !in || !out) return fail(CACFE_EINVAL, "ema: null argument");
  if (in == out) return fail(CACFE_EINVAL, "ema: in-place operation is not supported");
  if (B < 1 || B > 65535 || T < 1 || inner < 1 || outer_per_clip < 1) return fail(CACFE_ESHAPE, "ema: bad shape");
  CUDA_TRY(cudaSetDevice(p->device));
  cacfe::PcenArgs a{};
  a.in = in;
  a.out = out;
  a.init = init;
  a.T = T;
  a.inner = inner;
  a.rows_per_clip = (int)(outer_per_clip * inner);
  a.w = fminf(fmaxf(smooth, 0.0f), 1.0f);
  a.one_minus_w = 1.0f - a.w;
  const PcenGrid g = pcen_grid(a.rows_per_clip);
  cacfe::ema_kernel<<<dim3(g.gx, B), g.block, 0, (cudaStream_t)stream>>>(a);
  return check_launch(p, "ema");
}

int cacfe_ema(cacfe_plan* p, float smooth, const float* in, float* out, int B, long long outer_per_clip, int T, int inner,
              void* stream) {
  return cacfe_ema_init(p, smooth, in, nullptr, out, B, outer_per_clip, T, inner, stream);
}

static int launch_pcen(cacfe_plan* p, const cacfe_pcen_params* q, const float* in, float* out, int B,
                       long long outer_per_clip, int T, int inner, void* ws, cudaStream_t st) {
  cacfe::PcenArgs a{};
  int rc = fill_pcen_args(q, a);
  if (rc != CACFE_OK) return rc;
  a.in = in;
  a.out = out;
  a.T = T;
  a.inner = inner;
  a.rows_per_clip = (int)(outer_per_clip * inner);
  // time-contiguous rows with a small inner size (the image layout): the warp-scan kernel, one warp per row
  const long long rows = (long long)B * outer_per_clip;
  const bool scan = inner <= cacfe::kScanMaxC && (long long)T * inner <= cacfe::kScanMaxRow && T >= 32 &&
                    (q->norm_scope != CACFE_NORM_CLIP || outer_per_clip % cacfe::kScanWarps == 0);
  if (scan) {
    const unsigned ctas = (unsigned)((rows + cacfe::kScanWarps - 1) / cacfe::kScanWarps);
    const size_t smem = (size_t)cacfe::kScanWarps * T * inner * sizeof(float);
    if (q->norm_scope == CACFE_NORM_NONE) {
      cacfe::pcen_scan_kernel<cacfe::PCEN_RAW><<<ctas, cacfe::kScanWarps * 32, smem, st>>>(a, rows, outer_per_clip);
      return check_launch(p, "pcen (scan)");
    }
    if (!ws) return fail(CACFE_EINVAL, "pcen: workspace required for the min-max scope");
    a.partial = (float2*)ws;
    float2* ext = (float2*)((char*)ws + align256((size_t)ctas * sizeof(float2)));
    a.extremes = ext;
    a.per_clip_extremes = q->norm_scope == CACFE_NORM_CLIP;
    cacfe::pcen_scan_kernel<cacfe::PCEN_REDUCE><<<ctas, cacfe::kScanWarps * 32, smem, st>>>(a, rows, outer_per_clip);
    if (a.per_clip_extremes)
      cacfe::minmax_finalize_kernel<true><<<B, 256, 0, st>>>(a.partial, (int)(outer_per_clip / cacfe::kScanWarps), ext);
    else
      cacfe::minmax_finalize_kernel<true><<<1, 256, 0, st>>>(a.partial, (int)ctas, ext);
    cacfe::pcen_scan_kernel<cacfe::PCEN_APPLY><<<ctas, cacfe::kScanWarps * 32, smem, st>>>(a, rows, outer_per_clip);
    return check_launch(p, "pcen (scan)", 3);
  }
  const PcenGrid g = pcen_grid(a.rows_per_clip);
  dim3 grid(g.gx, B);
  static const int dyn = [] {   // residency probe (tools/probe_pcen_waves.py --residency): unused dynamic shared memory per block
    const char* e = getenv("CACFE_PCEN_DYN_SMEM");
    const int v = e ? atoi(e) : 0;
    if (v > 0) {
      cudaFuncSetAttribute((const void*)cacfe::pcen_kernel<cacfe::PCEN_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
      cudaFuncSetAttribute((const void*)cacfe::pcen_kernel<cacfe::PCEN_REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
      cudaFuncSetAttribute((const void*)cacfe::pcen_kernel<cacfe::PCEN_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
      cudaFuncSetAttribute((const void*)cacfe::pcen_kernel<cacfe::PCEN_RAW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
      cudaFuncSetAttribute((const void*)cacfe::pcen_kernel<cacfe::PCEN_REDUCE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
      cudaFuncSetAttribute((const void*)cacfe::pcen_kernel<cacfe::PCEN_APPLY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    }
    return v;
  }();
  if (q->norm_scope == CACFE_NORM_NONE) {
    pcen_pass<cacfe::PCEN_RAW>(p, grid, g.block, dyn, st, a);
    return check_launch(p, "pcen");
  }
  if (!ws) return fail(CACFE_EINVAL, "pcen: workspace required for the min-max scope");
  float2* partial = (float2*)ws;
  float2* extremes = (float2*)((char*)ws + align256((size_t)B * g.gx * sizeof(float2)));
  a.partial = partial;
  a.extremes = extremes;
  a.per_clip_extremes = q->norm_scope == CACFE_NORM_CLIP;
  pcen_pass<cacfe::PCEN_REDUCE>(p, grid, g.block, dyn, st, a);
  if (a.per_clip_extremes)
    cacfe::pcen_extremes_kernel<<<B, 256, 0, st>>>(partial, g.gx, extremes, a, 0);
  else
    cacfe::pcen_extremes_kernel<<<1, 256, 0, st>>>(partial, B * g.gx, extremes, a, 0);
  pcen_pass<cacfe::PCEN_APPLY>(p, grid, g.block, dyn, st, a);
  return check_launch(p, "pcen", 3);
}

int cacfe_pcen(cacfe_plan* p, const cacfe_pcen_params* q, const float* in, float* out, int B, long long outer_per_clip,
               int T, int inner, void* ws, void* stream) {
  if (!p || !q || !in || !out) return fail(CACFE_EINVAL, "pcen: null argument");
  if (in == out) return fail(CACFE_EINVAL, "pcen: in-place operation is not supported");
  if (B < 1 || B > 65535 || T < 1 || inner < 1 || outer_per_clip < 1 || outer_per_clip * inner > 2147483647LL)
    return fail(CACFE_ESHAPE, "pcen: bad shape");
  if (q->norm_scope < 0 || q->norm_scope > 2) return fail(CACFE_EINVAL, "pcen: unknown norm_scope");
  CUDA_TRY(cudaSetDevice(p->device));
  return launch_pcen(p, q, in, out, B, outer_per_clip, T, inner, ws, (cudaStream_t)stream);
}

// ---- PCEN backward -----------------------------------------------------------------------------
// lanes of a clip spread evenly over as few blocks of at most kBwdMaxThreads as it takes (160 mel lanes -> one block of 160)
static PcenGrid pcen_bwd_grid(long long rows_per_clip) {
  PcenGrid g;
  g.gx = (int)((rows_per_clip + cacfe::kBwdMaxThreads - 1) / cacfe::kBwdMaxThreads);
  const long long per = (rows_per_clip + g.gx - 1) / g.gx;
  g.block = (int)(((per + 31) / 32) * 32);
  return g;
}

static size_t pcen_bwd_layout(int B, long long rows_per_clip, size_t off[4]) {
  const PcenGrid g = pcen_grid(rows_per_clip);
  const int gx2 = pcen_bwd_grid(rows_per_clip).gx;
  const int gmax = g.gx > gx2 ? g.gx : gx2;
  size_t o = 0;
  off[0] = o; o += align256((size_t)B * g.gx * sizeof(float2));          // forward block extremes
  off[1] = o; o += align256((size_t)B * sizeof(float2));                 // (mn, mx) per entry
  off[2] = o; o += align256((size_t)B * gmax * 4 * sizeof(double));      // block partial sums
  off[3] = o; o += align256((size_t)B * sizeof(float4));                 // folded min-max terms
  return o;
}

size_t cacfe_pcen_backward_workspace_bytes(int B, long long outer_per_clip, int inner) {
  if (B < 1 || outer_per_clip < 1 || inner < 1) return 0;
  size_t off[4];
  return pcen_bwd_layout(B, outer_per_clip * inner, off);
}

int cacfe_pcen_backward(cacfe_plan* p, const cacfe_pcen_params* q, const float* x, const float* grad_out, float* grad_x,
                        float* grad_params, int B, long long outer_per_clip, int T, int inner, void* ws, void* stream) {
  if (!p || !q || !x || !grad_out || !grad_x || !grad_params || !ws) return fail(CACFE_EINVAL, "pcen_backward: null argument");
  if (B < 1 || B > 65535 || T < 1 || inner < 1 || outer_per_clip < 1 || outer_per_clip * inner > 2147483647LL)
    return fail(CACFE_ESHAPE, "pcen_backward: bad shape");
  if (T > cacfe::kBwdSeg * cacfe::kBwdMaxSeg)
    return fail(CACFE_ESHAPE, "pcen_backward: T=%d exceeds %d time steps", T, cacfe::kBwdSeg * cacfe::kBwdMaxSeg);
  if (q->norm_scope < 0 || q->norm_scope > 2) return fail(CACFE_EINVAL, "pcen_backward: unknown norm_scope");
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  cacfe::PcenBwdArgs a{};
  int rc = fill_pcen_args(q, a.f);
  if (rc != CACFE_OK) return rc;
  a.f.in = x;
  a.f.T = T;
  a.f.inner = inner;
  a.f.rows_per_clip = (int)(outer_per_clip * inner);
  a.g = grad_out;
  a.dx = grad_x;
  a.gain_raw = q->gain;
  a.root_raw = q->root;
  a.smooth_raw = q->smooth;
  a.scope = q->norm_scope;
  size_t off[4];
  pcen_bwd_layout(B, a.f.rows_per_clip, off);
  char* w = (char*)ws;
  float2* extremes = (float2*)(w + off[1]);
  double* partial = (double*)(w + off[2]);
  float4* fold = (float4*)(w + off[3]);
  a.extremes = extremes;
  a.partial = partial;
  a.fold = fold;
  const PcenGrid g = pcen_grid(a.f.rows_per_clip);
  int launches = 2;
  if (q->norm_scope != CACFE_NORM_NONE) {
    const int entries = q->norm_scope == CACFE_NORM_CLIP ? B : 1;
    const int per_entry = q->norm_scope == CACFE_NORM_CLIP ? g.gx : B * g.gx;
    a.f.partial = (float2*)(w + off[0]);
    pcen_pass<cacfe::PCEN_REDUCE>(p, dim3(g.gx, B), g.block, 0, st, a.f);
    cacfe::pcen_extremes_kernel<<<entries, 256, 0, st>>>(a.f.partial, per_entry, extremes, a.f, 1);   // (mn, mx) of p
    cacfe::pcen_bwd_reduce_kernel<<<dim3(g.gx, B), g.block, 0, st>>>(a);
    cacfe::pcen_bwd_fold_kernel<<<entries, 32, 0, st>>>(partial, per_entry, extremes, fold);
    launches += 4;
  }
  const PcenGrid g2 = pcen_bwd_grid(a.f.rows_per_clip);
  const int gx2 = g2.gx;
  const size_t ck = (size_t)((T + cacfe::kBwdSeg - 1) / cacfe::kBwdSeg) * g2.block * sizeof(float);
  cacfe::pcen_bwd_kernel<<<dim3(gx2, B), g2.block, ck, st>>>(a);
  cacfe::pcen_bwd_params_kernel<<<1, 32, 0, st>>>(partial, B * gx2, q->gain, q->root, q->smooth, grad_params);
  return check_launch(p, "pcen_backward", launches);
}

// ---- identifytracks.signal_noise, spectrogram -> connected components ---------------------------
static size_t signal_layout(int K, int T, size_t off[9]) {
  const size_t n = (size_t)K * T;
  size_t o = 0;
  off[0] = o; o += align256(1024 * sizeof(cacfe::Stats));   // block partials of the global maximum
  off[1] = o; o += align256(sizeof(cacfe::Stats));
  off[2] = o; o += align256(n * sizeof(float));              // [T][K] copy for the column medians
  off[3] = o; o += align256((size_t)K * sizeof(float));      // row medians
  off[4] = o; o += align256((size_t)T * sizeof(float));      // column medians
  off[5] = o; o += align256(n);                              // mask A
  off[6] = o; o += align256(n);                              // mask B
  off[7] = o; o += align256(n * sizeof(int));                // labels
  off[8] = o; o += align256(n * sizeof(int));                // slot of every root
  return o;
}

size_t cacfe_signal_workspace_bytes(int K, int T) {
  if (K < 1 || T < 1) return 0;
  size_t off[9];
  return signal_layout(K, T, off);
}

int cacfe_signal_components(cacfe_plan* p, const float* spec, int K, int T, int open_size, int dil_h, int dil_w, int ero_h,
                            int ero_w, unsigned char* mask_out, unsigned char* raw_mask_out, float* row_medians_out,
                            float* col_medians_out, int32_t* comps_out, int max_components, int32_t* n_components_out, void* ws,
                            void* stream) {
  if (!p || !spec || !ws || !comps_out || !n_components_out) return fail(CACFE_EINVAL, "signal_components: null argument");
  if (K < 1 || T < 1 || (long long)K * T > 2147483647LL) return fail(CACFE_ESHAPE, "signal_components: K=%d T=%d", K, T);
  if (open_size < 1 || dil_h < 1 || dil_w < 1 || ero_h < 1 || ero_w < 1 || max_components < 1)
    return fail(CACFE_EINVAL, "signal_components: structuring elements and max_components must be positive");
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  size_t off[9];
  signal_layout(K, T, off);
  char* w = (char*)ws;
  const long long n = (long long)K * T;
  cacfe::Stats* partial = (cacfe::Stats*)(w + off[0]);
  cacfe::Stats* stats = (cacfe::Stats*)(w + off[1]);
  float* spec_tk = (float*)(w + off[2]);
  float* rm = (float*)(w + off[3]);
  float* cm = (float*)(w + off[4]);
  unsigned char* ma = (unsigned char*)(w + off[5]);
  unsigned char* mb = (unsigned char*)(w + off[6]);
  int* labels = (int*)(w + off[7]);
  int* slot_of = (int*)(w + off[8]);
  int launches = 0;
  // a_max (identifytracks.py:79)
  const int sblocks = (int)std::min<long long>(1024, (n + 4095) / 4096);
  cacfe::stats_kernel<<<dim3(sblocks, 1), 256, 0, st>>>(spec, n, partial);
  cacfe::stats_finalize_kernel<<<1, 32, 0, st>>>(partial, sblocks, stats);
  // medians (:81-82): rows of [K][T]; columns = rows of the [T][K] copy
  {
    dim3 grid((T + cacfe::kTrTile - 1) / cacfe::kTrTile, (K + cacfe::kTrTile - 1) / cacfe::kTrTile, 1);
    cacfe::spec_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(spec, spec_tk, K, T);   // in [K][T] -> out [T][K]
  }
  cacfe::select_median_kernel<<<K, 256, 0, st>>>(spec, T, stats, rm);
  cacfe::select_median_kernel<<<T, 256, 0, st>>>(spec_tk, K, stats, cm);
  const int eblocks = (int)std::min<long long>(8 * p->sm_count, (n + 255) / 256);
  cacfe::signal_mask_kernel<<<eblocks, 256, 0, st>>>(spec, K, T, stats, rm, cm, ma);
  launches += 6;
  if (raw_mask_out) CUDA_TRY(cudaMemcpyAsync(raw_mask_out, ma, (size_t)n, cudaMemcpyDeviceToDevice, st));
  if (row_medians_out) CUDA_TRY(cudaMemcpyAsync(row_medians_out, rm, (size_t)K * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (col_medians_out) CUDA_TRY(cudaMemcpyAsync(col_medians_out, cm, (size_t)T * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // morphology (:94-101): open, dilate, erode -- each a horizontal and a vertical pass
  unsigned char *src = ma, *dst = mb;
  auto pass = [&](bool is_max, int size, int horizontal) {
    if (is_max)
      cacfe::morph_pass_kernel<true><<<eblocks, 256, 0, st>>>(src, dst, K, T, size, size / 2, horizontal);
    else
      cacfe::morph_pass_kernel<false><<<eblocks, 256, 0, st>>>(src, dst, K, T, size, size / 2, horizontal);
    std::swap(src, dst);
    ++launches;
  };
  pass(false, open_size, 1);
  pass(false, open_size, 0);
  pass(true, open_size, 1);
  pass(true, open_size, 0);
  pass(true, dil_w, 1);
  pass(true, dil_h, 0);
  pass(false, ero_w, 1);
  pass(false, ero_h, 0);
  if (mask_out) CUDA_TRY(cudaMemcpyAsync(mask_out, src, (size_t)n, cudaMemcpyDeviceToDevice, st));
  // connected components with stats (:106)
  cacfe::Component* comps = reinterpret_cast<cacfe::Component*>(comps_out);
  cacfe::ccl_clear_kernel<<<(max_components + 255) / 256, 256, 0, st>>>(comps, max_components, n_components_out);
  cacfe::ccl_init_kernel<<<eblocks, 256, 0, st>>>(src, labels, K, T);
  cacfe::ccl_merge_kernel<<<eblocks, 256, 0, st>>>(src, labels, K, T);
  cacfe::ccl_flatten_kernel<<<eblocks, 256, 0, st>>>(labels, n, slot_of, n_components_out, max_components);
  cacfe::ccl_stats_kernel<<<eblocks, 256, 0, st>>>(labels, K, T, slot_of, comps);
  launches += 5;
  return check_launch(p, "signal_components", launches);
}

int cacfe_mix_up(cacfe_plan* p, const float* one, const float* two, const float* lambda, float* out, int B, long long per_entry,
                 void* stream) {
  if (!p || !one || !two || !lambda || !out) return fail(CACFE_EINVAL, "mix_up: null argument");
  if (B < 1 || B > 65535 || per_entry < 1) return fail(CACFE_ESHAPE, "mix_up: bad shape");
  CUDA_TRY(cudaSetDevice(p->device));
  long long gx = (per_entry + 256 * 16 - 1) / (256 * 16);
  if (gx > 1024) gx = 1024;
  cacfe::mix_up_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(one, two, lambda, out, per_entry);
  return check_launch(p, "mix_up");
}

int cacfe_compress(cacfe_plan* p, int mode, float param, const float* in, float* out, long long entries,
                   long long per_entry, void* ws, void* stream) {
  if (!p || !in || !out) return fail(CACFE_EINVAL, "compress: null argument");
  if (in == out) return fail(CACFE_EINVAL, "compress: in-place operation is not supported");
  if (entries < 1 || entries > 65535 || per_entry < 1) return fail(CACFE_ESHAPE, "compress: bad shape");
  if (mode < 0 || mode > 4) return fail(CACFE_EINVAL, "compress: unknown mode %d", mode);
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = compress_blocks(entries, per_entry);
  dim3 grid(blocks, (unsigned)entries);
  int launches = 1;
  cacfe::Stats* stats = nullptr;
  if (mode != CACFE_COMPRESS_MAG_POW) {
    if (!ws) return fail(CACFE_EINVAL, "compress: workspace required for this mode");
    cacfe::Stats* partial = (cacfe::Stats*)ws;
    stats = (cacfe::Stats*)((char*)ws + align256((size_t)entries * blocks * sizeof(cacfe::Stats)));
    cacfe::stats_kernel<<<grid, 256, 0, st>>>(in, per_entry, partial);
    cacfe::stats_finalize_kernel<<<(unsigned)entries, 32, 0, st>>>(partial, blocks, stats);
    launches = 3;
  }
  switch (mode) {
    case CACFE_COMPRESS_MAG_POW:
      cacfe::compress_kernel<cacfe::COMPRESS_MAG_POW><<<grid, 256, 0, st>>>(in, out, per_entry, param, stats);
      break;
    case CACFE_COMPRESS_POWER_TO_DB:
      cacfe::compress_kernel<cacfe::COMPRESS_POWER_TO_DB><<<grid, 256, 0, st>>>(in, out, per_entry, param, stats);
      break;
    case CACFE_COMPRESS_MINMAX:
      cacfe::compress_kernel<cacfe::COMPRESS_MINMAX><<<grid, 256, 0, st>>>(in, out, per_entry, param, stats);
      break;
    case CACFE_COMPRESS_MEAN_SUB:
      cacfe::compress_kernel<cacfe::COMPRESS_MEAN_SUB><<<grid, 256, 0, st>>>(in, out, per_entry, param, stats);
      break;
    default:
      cacfe::compress_kernel<cacfe::COMPRESS_STD><<<grid, 256, 0, st>>>(in, out, per_entry, param, stats);
      break;
  }
  return check_launch(p, "compress", launches);
}

int cacfe_sosfilt(cacfe_plan* p, const double* sos_host, int n_sections, const float* in, float* out, long long rows,
                  long long n, void* stream) {
  if (!p || !sos_host || !in || !out) return fail(CACFE_EINVAL, "sosfilt: null argument");
  if (n_sections < 1 || n_sections > cacfe::kSosMaxSections)
    return fail(CACFE_EINVAL, "sosfilt: 1..%d sections supported (got %d)", cacfe::kSosMaxSections, n_sections);
  if (rows < 1 || n < 1) return fail(CACFE_ESHAPE, "sosfilt: bad shape");
  CUDA_TRY(cudaSetDevice(p->device));
  thread_local cacfe::SosArgs a;   // 9 KB of kernel parameters, staged per host thread
  a.in = in;
  a.out = out;
  a.rows = rows;
  a.n = n;
  a.n_sections = n_sections;
  for (int s = 0; s < n_sections; ++s) {
    const double a0 = sos_host[6 * s + 3];
    if (a0 == 0.0) return fail(CACFE_EINVAL, "sosfilt: a0 == 0 in section %d", s);
    cacfe::SosSection& q = a.sec[s];
    q.b0 = sos_host[6 * s + 0] / a0;   // scipy normalises by a0 as well
    q.b1 = sos_host[6 * s + 1] / a0;
    q.b2 = sos_host[6 * s + 2] / a0;
    q.a1 = sos_host[6 * s + 4] / a0;
    q.a2 = sos_host[6 * s + 5] / a0;
    q.c1 = q.b1 - q.a1 * q.b0;
    q.c2 = q.b2 - q.a2 * q.b0;
    // P = A^32 with A = [[-a1, 1], [-a2, 0]] and its powers 0..32, in long double
    long double A[4] = {-(long double)q.a1, 1.0L, -(long double)q.a2, 0.0L}, P[4] = {1.0L, 0.0L, 0.0L, 1.0L};
    auto mul = [](const long double* x, const long double* y, long double* z) {
      const long double r[4] = {x[0] * y[0] + x[1] * y[2], x[0] * y[1] + x[1] * y[3], x[2] * y[0] + x[3] * y[2], x[2] * y[1] + x[3] * y[3]};
      for (int i = 0; i < 4; ++i) z[i] = r[i];
    };
    for (int i = 0; i < cacfe::kSosRun; ++i) mul(P, A, P);
    long double M[4] = {1.0L, 0.0L, 0.0L, 1.0L};
    for (int m = 0; m <= cacfe::kSosRun; ++m) {
      for (int i = 0; i < 4; ++i) q.pw[m][i] = (double)M[i];
      mul(M, P, M);
    }
  }
  if (rows > 2147483647LL) return fail(CACFE_ESHAPE, "sosfilt: too many rows for one launch");
  cacfe::sosfilt_scan_kernel<<<(unsigned)rows, cacfe::kSosThreads, 0, (cudaStream_t)stream>>>(a);
  return check_launch(p, "sosfilt");
}

int cacfe_frontend_pcen(cacfe_plan* p, const cacfe_pcen_params* q, const float* raw, float* out, int B, void* ws,
                        void* stream) {
  if (!p || !q || !raw || !out || !ws) return fail(CACFE_EINVAL, "frontend_pcen: null argument");
  if (B < 1 || B > 65535) return fail(CACFE_ESHAPE, "frontend_pcen: B=%d out of range", B);
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)ws;
  void* ws_front = w;
  float* mel = (float*)(w + frontend_ws_bytes(B));
  void* ws_pcen = w + frontend_ws_bytes(B) + align256((size_t)B * p->n_frames * p->cfg.n_mels * sizeof(float));
  int rc = launch_frontend(p, raw, mel, B, CACFE_LAYOUT_BTM, 1, ws_front, st);
  if (rc != CACFE_OK) return rc;
  return launch_pcen(p, q, mel, out, B, 1, p->n_frames, p->cfg.n_mels, ws_pcen, st);
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer pipeline
// ---------------------------------------------------------------------------------------------------------
}  // extern "C"

struct cacfe_hostpipe {
  cacfe_plan* plan = nullptr;
  int max_B = 0, chunk = 0;
  cudaStream_t stream[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaEvent_t all_reduced = nullptr;
  float* d_in[2] = {nullptr, nullptr};   // [chunk][n_samples]
  short* d_in16[2] = {nullptr, nullptr};  // [chunk][n_samples] 16-bit PCM staging, allocated by the first cacfe_hostpipe_run_pcm16
  float* d_feat = nullptr;               // [max_B] mel features (layout of the run)
  float* d_out[2] = {nullptr, nullptr};  // [chunk] PCEN output staging
  void* d_ws[2] = {nullptr, nullptr};    // per-stream front-end workspace
  float2* d_partial = nullptr;           // [max_B][gx]
  float2* d_extremes = nullptr;          // [max_B]
  size_t bytes = 0;
};

extern "C" {

void cacfe_hostpipe_destroy(cacfe_hostpipe* h) {
  if (!h) return;
  cudaSetDevice(h->plan->device);
  for (int i = 0; i < 2; ++i) {
    if (h->stream[i]) cudaStreamSynchronize(h->stream[i]);
    cudaFree(h->d_in[i]);
    cudaFree(h->d_in16[i]);
    cudaFree(h->d_out[i]);
    cudaFree(h->d_ws[i]);
    if (h->done[i]) cudaEventDestroy(h->done[i]);
    if (h->stream[i]) cudaStreamDestroy(h->stream[i]);
  }
  if (h->all_reduced) cudaEventDestroy(h->all_reduced);
  cudaFree(h->d_feat);
  cudaFree(h->d_partial);
  cudaFree(h->d_extremes);
  delete h;
}

int cacfe_hostpipe_create(cacfe_plan* p, int max_B, int chunk, cacfe_hostpipe** out) {
  if (!p || !out) return fail(CACFE_EINVAL, "hostpipe_create: null argument");
  *out = nullptr;
  if (max_B < 1 || max_B > 65535 || chunk < 1) return fail(CACFE_ESHAPE, "hostpipe_create: max_B=%d chunk=%d", max_B, chunk);
  if (chunk > max_B) chunk = max_B;
  CUDA_TRY(cudaSetDevice(p->device));
  cacfe_hostpipe* h = new cacfe_hostpipe();
  h->plan = p;
  h->max_B = max_B;
  h->chunk = chunk;
  const size_t feat_clip = (size_t)p->n_frames * p->cfg.n_mels * p->cfg.channels * sizeof(float);
  const size_t pcen_clip = (size_t)p->n_frames * p->cfg.n_mels * sizeof(float);
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) {
    if (e == cudaSuccess) {
      e = cudaMalloc(ptr, bytes);
      if (e == cudaSuccess) h->bytes += bytes;
    }
  };
  for (int i = 0; i < 2; ++i) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->done[i], cudaEventDisableTiming);
    alloc((void**)&h->d_in[i], (size_t)chunk * p->cfg.n_samples * sizeof(float));
    alloc((void**)&h->d_out[i], (size_t)chunk * pcen_clip);
    alloc(&h->d_ws[i], frontend_ws_bytes(chunk) + align256((size_t)chunk * pcen_clip));  // + [chunk][T][M] staging
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->all_reduced, cudaEventDisableTiming);
  alloc((void**)&h->d_feat, (size_t)max_B * (feat_clip > pcen_clip ? feat_clip : pcen_clip));
  alloc((void**)&h->d_partial, (size_t)max_B * pcen_grid(p->cfg.n_mels).gx * sizeof(float2));
  alloc((void**)&h->d_extremes, (size_t)max_B * sizeof(float2));
  if (e != cudaSuccess) {
    cacfe_hostpipe_destroy(h);
    return fail(e == cudaErrorMemoryAllocation ? CACFE_ENOMEM : CACFE_ECUDA, "hostpipe_create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return CACFE_OK;
}

size_t cacfe_hostpipe_device_bytes(const cacfe_hostpipe* h) { return h ? h->bytes : 0; }

}  // extern "C"

// One body for both host formats: `host_in` (float32) or `host_pcm` (16-bit PCM, converted on the device as s / 32768 -- what
// soundfile / librosa.load return for a 16-bit file, audiowriter.py:352 -- so that the features are bit-identical to the
// float32 call on the converted samples while the upload moves half the bytes).
static int hostpipe_run_impl(cacfe_hostpipe* h, const cacfe_pcen_params* q, const float* host_in, const short* host_pcm,
                             float* host_out, int B) {
  if (!h || (!host_in && !host_pcm) || !host_out) return fail(CACFE_EINVAL, "hostpipe_run: null argument");
  if (B < 1 || B > h->max_B) return fail(CACFE_ESHAPE, "hostpipe_run: B=%d exceeds max_B=%d", B, h->max_B);
  cacfe_plan* p = h->plan;
  CUDA_TRY(cudaSetDevice(p->device));
  const int ns = p->cfg.n_samples;
  const int nchunks = (B + h->chunk - 1) / h->chunk;
  if (host_pcm) {
    for (int i = 0; i < 2; ++i)
      if (!h->d_in16[i]) {
        const size_t bytes = (size_t)h->chunk * ns * sizeof(short);
        cudaError_t e = cudaMalloc((void**)&h->d_in16[i], bytes);
        if (e != cudaSuccess)
          return fail(e == cudaErrorMemoryAllocation ? CACFE_ENOMEM : CACFE_ECUDA, "hostpipe_run_pcm16: %s", cudaGetErrorString(e));
        h->bytes += bytes;
      }
  }
  // chunk b0.. of the caller's buffer -> d_in[s] on stream st
  auto upload = [&](int s, int b0, int nb, cudaStream_t st) -> cudaError_t {
    if (!host_pcm)
      return cudaMemcpyAsync(h->d_in[s], host_in + (size_t)b0 * ns, (size_t)nb * ns * sizeof(float), cudaMemcpyHostToDevice, st);
    cudaError_t e = cudaMemcpyAsync(h->d_in16[s], host_pcm + (size_t)b0 * ns, (size_t)nb * ns * sizeof(short),
                                    cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    const long long n = (long long)nb * ns;
    cacfe::pcm16_to_f32_kernel<<<(unsigned)std::min<long long>((n / 8 + 255) / 256 + 1, 148 * 16), 256, 0, st>>>(h->d_in16[s], h->d_in[s], n);
    p->launches.fetch_add(1);
    return cudaGetLastError();
  };
  if (q == nullptr) {
    // mel image only: every chunk is independent -> H2D / kernels / D2H fully pipelined on two streams
    const size_t feat_clip = (size_t)p->n_frames * p->cfg.n_mels * p->cfg.channels;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1, b0 = c * h->chunk, nb = (B - b0 < h->chunk) ? B - b0 : h->chunk;
      cudaStream_t st = h->stream[s];
      CUDA_TRY(upload(s, b0, nb, st));
      float* feat = h->d_feat + (size_t)b0 * feat_clip;
      int rc = launch_frontend(p, h->d_in[s], feat, nb, p->cfg.out_layout, p->cfg.channels, h->d_ws[s], st,
                               (float*)((char*)h->d_ws[s] + frontend_ws_bytes(h->chunk)));
      if (rc != CACFE_OK) return rc;
      CUDA_TRY(cudaMemcpyAsync(host_out + (size_t)b0 * feat_clip, feat, (size_t)nb * feat_clip * sizeof(float),
                               cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream[0]));
    CUDA_TRY(cudaStreamSynchronize(h->stream[1]));
    return CACFE_OK;
  }
  if (q->norm_scope < 0 || q->norm_scope > 2) return fail(CACFE_EINVAL, "hostpipe_run: unknown norm_scope");
  cacfe::PcenArgs a{};
  int rc = fill_pcen_args(q, a);
  if (rc != CACFE_OK) return rc;
  const int T = p->n_frames, M = p->cfg.n_mels;
  const size_t clip = (size_t)T * M;
  const PcenGrid g = pcen_grid(M);
  a.T = T;
  a.inner = M;
  a.rows_per_clip = M;
  a.per_clip_extremes = q->norm_scope == CACFE_NORM_CLIP;
  a.extremes = h->d_extremes;
  const bool global = q->norm_scope == CACFE_NORM_TENSOR;
  // phase 1: per chunk H2D -> min/max -> mel [b][T][M] (device resident) -> PCEN reduce (or the final PCEN
  // when the scope allows the chunk to finish on its own).
  // With the tensor-global scope a run is an upload phase followed by a download phase; several pipes driven from
  // several host threads (the way to keep both PCIe directions busy) must not split the upload link between them,
  // so the upload phase of a device is taken by one pipe at a time.
  std::unique_lock<std::mutex> upload_gate;
  if (global) upload_gate = std::unique_lock<std::mutex>(g_upload_gate[p->device & 15]);
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1, b0 = c * h->chunk, nb = (B - b0 < h->chunk) ? B - b0 : h->chunk;
    cudaStream_t st = h->stream[s];
    CUDA_TRY(upload(s, b0, nb, st));
    float* mel = h->d_feat + (size_t)b0 * clip;
    rc = launch_frontend(p, h->d_in[s], mel, nb, CACFE_LAYOUT_BTM, 1, h->d_ws[s], st);
    if (rc != CACFE_OK) return rc;
    cacfe::PcenArgs ac = a;
    ac.in = mel;
    ac.out = h->d_out[s];
    ac.partial = h->d_partial + (size_t)b0 * g.gx;
    ac.extremes = h->d_extremes + b0;
    if (q->norm_scope == CACFE_NORM_NONE) {
      pcen_pass<cacfe::PCEN_RAW>(p, dim3(g.gx, nb), g.block, 0, st, ac);
    } else {
      pcen_pass<cacfe::PCEN_REDUCE>(p, dim3(g.gx, nb), g.block, 0, st, ac);
      if (!global) {
        cacfe::pcen_extremes_kernel<<<nb, 256, 0, st>>>(ac.partial, g.gx, h->d_extremes + b0, ac, 0);
        pcen_pass<cacfe::PCEN_APPLY>(p, dim3(g.gx, nb), g.block, 0, st, ac);
      }
    }
    if ((rc = check_launch(p, "hostpipe", global ? 1 : (q->norm_scope == CACFE_NORM_NONE ? 1 : 3))) != CACFE_OK) return rc;
    if (!global)
      CUDA_TRY(cudaMemcpyAsync(host_out + (size_t)b0 * clip, h->d_out[s], (size_t)nb * clip * sizeof(float),
                               cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(h->done[s], st));
  }
  if (global) {
    // the uploads are in flight: wait for the last one of each stream, then let the next pipe upload
    CUDA_TRY(cudaEventSynchronize(h->done[0]));
    if (nchunks > 1) CUDA_TRY(cudaEventSynchronize(h->done[1]));
    upload_gate.unlock();
    // phase 2: tensor-global extremes need every chunk (tfpcen.py:105-110)
    CUDA_TRY(cudaStreamWaitEvent(h->stream[0], h->done[1], 0));
    cacfe::pcen_extremes_kernel<<<1, 256, 0, h->stream[0]>>>(h->d_partial, B * g.gx, h->d_extremes, a, 0);
    if ((rc = check_launch(p, "hostpipe")) != CACFE_OK) return rc;
    CUDA_TRY(cudaEventRecord(h->all_reduced, h->stream[0]));
    CUDA_TRY(cudaStreamWaitEvent(h->stream[1], h->all_reduced, 0));
    // phase 3: apply + D2H per chunk, alternating streams
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1, b0 = c * h->chunk, nb = (B - b0 < h->chunk) ? B - b0 : h->chunk;
      cudaStream_t st = h->stream[s];
      cacfe::PcenArgs ac = a;
      ac.in = h->d_feat + (size_t)b0 * clip;
      ac.out = h->d_out[s];
      ac.extremes = h->d_extremes;
      ac.per_clip_extremes = 0;
      pcen_pass<cacfe::PCEN_APPLY>(p, dim3(g.gx, nb), g.block, 0, st, ac);
      if ((rc = check_launch(p, "hostpipe")) != CACFE_OK) return rc;
      CUDA_TRY(cudaMemcpyAsync(host_out + (size_t)b0 * clip, h->d_out[s], (size_t)nb * clip * sizeof(float),
                               cudaMemcpyDeviceToHost, st));
    }
  }
  CUDA_TRY(cudaStreamSynchronize(h->stream[0]));
  CUDA_TRY(cudaStreamSynchronize(h->stream[1]));
  return CACFE_OK;
}

extern "C" {

int cacfe_hostpipe_run(cacfe_hostpipe* h, const cacfe_pcen_params* q, const float* host_in, float* host_out, int B) {
  return hostpipe_run_impl(h, q, host_in, nullptr, host_out, B);
}

int cacfe_hostpipe_run_pcm16(cacfe_hostpipe* h, const cacfe_pcen_params* q, const int16_t* host_pcm, float* host_out, int B) {
  if (!host_pcm) return fail(CACFE_EINVAL, "hostpipe_run_pcm16: null argument");
  return hostpipe_run_impl(h, q, nullptr, reinterpret_cast<const short*>(host_pcm), host_out, B);
}

int cacfe_pcm16_to_f32(cacfe_plan* p, const int16_t* in_dev, float* out_dev, long long n, void* stream) {
  if (!p || !in_dev || !out_dev) return fail(CACFE_EINVAL, "pcm16_to_f32: null argument");
  if (n < 1) return fail(CACFE_ESHAPE, "pcm16_to_f32: n=%lld", n);
  CUDA_TRY(cudaSetDevice(p->device));
  cacfe::pcm16_to_f32_kernel<<<(unsigned)std::min<long long>((n / 8 + 255) / 256 + 1, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const short*>(in_dev), out_dev, n);
  return check_launch(p, "pcm16_to_f32", 1);
}

int cacfe_host_register(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(CACFE_EINVAL, "host_register: null argument");
  CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return CACFE_OK;
}

int cacfe_host_unregister(void* ptr) {
  if (!ptr) return fail(CACFE_EINVAL, "host_unregister: null argument");
  CUDA_TRY(cudaHostUnregister(ptr));
  return CACFE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// DLPack
// ---------------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {
// dlpack.h v0.8 ABI (stable since v0.2): only the fields read here.
struct DLDevice { int32_t device_type; int32_t device_id; };
struct DLDataType { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensor {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides;
  uint64_t byte_offset;
};
struct DLManagedTensorABI {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(DLManagedTensorABI*);
};
constexpr int kDLCUDA = 2, kDLFloat = 2;

int dl_check(const cacfe_plan* p, const DLManagedTensorABI* m, const char* what, const DLTensor** out) {
  if (!m) return fail(CACFE_EINVAL, "%s: null DLManagedTensor", what);
  const DLTensor& t = m->dl_tensor;
  if (t.device.device_type != kDLCUDA) return fail(CACFE_EDEVICE, "%s: tensor is not on a CUDA device", what);
  if (t.device.device_id != p->device)
    return fail(CACFE_EDEVICE, "%s: tensor on cuda:%d, plan on cuda:%d", what, t.device.device_id, p->device);
  if (t.dtype.code != kDLFloat || t.dtype.bits != 32 || t.dtype.lanes != 1) return fail(CACFE_EDTYPE, "%s: float32 required", what);
  if (t.strides) {
    int64_t expect = 1;
    for (int i = t.ndim - 1; i >= 0; --i) {
      if (t.shape[i] != 1 && t.strides[i] != expect) return fail(CACFE_ESHAPE, "%s: tensor must be C-contiguous", what);
      expect *= t.shape[i];
    }
  }
  *out = &t;
  return CACFE_OK;
}
float* dl_ptr(const DLTensor* t) { return (float*)((char*)t->data + t->byte_offset); }
int64_t dl_numel(const DLTensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->shape[i];
  return n;
}
}  // namespace

extern "C" {

int cacfe_frontend_dlpack(cacfe_plan* p, struct DLManagedTensor* raw_, struct DLManagedTensor* feat_, void* ws, void* stream) {
  if (!p) return fail(CACFE_EINVAL, "frontend_dlpack: null plan");
  const DLTensor *raw, *feat;
  int rc;
  if ((rc = dl_check(p, (const DLManagedTensorABI*)raw_, "frontend_dlpack(raw)", &raw)) != CACFE_OK) return rc;
  if ((rc = dl_check(p, (const DLManagedTensorABI*)feat_, "frontend_dlpack(feat)", &feat)) != CACFE_OK) return rc;
  if (raw->ndim != 2 || raw->shape[1] != p->cfg.n_samples)
    return fail(CACFE_ESHAPE, "frontend_dlpack: raw must be [B][%d]", p->cfg.n_samples);
  const int64_t B = raw->shape[0];
  const int64_t want = B * p->n_frames * p->cfg.n_mels * p->cfg.channels;
  if (dl_numel(feat) != want) return fail(CACFE_ESHAPE, "frontend_dlpack: feat has %lld elements, expected %lld",
                                          (long long)dl_numel(feat), (long long)want);
  return cacfe_frontend(p, dl_ptr(raw), dl_ptr(feat), (int)B, ws, stream);
}

int cacfe_pcen_dlpack(cacfe_plan* p, const cacfe_pcen_params* q, struct DLManagedTensor* in_, struct DLManagedTensor* out_,
                      void* ws, void* stream) {
  if (!p) return fail(CACFE_EINVAL, "pcen_dlpack: null plan");
  const DLTensor *in, *out;
  int rc;
  if ((rc = dl_check(p, (const DLManagedTensorABI*)in_, "pcen_dlpack(in)", &in)) != CACFE_OK) return rc;
  if ((rc = dl_check(p, (const DLManagedTensorABI*)out_, "pcen_dlpack(out)", &out)) != CACFE_OK) return rc;
  if (in->ndim != 3) return fail(CACFE_ESHAPE, "pcen_dlpack: [batch, time, filters] required (tfpcen.py:34)");
  if (dl_numel(in) != dl_numel(out)) return fail(CACFE_ESHAPE, "pcen_dlpack: in/out sizes differ");
  return cacfe_pcen(p, q, dl_ptr(in), dl_ptr(out), (int)in->shape[0], 1, (int)in->shape[1], (int)in->shape[2], ws, stream);
}

}  // extern "C"
