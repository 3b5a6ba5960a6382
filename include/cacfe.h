/* cacfe -- B200-native (sm_100a) audio front-end, C ABI.
 *
 * Drop-in boundary for the per-clip feature path of TheCacophonyProject/audio-training.  The reference has
 * no FFI of its own: its "operator interface" is a set of Python callables / Keras layers.  Each entry point
 * below names the reference code it replaces (paths relative to the reference tree); the Python mirror in
 * audio-training_b200/ binds them through ctypes and keeps the reference's names and signatures.
 *
 * Conventions
 *   - every pointer called *_dev is device memory on the plan's device, float32, C-contiguous, and is only
 *     borrowed for the duration of the call; the caller owns all input, output and workspace buffers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous
 *     on it and never synchronise the device unless stated;
 *   - return value: CACFE_OK (0) or a negative cacfe_status; cacfe_last_error() gives the message;
 *   - a plan is immutable after creation and may be shared between threads provided each concurrent call
 *     uses its own workspace and stream;
 *   - there is no CPU fallback: every compute entry fails with CACFE_ECUDA when no sm_100 device is usable.
 */
#ifndef CACFE_H_
#define CACFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CACFE_VERSION 103 /* 0.1.3: + cacfe_ema_init, CACFE_COMPRESS_MEAN_SUB */

typedef enum cacfe_status {
  CACFE_OK = 0,
  CACFE_EINVAL = -1,  /* bad argument / unsupported configuration */
  CACFE_ESHAPE = -2,
  CACFE_EDTYPE = -3,
  CACFE_EDEVICE = -4,
  CACFE_EALIGN = -5,
  CACFE_ECUDA = -6,   /* CUDA runtime / launch failure */
  CACFE_ENOMEM = -7
} cacfe_status;

/* how frames are cut from a clip */
typedef enum cacfe_framing {
  CACFE_FRAME_TF_PAD_END = 0,     /* tf.signal.stft(pad_end=True): T = ceil(N/hop), zeros after the clip (tfdataset.py:2026-2034) */
  CACFE_FRAME_CENTER_ZERO = 1,    /* librosa.stft(center=True, pad_mode="constant"): T = 1 + N/hop  (predict_utils.py:194) */
  CACFE_FRAME_CENTER_REFLECT = 2, /* librosa < 0.10 default pad_mode="reflect" */
  CACFE_FRAME_NO_PAD = 3          /* tf.signal.stft(pad_end=False): T = 1 + (N - n_fft)/hop (tfdataset.py:1824-1832) */
} cacfe_framing;

typedef enum cacfe_layout {
  CACFE_LAYOUT_BMTC = 0, /* [B][n_mels][T][channels]  -- what raw_to_mel / get_spect return (tfdataset.py:2052-2053) */
  CACFE_LAYOUT_BTM = 1   /* [B][T][n_mels]            -- what tfpcen.PCEN consumes (tfpcen.py:34) */
} cacfe_layout;

typedef enum cacfe_mel_impl {
  CACFE_MEL_BANDED_FP32 = 0, /* banded FP32 epilogue (the filterbank is 0.56 % dense) */
  CACFE_MEL_TC_3XTF32 = 1    /* dense tcgen05 3xTF32 GEMM (stored-spectrogram path) */
} cacfe_mel_impl;

typedef struct cacfe_config {
  int32_t sr;          /* 48000 */
  int32_t n_samples;   /* samples per clip, 144000 */
  int32_t n_fft;       /* frame length == FFT length; 4096 (fused kernels); 512, 1024, 2048 ride the 4096-point kernel */
  int32_t hop;         /* 281 */
  int32_t framing;     /* cacfe_framing */
  int32_t n_mels;      /* 160 */
  double fmin;         /* 100 (tfdataset.py:56) or 500 (tfdataset.py:47) */
  double fmax;         /* 11000 */
  double break_freq;   /* 1000; custommel.py:6-8 */
  int32_t power;       /* 2: |X|^2 (raw / inference paths); 1: |X| (stored-spectrogram path, tfdataset.py:1085-1088) */
  int32_t channels;    /* channel replication of the BMTC image: 1 or 3 (tfdataset.py:2053) */
  int32_t out_layout;  /* cacfe_layout */
  int32_t mel_impl;    /* cacfe_mel_impl */
  int32_t normalize;   /* 1: fuse the per-clip min-max normalisation (tfdataset.py:1916-1934) into cacfe_frontend */
  int32_t reserved;
  const float* filterbank; /* optional host [n_mels][n_fft/2+1] row-major bank (e.g. numpy custommel.mel_f output);
                              NULL: built natively from sr/n_mels/fmin/fmax/n_fft/break_freq as custommel.py:18-54 */
} cacfe_config;

/* PCEN layer weights (tfpcen.py:48-87); scope of the final min-max (tfpcen.py:105-110) */
typedef enum cacfe_norm_scope {
  CACFE_NORM_TENSOR = 0, /* reference semantics: min/max over the whole call (batch included) */
  CACFE_NORM_CLIP = 1,   /* per clip */
  CACFE_NORM_NONE = 2    /* no min-max */
} cacfe_norm_scope;

typedef struct cacfe_pcen_params {
  float gain;   /* 0.98, clamped to <= 1 (tfpcen.py:90) */
  float bias;   /* 2.0 */
  float root;   /* 2.0, clamped to >= 1 (tfpcen.py:91) */
  float smooth; /* 0.04, clamped to [0, 1] (tfpcen.py:35) */
  float eps;    /* 1e-6 */
  int32_t norm_scope;
} cacfe_pcen_params;

typedef enum cacfe_compress_mode {
  CACFE_COMPRESS_MAG_POW = 0,     /* badwinner2.MagTransform: x ** param, param = sigmoid(a) (badwinner2.py:47-49) */
  CACFE_COMPRESS_POWER_TO_DB = 1, /* tfdataset.power_to_db (tfdataset.py:1906-1913) */
  CACFE_COMPRESS_MINMAX = 2,      /* normalize_minmax (tfpcen.py:105-110, tfdataset.py:1897-1902) */
  CACFE_COMPRESS_STD = 3,         /* normalize_std (tfdataset.py:1883-1893) */
  CACFE_COMPRESS_MEAN_SUB = 4     /* get_spect(mean_sub=True): x - mean over each entry = one mel row (predict_utils.py:233-236) */
} cacfe_compress_mode;

typedef struct cacfe_plan cacfe_plan;
typedef struct cacfe_hostpipe cacfe_hostpipe;

/* ---- host-only helpers (no GPU needed) ------------------------------------------------------------------ */
int cacfe_version(void);
const char* cacfe_last_error(void); /* message of the last failing call on this thread */
/* custommel.mel_f (custommel.py:18-54): f64 math, f32 result, out = host [n_mels][n_fft/2+1] */
int cacfe_mel_filterbank(int sr, int n_mels, double fmin, double fmax, int n_fft, double break_freq, float* out);
/* frame count for a framing mode (tf.signal.frame / librosa.stft conventions) */
int cacfe_num_frames(int n_samples, int n_fft, int hop, int framing);

/* ---- plan ---------------------------------------------------------------------------------------------- */
int cacfe_plan_create(const cacfe_config* cfg, int device, cacfe_plan** out);
void cacfe_plan_destroy(cacfe_plan* plan);
int cacfe_plan_num_frames(const cacfe_plan* plan);
int cacfe_plan_num_bins(const cacfe_plan* plan);
/* copies the bank the plan uses to host memory [n_mels][n_fft/2+1] */
int cacfe_plan_filterbank(const cacfe_plan* plan, float* out);
/* inclusive range of FFT bins with a non-zero weight (9..938 for the default bank) */
int cacfe_plan_bin_range(const cacfe_plan* plan, int* lo, int* hi);
/* bytes of device workspace the calls below need for a batch of B clips (max over all entry points) */
size_t cacfe_workspace_bytes(const cacfe_plan* plan, int B);
/* workspace of a cacfe_pcen / cacfe_compress call on tensors that are not plan shaped */
size_t cacfe_pcen_workspace_bytes(int B, long long outer_per_clip, int inner);
size_t cacfe_compress_workspace_bytes(long long entries, long long per_entry);
/* kernels launched by this plan so far (for bench accounting) */
long long cacfe_plan_launch_count(const cacfe_plan* plan);
/* enable != 0: bracket every fused STFT/mel kernel launch with CUDA events on its stream (single-threaded use);
 * cacfe_plan_profile_read synchronises them, returns the summed duration and count, and resets the list. */
int cacfe_plan_profile(cacfe_plan* plan, int enable);
/* enable == 1: use the generic (non-TMA, one CTA per tile) form of the fused kernel even where the streaming form
 * applies; both compute the same features (parity tests exercise both).  enable == 2: the streaming kernel without its
 * compile-time specialisations of the common configurations, and the PCEN passes without their root-2 instantiation
 * (bit-identical results; tests).  0: default. */
int cacfe_plan_force_generic(cacfe_plan* plan, int enable);
int cacfe_plan_profile_read(cacfe_plan* plan, double* k1_ms, long long* k1_launches);

/* ---- a1: normalize(input, y) tfdataset.py:1916-1934 / normalize_data predict_utils.py:153-160 ------------
 * rows x n floats, reduction over the last axis, reference operation order with a true f32 division. */
int cacfe_normalize(cacfe_plan* plan, const float* in_dev, float* out_dev, long long rows, long long n,
                    void* workspace_dev, void* stream);

/* ---- a1-a7: raw clips -> mel image.  raw_to_mel tfdataset.py:2007-2059 (framing TF_PAD_END, power 2, channels 3)
 * and get_spect predict_utils.py:163-239 (framing CENTER_*, channels 1).  raw_dev [B][n_samples];
 * feat_dev [B][n_mels][T][channels] or [B][T][n_mels]. */
int cacfe_frontend(cacfe_plan* plan, const float* raw_dev, float* feat_dev, int B, void* workspace_dev, void* stream);

/* ---- next row 1: the stored `audio/spectogram` field.  np.abs(librosa.stft(normed, n_fft, hop)) of
 * audiodataset.load_data (audiodataset.py:1301-1303) with framing CENTER_*, power 1, normalize 1; any framing / power
 * of the plan otherwise.  raw_dev [B][n_samples] -> spec_dev [B][n_fft/2+1][T] (t contiguous, what path C reads). */
size_t cacfe_stft_workspace_bytes(const cacfe_plan* plan, int B);
int cacfe_stft(cacfe_plan* plan, const float* raw_dev, float* spec_dev, int B, void* workspace_dev, void* stream);
/* the same, and (plans with normalize = 1) range_min_dev [B][2] receives each clip's (max - min, min): max - min == 0 is the
 * reference's silent-window test  a_max == a_min  (audiodataset.py:1311-1323), answered by the pass the normalisation
 * needs anyway.  range_min_dev may be NULL. */
int cacfe_stft_stats(cacfe_plan* plan, const float* raw_dev, float* spec_dev, float* range_min_dev, int B,
                     void* workspace_dev, void* stream);

/* ---- a9: stored spectrogram -> mel.  tfdataset.py:1082-1099.  spec_dev [B][n_fft/2+1][T]. */
int cacfe_mel_from_spectrogram(cacfe_plan* plan, const float* spec_dev, float* feat_dev, int B, int T, void* stream);

/* ---- a10: ExponentialMovingAverage.call tfpcen.py:33-39 on x[outer][T][inner], recurrence along T. */
int cacfe_ema(cacfe_plan* plan, float smooth, const float* in_dev, float* out_dev, int B, long long outer_per_clip,
              int T, int inner, void* stream);
/* the same with tf.scan's initializer (tfpcen.py:33-38) given explicitly: init_dev [B * outer_per_clip][inner], the state
 * before the first step; NULL = inputs[:, 0, :], what PCEN.call passes (tfpcen.py:92). */
int cacfe_ema_init(cacfe_plan* plan, float smooth, const float* in_dev, const float* init_dev, float* out_dev, int B,
                   long long outer_per_clip, int T, int inner, void* stream);

/* ---- a10-a12: PCEN.call tfpcen.py:89-99 on x[B * outer_per_clip][T][inner].
 * [B,T,F] (the reference contract): outer_per_clip = 1, inner = F.  in_dev != out_dev. */
int cacfe_pcen(cacfe_plan* plan, const cacfe_pcen_params* params, const float* in_dev, float* out_dev, int B,
               long long outer_per_clip, int T, int inner, void* workspace_dev, void* stream);

/* ---- backward of PCEN.call (tfpcen.py:89-99 incl. normalize_minmax :105-110), what TensorFlow's autodiff
 * computes for that graph.  x_dev, grad_out_dev, grad_x_dev: [B * outer_per_clip][T][inner]; grad_params_dev: device float[4] =
 * dL/d(gain, bias, root, smooth) (0 for a parameter that sits outside its clip range).  T <= 2048. */
size_t cacfe_pcen_backward_workspace_bytes(int B, long long outer_per_clip, int inner);
int cacfe_pcen_backward(cacfe_plan* plan, const cacfe_pcen_params* params, const float* x_dev, const float* grad_out_dev,
                        float* grad_x_dev, float* grad_params_dev, int B, long long outer_per_clip, int T, int inner,
                        void* workspace_dev, void* stream);

/* ---- identifytracks.signal_noise (identifytracks.py:51-143) from the magnitude spectrogram spec_dev[K][T]
 * to the connected components of the signal mask: a_max, row / column medians, threshold (:79-91), open open_size x open_size,
 * dilate dil_h x dil_w, erode ero_h x ero_w (:94-101; OpenCV conventions -- pass 3 x 3 where the reference hands cv2 an empty
 * kernel), 8-connected components with statistics (:106).  Bit-exact against numpy / OpenCV on the same spectrogram.
 * comps_dev: [max_components][6] int32 = (min_x, min_y, max_x, max_y, area, order_key); order_key reproduces OpenCV's label
 * order; n_components_dev: one int32 (may exceed max_components: the list is then truncated).  Optional device outputs (may be
 * NULL): final mask [K][T] uint8, thresholded mask before the morphology, row medians [K], column medians [T]. */
size_t cacfe_signal_workspace_bytes(int K, int T);
int cacfe_signal_components(cacfe_plan* plan, const float* spec_dev, int K, int T, int open_size, int dil_h, int dil_w, int ero_h,
                            int ero_w, unsigned char* mask_dev, unsigned char* raw_mask_dev, float* row_medians_dev,
                            float* col_medians_dev, int32_t* comps_dev, int max_components, int32_t* n_components_dev,
                            void* workspace_dev, void* stream);

/* ---- tfdataset.mix_up tfdataset.py:929-955: out[b] = one[b] * lambda[b] + two[b] * (1 - lambda[b]) over per_entry floats per
 * batch entry (f32, the reference's operation order); lambda_dev: device float[B].  The Beta / Bernoulli draws stay on the host. */
int cacfe_mix_up(cacfe_plan* plan, const float* one_dev, const float* two_dev, const float* lambda_dev, float* out_dev, int B,
                 long long per_entry, void* stream);

/* ---- a12-a14: point-wise compression with a tensor- (entries = 1) or clip-wide (entries = B) statistic. */
int cacfe_compress(cacfe_plan* plan, int mode, float param, const float* in_dev, float* out_dev, long long entries,
                   long long per_entry, void* workspace_dev, void* stream);

/* ---- a15: butter_function / butter_bandpass_filter tfdataset.py:2062-2077 (scipy.signal.sosfilt along the last axis,
 * float64 recurrence, float32 result).  sos_host: host [n_sections][6] as scipy.signal.butter(..., output="sos") returns. */
int cacfe_sosfilt(cacfe_plan* plan, const double* sos_host, int n_sections, const float* in_dev, float* out_dev,
                  long long rows, long long n, void* stream);

/* ---- raw clips -> mel [B][T][n_mels] -> PCEN, device pointers.  normalize + raw_to_mel + PCEN in one call;
 * the mel intermediate lives in the workspace. */
int cacfe_frontend_pcen(cacfe_plan* plan, const cacfe_pcen_params* params, const float* raw_dev, float* out_dev, int B,
                        void* workspace_dev, void* stream);

/* ---- host-buffer pipeline (what a tf.data / numpy caller sees): chunked H2D -> front-end [-> PCEN] -> D2H
 * on two streams.  Owns its device staging (allocated once, here).  max_B bounds the batch of one run. */
int cacfe_hostpipe_create(cacfe_plan* plan, int max_B, int chunk_clips, cacfe_hostpipe** out);
void cacfe_hostpipe_destroy(cacfe_hostpipe* pipe);
/* params == NULL: output is the mel image in the plan's layout; else PCEN output [B][T][n_mels].
 * host_in [B][n_samples], host_out sized accordingly; both should be page-locked (cacfe_host_register).
 * Synchronous: returns when host_out is complete. */
int cacfe_hostpipe_run(cacfe_hostpipe* pipe, const cacfe_pcen_params* params, const float* host_in, float* host_out, int B);
/* The same run from 16-bit PCM host samples [B][n_samples] (extension: the reference's callables take float32 only).  The
 * decoders the reference reads its audio with (librosa.load -> soundfile, audiowriter.py:352,497, audiosplitter.py:20) turn a
 * 16-bit file into float32 as s / 32768; this entry uploads the 16-bit samples (half the bytes of the float32 upload, which is
 * what bounds the host path) and applies that conversion on the device, so the features are bit-identical to
 * cacfe_hostpipe_run on the converted samples.  The staging for it is allocated by the first call. */
int cacfe_hostpipe_run_pcm16(cacfe_hostpipe* pipe, const cacfe_pcen_params* params, const int16_t* host_pcm, float* host_out,
                             int B);
/* Device form of the conversion: out_dev[i] = in_dev[i] / 32768 for n samples. */
int cacfe_pcm16_to_f32(cacfe_plan* plan, const int16_t* in_dev, float* out_dev, long long n, void* stream);
size_t cacfe_hostpipe_device_bytes(const cacfe_hostpipe* pipe);
int cacfe_host_register(void* ptr, size_t bytes);
int cacfe_host_unregister(void* ptr);

/* ---- DLPack front doors: borrow a DLManagedTensor* (never stored, deleter never called), check
 * kDLCUDA / float32 / C-contiguous / device, then forward to the raw-pointer entry points. */
struct DLManagedTensor;
int cacfe_frontend_dlpack(cacfe_plan* plan, struct DLManagedTensor* raw, struct DLManagedTensor* feat,
                          void* workspace_dev, void* stream);
int cacfe_pcen_dlpack(cacfe_plan* plan, const cacfe_pcen_params* params, struct DLManagedTensor* in,
                      struct DLManagedTensor* out, void* workspace_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CACFE_H_ */
