/*
 * salsa_b200 -- C ABI of the B200-native SALSA hot path (libsalsa_b200.so).
 *
 * The reference (thomeou/SALSA) has no FFI layer: its seams are Python callables
 * (SURVEY.md section 8b).  Each entry point below cites the reference callable it replaces;
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - unless a name ends in _host, every pointer is a DEVICE pointer owned by the caller, and
 *     the call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default).
 *   - return value: 0 on success, negative SALSA_E* code on failure; the message of the
 *     last failure on the calling thread is returned by salsa_last_error().
 *   - no hidden device allocations after salsa_workspace_bytes() has been honoured: all
 *     scratch lives in the caller-provided workspace.  (One-time per-device twiddle/window
 *     tables, ~30 KB, are created on first use.)
 *   - thread-safe for distinct streams.
 *
 * Data layouts (row-major, innermost last)
 *   audio     float32  [n_clips][n_chans=4][n_samples]
 *   feature   float32  [n_clips][7][n_frames][feat_dim]       (the h5 'feature' array,
 *                                                               salsa_feature_extraction.py:377-382)
 *   X         float32x2 [n_clips][n_frames][n_chans][n_bins]   complex64 STFT, bins lower..upper-1
 *   power0    float64  [n_clips][n_frames][n_bins]             |X[ch 0]|^2 (tracker input)
 *   mask      uint32   [n_clips][n_frames][(n_bins+31)/32]     bit b of word w = bin 32w+b selected
 *   eig       float32  [n_clips][3][n_frames][n_bins]
 */
#ifndef SALSA_B200_H
#define SALSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SALSA_OK 0
#define SALSA_EINVAL (-1)  /* bad argument (mirrors the reference's ValueError / assert) */
#define SALSA_ECUDA (-2)   /* CUDA runtime error */
#define SALSA_ENOMEM (-3)  /* workspace too small */

#define SALSA_FORMAT_FOA 0 /* audio_format='foa': Re(u[1:]/u[0]) normalised  (:117-120) */
#define SALSA_FORMAT_MIC 1 /* audio_format='mic': angle(u[1:] conj u[0])/(delta*bin) (:121-123) */

#define SALSA_LITE_NIPD 0 /* feature_type='salsa_lite' (salsa_lite_feature_extraction.py:114-115) */
#define SALSA_LITE_IPD 1  /* feature_type='salsa_ipd'  (:112-113) */

/* Parameters of one extraction job; field names follow extract_features()
 * (dataset/salsa_feature_extraction.py:265-311) and the data config yml. */
typedef struct salsa_params {
    int32_t n_clips;
    int32_t n_chans;    /* must be 4 (n_mics, :297) */
    int32_t n_samples;  /* per channel */
    int32_t fs;         /* 24000 */
    int32_t n_fft;      /* 512 (256 is not implemented) */
    int32_t hop_len;    /* 300 */
    int32_t win_len;    /* <= n_fft */
    int32_t lower_bin;  /* first spatial bin, max(1, floor(fmin_doa*n_fft/fs)) (:302-304) */
    int32_t upper_bin;  /* one past the last spatial bin (:303) */
    int32_t audio_format;          /* SALSA_FORMAT_* */
    int32_t is_tracking;           /* noise-floor tracking + coherence test (:89-112) */
    int32_t is_compress_high_freq; /* 200-band log-linear spectrogram (:153-175) */
    int32_t n_hopframes;           /* must be 3 ("do not change", :267) */
    int32_t stft_precision;        /* 64: float64 transform rounded to complex64 as librosa does;
                                      32: float32 transform (faster, ~1e-7 relative) */
    double cond_num;               /* coherence threshold, s[0] > s[1]*cond_num (:106) */
    const double *window;          /* HOST pointer to n_fft float64 window values, or NULL for the
                                      periodic Hann window zero-padded from win_len */
} salsa_params_t;

const char *salsa_last_error(void);
const char *salsa_version(void);

/* Frames of the centred STFT: 1 + n_samples / hop_len (librosa.stft, center=True). */
int32_t salsa_n_frames(int32_t n_samples, int32_t hop_len);
/* Feature width: 200 / n_fft/2 for SALSA (:306-313). */
int32_t salsa_feat_dim(const salsa_params_t *p);

/* ---- op level -------------------------------------------------------------------------------- */

/* librosa.stft for every clip and channel (call sites :186-192, :359-365) fused with
 * MagStftExtractor.extract (:177-201).  Any of X / logspec / power0 may be NULL.
 *   X       complex64, bins lower_bin..upper_bin-1
 *   logspec float32 [n_clips][n_chans][n_frames][feat_dim]  = 10 log10(max(1e-10, W |X|^2))
 *   power0  float64 |X[ch 0]|^2 of the spatial bins (input of salsa_tracker) */
int salsa_stft(const salsa_params_t *p, const float *audio, float *X, float *logspec, double *power0,
               void *stream);

/* Noise-floor tracker of extract_normalized_eigenvector (:26-93): RMS over frames t, t-1, t-2 of
 * |X0|^2 (wrapped), initial floor 0.5*mean of the first 5 frames, up/down tracking in float64,
 * select bins with signal > 1.5 * floor.  n_bins = upper_bin - lower_bin. */
int salsa_tracker(const double *power0, uint32_t *mask, int32_t n_clips, int32_t n_frames, int32_t n_bins,
                  void *stream);

/* Layout / precision adapter for the seam extract_normalized_eigenvector(X, ...) (:17-24): X_ref is
 * one clip of complex128 laid out (n_bins, n_frames, n_chans); writes the internal complex64
 * [n_frames][n_chans][n_bins] array and (if power0 is not NULL) |X_ref[:, :, 0]|^2 in float64. */
int salsa_spectrum_from_reference(const double *X_ref, float *X, double *power0, int32_t n_bins,
                                  int32_t n_frames, int32_t n_chans, void *stream);

/* Covariance over 7 wrapped frames, principal eigenvector, coherence test and FOA / MIC
 * normalisation (:96-127) for the bins selected by `mask` (all bins when mask is NULL, i.e.
 * is_tracking=False).  Output eig [n_clips][3][n_frames][n_bins], zeros where not valid. */
int salsa_eigenvector(const salsa_params_t *p, const float *X, const uint32_t *mask, float *eig,
                      int32_t n_frames, void *stream);

/* ---- clip level ------------------------------------------------------------------------------ */

/* Bytes of device scratch needed by salsa_extract for these parameters: the complex64 spectrum of the spatial bins in
 * tiles of 32 bins ([clip][frame][tile][4][32], 29.5 MB per 60 s FOA clip) and two bit masks.  With the environment
 * variable SALSA_B200_PIPELINE=fused (read at every call) the spectrum stays in shared memory and the scratch holds
 * |X0|^2 instead (7.5 MB per clip); results are the same, the fused arrangement is 1.5x slower. */
size_t salsa_workspace_bytes(const salsa_params_t *p);

/* Per-clip body of extract_features() (dataset/salsa_feature_extraction.py:353-377) for a batch of
 * clips resident in HBM: audio -> feature [n_clips][7][n_frames][feat_dim]. */
int salsa_extract(const salsa_params_t *p, const float *audio, float *feature, void *workspace,
                  size_t workspace_bytes, void *stream);

/* Per-clip body of SALSA-Lite / SALSA-IPD (dataset/salsa_lite_feature_extraction.py:94-123):
 * audio -> feature [n_clips][7][n_frames][cutoff_bin - lower_bin]; upper_bin is applied in cropped
 * coordinates like the reference (:120).  `mode` is SALSA_LITE_*. */
int salsa_lite_extract(const salsa_params_t *p, int32_t cutoff_bin, int32_t mode, const float *audio,
                       float *feature, void *stream);

/* LinSpecIvExtractor.extract (dataset/feature_extraction.py:273-358), the FOA "linspeciv" feature family on the same STFT
 * front-end: audio -> feature [n_clips][7][n_frames][200] = four log-linear spectrogram channels (as salsa_extract) + the
 * three intensity-vector channels W (Re(conj(X0) Xc) / (|IV| + 1e-8)).  lower_bin / upper_bin / audio_format of `p` are
 * ignored; win_len / window apply to both parts, as in the reference.  Scratch: the complex64 spectrum (39 MB per 60 s clip). */
size_t salsa_linspec_iv_workspace_bytes(const salsa_params_t *p);
int salsa_linspec_iv(const salsa_params_t *p, const float *audio, float *feature, void *workspace, size_t workspace_bytes,
                     void *stream);

/* LogSpecGccExtractor.extract (dataset/feature_extraction.py:362-482), the MIC "linspecgcc" family: audio -> feature
 * [n_clips][10][n_frames][200] = four log-linear spectrogram channels + GCC-PHAT of the six channel pairs (sig m, ref n,
 * n < m; the 200 lags -100 .. 99 of the inverse transform of the cross-spectrum phase of a 1024-point STFT).  Needs the
 * compressed layout, win_len = n_fft = 512 and the built-in Hann window.  The workspace (236 MB + 148 MB per 60 s clip) holds the
 * three 512-point spectra that make up the 1024-point one and the GEMM operand of the inverse transform: split large batches. */
size_t salsa_logspec_gcc_workspace_bytes(const salsa_params_t *p);
int salsa_logspec_gcc(const salsa_params_t *p, const float *audio, float *feature, void *workspace, size_t workspace_bytes,
                      void *stream);

/* Same two entry points for HOST buffers (pinned memory recommended): clips are streamed through
 * the GPU in chunks, overlapping host->device copy, kernels and device->host copy.  Synchronous. */
int salsa_extract_host(const salsa_params_t *p, const float *audio_host, float *feature_host,
                       int32_t clips_per_chunk);
int salsa_lite_extract_host(const salsa_params_t *p, int32_t cutoff_bin, int32_t mode,
                            const float *audio_host, float *feature_host, int32_t clips_per_chunk);

/* salsa_extract_host for 16-bit PCM input: audio_host int16 [n_clips][4][n_samples], the samples of the dataset's wav
 * files before librosa.load / soundfile turn them into float32 = sample / 32768 (salsa_feature_extraction.py:353).  Half
 * the host-to-device bytes; the conversion runs on the device and the features are bit-identical to feeding the float32
 * audio. */
int salsa_extract_host_pcm16(const salsa_params_t *p, const int16_t *audio_host, float *feature_host,
                             int32_t clips_per_chunk);

/* The conversion alone, on device buffers (both 16-byte aligned): audio[i] = pcm[i] / 32768. */
int salsa_pcm16_to_float(const int16_t *pcm, float *audio, int64_t n, void *stream);

/* Statistics of compute_scaler() (dataset/salsa_feature_extraction.py:204-262): adds, for spectrogram channels 0..3
 * and every frequency, the sum and the sum of squares over all frames of all clips of `feature`
 * ([n_clips][n_feat_chans][n_frames][feat_dim]) to sums (float64 [4][feat_dim][2], zeroed by the caller before the
 * first batch).  mean = s1 / n, std = sqrt(s2 / n - mean^2) with n = frames seen (StandardScaler's population
 * variance); across GPUs the sums are all-reduced first. */
int salsa_scaler_accumulate(const float *feature, int32_t n_clips, int32_t n_feat_chans, int32_t n_frames,
                            int32_t feat_dim, double *sums, void *stream);

/* The host-buffer entry points keep their streams and device staging buffers between calls (per host thread);
 * this frees them. */
int salsa_host_release(void);

/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
uint64_t salsa_launch_count(int reset);

/* Per-kernel timing with CUDA events recorded on the launching stream around every kernel this
 * library launches from the calling thread.  salsa_profile_read() synchronises, then returns the
 * number of distinct kernels and fills names (32 bytes each), summed milliseconds and launch counts
 * since the previous read. */
int salsa_profile_enable(int on);
int salsa_profile_read(int32_t max_entries, char *names, double *total_ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* SALSA_B200_H */
