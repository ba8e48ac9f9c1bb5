// SALSA / SALSA-Lite feature kernels for sm_100a.
//
//   stft_kernel          a1 + a3  librosa.stft (+ MagStftExtractor.extract; the spectrum goes to the tiled X of the clip path)
//   tracker_kernel       a4 + a5  noise-floor tracker (float64 recurrence over frames)
//   eig_tile_kernel      a6 - a9  clip path: covariance / eigenvector / coherence / normalisation from TMA-staged tiles of X,
//                        (3, T, F) spatial rows of the feature tensor; eig_redo_kernel = its float64 second opinion
//   eig_kernel           a6 - a8  the same step for the op-level seam extract_normalized_eigenvector(X, ...)
//   salsa_fused_kernel   a1, a3, a6 - a9 in one pass (alternative arrangement): STFT -> shared-memory ring of frames ->
//                        eigenvector step -> (7, T, F) feature rows; X never touches HBM
//   lite_kernel          a10      SALSA-Lite / SALSA-IPD
//   scaler_accumulate_kernel  a11 compute_scaler statistics
//
// (row numbers: SURVEY.md section 8a; reference lines are cited at each function.)
#pragma once
#include "eig.cuh"
#include "fft.cuh"
#include <cuda_bf16.h>

#include "tc_ptx.cuh"

namespace salsa {

#ifndef STFT_MINB
#define STFT_MINB 2                      // resident CTAs per SM stft_kernel is compiled for (register budget 65536 / (256 MINB))
#endif
#ifndef STFT_TAB
#define STFT_TAB 2                       // pass twiddles of the transform from shared-memory tables: bit 0 pass 1, bit 1 pass 2
#endif
#ifndef STFT_PREFETCH
#define STFT_PREFETCH 1                  // stft_kernel requests a warp's next frame while it transforms the current one
#endif
constexpr int kWarps = 8;                // warps per CTA in the STFT-bearing kernels
constexpr int kThreads = kWarps * 32;
constexpr float kAmin = 1e-10f;          // power_to_db amin (salsa_feature_extraction.py:195)
constexpr int kTileBins = 32;            // clip path: X lives in HBM in tiles of 32 bins (see eig_tile_kernel)
constexpr int kTileFrameElems = 4 * kTileBins;       // complex values per (tile, frame)

// Layout of the log-linear bands (MagStftExtractor.__init__, :153-175):
// band < n_lin -> bin band+1 (weight 1); n_lin <= band < n_out -> 8 bins (the last one fewer,
// clipped at n_fft/2) starting at n_lin+1+8*(band-n_lin), weight 1/8.
struct BandLayout {
    int n_lin;
    int n_out;
};

struct StftArgs {
    const float* audio;   // [clip][n_chans][n_samples]
    int n_chans;
    int n_samples;
    int hop;
    int n_frames;
    int lower;            // spatial bins lower..upper-1
    int upper;
    int ch_count;         // channels 0..ch_count-1 are transformed
    int frames_per_block;
    BandLayout bands;
    float2* X;            // [clip][frame][n_chans][x_pitch] or null (bins lower..upper-1 of each row are written)
    int x_pitch;          // row layout: row length of X in complex values (element b of a row = bin lower + b)
    int x_tiles;          // > 0: tiled layout X[clip][frame][tile][channel][32] of the clip path with this many tiles
    float* spec;          // log-linear spectrogram or null
    long long spec_clip_stride;
    long long spec_chan_stride;   // row (frame) stride is bands.n_out
    double* power0;       // [clip][frame][upper-lower] or null
};

template <typename T, int WARPS>
struct FftSmemW {
    T win[kNfft];
    Cx<T> scratch[WARPS][kScratchElems];
    Cx<T> twiddles[kTwiddleTabElems];      // pass twiddles of the TAB variants of the transform (fft.cuh)
};
template <typename T>
using FftSmem = FftSmemW<T, kWarps>;

#ifndef STFT_WARPS
#define STFT_WARPS 8                     // warps per CTA of stft_kernel (a multiple of the channel count)
#endif
constexpr int kStftWarps = STFT_WARPS;
constexpr int kStftThreads = kStftWarps * 32;

template <typename S, typename T>
__device__ __forceinline__ void load_fft_smem(S& s, const FftTables<T>& tb) {
    if (tb.window)
        for (int i = threadIdx.x; i < kNfft; i += blockDim.x) s.win[i] = tb.window[i];
    load_twiddle_tables(s.twiddles, tb);
}

// Warp index as a value the compiler KNOWS is the same in every lane (a shuffle from lane 0): with `threadIdx.x >> 5`
// ptxas treats every loop over a warp's work items as potentially divergent, wraps each __shfl_sync in a WARPSYNC /
// ENDCOLLECTIVE pair and duplicates code (stft_kernel: 2160 -> 1552 SASS instructions with this one line).
__device__ __forceinline__ int uniform_warp_index() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// power_to_db(ref=1, amin=1e-10, top_db=None): 10 log10(max(amin, p)) (:195).  The argument is never
// denormal (>= amin), so the MUFU.LG2 approximation applies directly; its error (<= 2^-22 absolute
// plus 2 ulp) is below 2e-5 dB over the whole [-100, +100] dB range.
__device__ __forceinline__ float power_db(float p) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(kAmin, p)));     // bare MUFU.LG2: no denormal pre-scaling
    return 3.01029995663981195f * l;
}

__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// |X|^2 for the spectrogram.  The reference takes np.abs(complex64)**2 in float32 (:186-194); the
// sum of squares differs from hypot()^2 by about one float32 ulp (5e-7 dB) and can neither overflow
// nor matter below the 1e-10 amin clamp for audio-range spectra.
__device__ __forceinline__ float power_f32(float re, float im) { return fmaf(re, re, im * im); }

// One warp, after the transform: writes the log-linear spectrogram row of this (frame, channel).
// p[j] is the power of bin lane + 32 j, p_nyq the power of the Nyquist bin (only the uncompressed
// layout reaches it: its last band is bin n_fft/2, :172-175).  The compressed layout is the
// n_fft = 512 one: 192 linear bands, then 8 bands of 8 bins (the last of 7) weighted 1/8 (:153-162).
__device__ __forceinline__ void write_logspec_row(const float (&p)[8], float p_nyq, float* row, BandLayout bands,
                                                  int lane) {
    const bool compress = bands.n_out > bands.n_lin;
    if (!compress) {
        // linear layout: band = bin - 1 for bins 1..n_fft/2
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int band = lane + 32 * j - 1;
            if (band >= 0 && band < bands.n_lin) row[band] = power_db(p[j]);
        }
        if (lane == 0) row[kHalf - 1] = power_db(p_nyq);
        return;
    }
    // compressed layout (n_lin = 192): bands 0..191 = bins 1..192, i.e. p[0..5] of every lane but bin 0, and bin 192
#pragma unroll
    for (int j = 0; j < 6; ++j)
        if (j > 0 || lane > 0) row[lane + 32 * j - 1] = power_db(p[j]);
    if (lane == 0) row[191] = power_db(p[6]);
    // r6 / r7[lane] = power of bin 193 + lane / 225 + lane (bin 256 is not part of the last band)
    const float up6 = __shfl_down_sync(0xffffffffu, p[6], 1), up7 = __shfl_down_sync(0xffffffffu, p[7], 1);
    const float first7 = __shfl_sync(0xffffffffu, p[7], 0);
    float r6 = lane == 31 ? first7 : up6;
    float r7 = lane == 31 ? 0.0f : up7;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
        r6 += __shfl_xor_sync(0xffffffffu, r6, m);
        r7 += __shfl_xor_sync(0xffffffffu, r7, m);
    }
    const float v6 = __shfl_sync(0xffffffffu, r6, 8 * (lane & 3)), v7 = __shfl_sync(0xffffffffu, r7, 8 * (lane & 3));
    if (lane < 8) row[bands.n_lin + lane] = power_db(0.125f * (lane < 4 ? v6 : v7));
}

// The same for the layout of warp_fft_passes_paired: p[g] for g < 4 is the power of bin lane + 32 g, for g >= 4 of bin
// 32 g + rl with rl = (32 - lane) & 31 (upper groups in reversed lane order).  In that order the 8 bins of a compressed
// band are the 8 lanes of an aligned octet: bins 193..200 are lanes 31..24 of group 6, ..., bins 217..223 + 224 are lanes
// 7..1 of group 6 + lane 0 of group 7, bins 225..232 lanes 31..24 of group 7, ..., bins 249..255 lanes 7..1.
__device__ __forceinline__ void write_logspec_row_paired(const float (&p)[8], float p_nyq, float* row, BandLayout bands, int lane,
                                                         int rl) {
    const bool compress = bands.n_out > bands.n_lin;
    if (!compress) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int band = (g < 4 ? lane : rl) + 32 * g - 1;
            if (band >= 0) row[band] = power_db(p[g]);
        }
        if (lane == 0) row[kHalf - 1] = power_db(p_nyq);
        return;
    }
#pragma unroll
    for (int g = 0; g < 6; ++g) {
        const int band = (g < 4 ? lane : rl) + 32 * g - 1;
        if (g > 0 || lane > 0) row[band] = power_db(p[g]);
    }
    if (lane == 0) row[191] = power_db(p[6]);           // bin 192
    float r6 = lane == 0 ? p[7] : p[6];                  // lane 0: bin 224 closes the band of bins 217..224
    float r7 = lane == 0 ? 0.0f : p[7];                  // the last band has 7 bins (249..255)
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
        r6 += __shfl_xor_sync(0xffffffffu, r6, m);
        r7 += __shfl_xor_sync(0xffffffffu, r7, m);
    }
    // band 192 + i (i = 0..3) = octet 3 - i of r6, band 196 + i = octet 3 - i of r7
    const int src = 8 * (3 - (lane & 3));
    const float v6 = __shfl_sync(0xffffffffu, r6, src), v7 = __shfl_sync(0xffffffffu, r7, src);
    if (lane < 8) row[bands.n_lin + lane] = power_db(0.125f * (lane < 4 ? v6 : v7));
}

// ------------------------------------------------------------------------------------------------
// stft_kernel: grid (frame blocks, clips); one warp per (frame, channel) item.
// ------------------------------------------------------------------------------------------------
// OUT: which outputs exist, known at compile time on the clip path so that no pointer is tested per bin
// (kStftAny: test the pointers of StftArgs at run time -- the op-level entry point).
constexpr int kStftX = 1, kStftPower0 = 2, kStftSpec = 4, kStftAny = -1;
// kStftTiled: X goes to the tiled layout of the clip path (lower_bin < 32); kStftTiledAny: same for any lower_bin
constexpr int kStftTiled = 8, kStftTiledAny = 16;

template <typename T, int CH, int OUT>
__global__ void __launch_bounds__(kStftThreads, STFT_MINB) stft_kernel(StftArgs a, FftTables<T> tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = FftSmemW<T, kStftWarps>;
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    load_fft_smem(s, tb);
    __syncthreads();
    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    LaneTwiddles<T> tw = lane_twiddles(tb, lane);
    use_twiddle_tables(tw, s.twiddles, lane);
    const int clip = blockIdx.y;
    const int f0 = blockIdx.x * a.frames_per_block;
    const int f1 = min(a.n_frames, f0 + a.frames_per_block);
    const int nb = a.upper - a.lower;
    const bool has_x = OUT == kStftAny ? a.X != nullptr : (OUT & kStftX) != 0;
    const bool has_p0 = OUT == kStftAny ? a.power0 != nullptr : (OUT & kStftPower0) != 0;
    const bool has_spec = OUT == kStftAny ? a.spec != nullptr : (OUT & kStftSpec) != 0;
    constexpr bool tiled = OUT != kStftAny && (OUT & (kStftTiled | kStftTiledAny)) != 0;
    // a warp keeps its channel: items (frame, channel) are dealt frame-major, so warp w walks the frames
    // f0 + w / CH, + kWarps / CH, ... of channel w % CH
    static_assert(kStftWarps % CH == 0, "channels divide the warps of a CTA");
    constexpr int kFrameStep = kStftWarps / CH;
    const int ch = warp % CH;
    const float* chan_audio = a.audio + ((long long)clip * a.n_chans + ch) * a.n_samples;
    Cx<T>* scratch = s.scratch[warp];
    const T* win = tb.window ? s.win : nullptr;
    const int rl = (32 - lane) & 31;             // position of this lane's bins in the upper groups (reversed order)
    const int lo_off = lane - a.lower + (lane < a.lower ? kTileBins - kTileFrameElems : 0);
    const int hi_off = rl - a.lower + (rl < a.lower ? kTileBins - kTileFrameElems : 0);
    float2 raw[8];
    int t = f0 + warp / CH;
    if (STFT_PREFETCH && t < f1) load_frame(chan_audio, a.n_samples, t * a.hop - kNfft / 2, lane, raw);
    for (; t < f1; t += kFrameStep) {
        Cx<T> v[8];
        if (!STFT_PREFETCH) load_frame(chan_audio, a.n_samples, t * a.hop - kNfft / 2, lane, raw);
        window_frame<T>(raw, win, tw, lane, v);
        // `raw` is consumed: the samples of this warp's next frame are requested before the passes
        if (STFT_PREFETCH && t + kFrameStep < f1) load_frame(chan_audio, a.n_samples, (t + kFrameStep) * a.hop - kNfft / 2, lane, raw);
        Cx<T> lo[4], hr[4], x128;
        warp_fft_passes_paired<T, STFT_TAB>(v, tw, scratch, lane, lo, hr, x128);
        // librosa stores complex64: round, then put the upper half into groups (reversed lane order, fft.cuh)
        float2 hf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) hf[j] = make_float2((float)hr[j].re, (float)hr[j].im);
        const float2 x128f = make_float2((float)x128.re, (float)x128.im);
        float2 xf[8];
#pragma unroll
        for (int g = 0; g < 4; ++g) xf[g] = make_float2((float)lo[g].re, (float)lo[g].im);
#pragma unroll
        for (int g = 4; g < 8; ++g) xf[g] = upper_group(hf, x128f, g, lane);
        const long long o = ((long long)clip * a.n_frames + t);
        float2* xrow = a.X + (o * a.n_chans + ch) * a.x_pitch - a.lower;
        double* prow = a.power0 + o * nb - a.lower;
        float2* xtile = a.X + o * a.x_tiles * kTileFrameElems + ch * kTileBins;
        float p[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int k = 32 * g + (g < 4 ? lane : rl);
            const float re = xf[g].x, im = xf[g].y;
            p[g] = power_f32(re, im);
            if (tiled && (unsigned)(k - a.lower) < (unsigned)nb) {
                // tiled layout: spatial bin b = k - lower -> tile b / 32, position b % 32.  With lower < 32 that is tile g
                // (g - 1 for the `lower` lanes at the bottom of the group) at a per-lane offset that does not depend on g
                if (OUT & kStftTiled) xtile[g * kTileFrameElems + (g < 4 ? lo_off : hi_off)] = xf[g];
                else xtile[((k - a.lower) >> 5) * kTileFrameElems + ((k - a.lower) & 31)] = xf[g];
            }
            if (!tiled && (has_x || has_p0) && (unsigned)(k - a.lower) < (unsigned)nb) {
                if (has_x) xrow[k] = xf[g];
                // np.abs(complex128) ** 2 (:53-55) up to one float64 ulp
                if (has_p0 && ch == 0) prow[k] = fma((double)re, (double)re, (double)im * (double)im);
            }
        }
        if (!tiled && has_x && a.upper > kHalf && lane == 0) xrow[kHalf] = make_float2(hf[0].x, 0.0f);     // the Nyquist bin (real)
        if (has_spec) {
            const float p_nyq = hf[0].x * hf[0].x;        // lane 0: hr[0] = X[256], real
            float* row = a.spec + clip * a.spec_clip_stride + ch * a.spec_chan_stride + (long long)t * a.bands.n_out;
            write_logspec_row_paired(p, p_nyq, row, a.bands, lane, rl);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// stft256_kernel: the n_fft = 256 configuration the reference also accepts (salsa_feature_extraction.py:151-152, :163-170,
// :300-306).  No shipped config uses it, so this is the plain form of stft_kernel: one warp per (frame, channel), the 256 real
// samples go through the 256-point COMPLEX transform of fft.cuh with zero imaginary parts (Z[k] is then the real
// transform's bin k for k <= 128; no split step), outputs selected by the pointers of StftArgs.  The window always comes
// from the table (entries 0..255 of FftTables::window).
// ------------------------------------------------------------------------------------------------
constexpr int kNfft256 = 256;

// log-linear row of the n_fft = 256 layouts: p[g] = power of bin lane + 32 g (g = 0..3), p_nyq = power of bin 128.
// Compressed (:163-170): bands 0..95 = bins 1..96, bands 96..98 = 8 bins from 97 + 8 i, band 99 = bins 121..127, all / 8.
__device__ __forceinline__ void write_logspec_row256(const float (&p)[4], float p_nyq, float* row, BandLayout bands, int lane) {
    const bool compress = bands.n_out > bands.n_lin;
    if (!compress) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int band = lane + 32 * g - 1;
            if (band >= 0) row[band] = power_db(p[g]);
        }
        if (lane == 0) row[kNfft256 / 2 - 1] = power_db(p_nyq);
        return;
    }
#pragma unroll
    for (int g = 0; g < 3; ++g)
        if (g > 0 || lane > 0) row[lane + 32 * g - 1] = power_db(p[g]);
    if (lane == 0) row[95] = power_db(p[3]);                        // bin 96
    const float up = __shfl_down_sync(0xffffffffu, p[3], 1);         // lane l: power of bin 97 + l
    float r = lane == 31 ? 0.0f : up;                                // the last band has 7 bins (121..127)
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) r += __shfl_xor_sync(0xffffffffu, r, m);
    const float v = __shfl_sync(0xffffffffu, r, 8 * (lane & 3));
    if (lane < 4) row[bands.n_lin + lane] = power_db(0.125f * v);
}

template <typename T>
__global__ void __launch_bounds__(kStftThreads) stft256_kernel(StftArgs a, FftTables<T> tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = FftSmemW<T, kStftWarps>;
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    load_fft_smem(s, tb);
    __syncthreads();
    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    const LaneTwiddles<T> tw = lane_twiddles(tb, lane);
    const int clip = blockIdx.y;
    const int f0 = blockIdx.x * a.frames_per_block;
    const int f1 = min(a.n_frames, f0 + a.frames_per_block);
    const int nb = a.upper - a.lower;
    Cx<T>* scratch = s.scratch[warp];
    for (int item = warp; item < (f1 - f0) * a.ch_count; item += kStftWarps) {
        const int t = f0 + item / a.ch_count, ch = item % a.ch_count;
        const float* x = a.audio + ((long long)clip * a.n_chans + ch) * a.n_samples;
        const int start = t * a.hop - kNfft256 / 2;
        Cx<T> v[8], z[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int m = lane + 32 * n1;
            v[n1] = {(T)x[reflect_index(start + m, a.n_samples)] * s.win[m], (T)0};
        }
        warp_fft_core<T, 0>(v, tw, scratch, lane, z);               // z[i] = Z[lane + 32 i]
        const long long o = ((long long)clip * a.n_frames + t);
        float p[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int k = lane + 32 * g;
            const float2 xf = make_float2((float)z[g].re, (float)z[g].im);     // librosa stores complex64
            p[g] = power_f32(xf.x, xf.y);
            if ((unsigned)(k - a.lower) < (unsigned)nb) {
                const int b = k - a.lower;
                if (a.X && a.x_tiles > 0) a.X[o * a.x_tiles * kTileFrameElems + (b >> 5) * kTileFrameElems + ch * kTileBins + (b & 31)] = xf;
                else if (a.X) a.X[(o * a.n_chans + ch) * a.x_pitch + b] = xf;
                if (a.power0 && ch == 0) a.power0[o * nb + b] = fma((double)xf.x, (double)xf.x, (double)xf.y * (double)xf.y);
            }
        }
        if (a.spec) {
            const float nyq = (float)z[4].re;                       // lane 0: Z[128], real for a real signal
            float* row = a.spec + clip * a.spec_clip_stride + ch * a.spec_chan_stride + (long long)t * a.bands.n_out;
            write_logspec_row256(p, nyq * nyq, row, a.bands, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tracker_kernel: one thread per (clip, bin), one warp per 32-bin mask word; sequential over frames.
// Reference: salsa_feature_extraction.py:26-36 (constants), :49-58 (signal, initial floor), :63-87.
// ------------------------------------------------------------------------------------------------
struct TrackerConsts {
    double floor_up, floor_up_slow, floor_down, snr_ratio, floor_min;
    int n_sig_frames, n_init_frames;
};

// Input of the tracker: |X0|^2 in float64, either precomputed (power0, [clip][frame][n_bins]) or taken from channel 0
// of the clip path's tiled complex64 spectrum -- the same fma as stft_kernel's power0, so both give identical bits.
struct TrackerPower0 {
    using Raw = double;
    const double* p;
    long long clip_stride, frame_stride;
    __device__ __forceinline__ Raw load(int clip, int t, int b) const { return p[clip * clip_stride + t * frame_stride + b]; }
    static __device__ __forceinline__ double power(Raw r) { return r; }
};

// clip path: channel 0 of the tiled spectrum
struct TrackerTiles {
    using Raw = float2;
    const float2* x;
    int n_tiles, n_frames;
    __device__ __forceinline__ Raw load(int clip, int t, int b) const {
        return __ldg(x + (((long long)clip * n_frames + t) * n_tiles + (b >> 5)) * kTileFrameElems + (b & 31));
    }
    static __device__ __forceinline__ double power(Raw v) { return fma((double)v.x, (double)v.x, (double)v.y * (double)v.y); }
};

// bins first_bin .. n_bins - 1 are tracked (bits of the bins below first_bin stay clear).
// The kernel is one dependent chain of 4801 steps per warp, so its duration hardly depends on the batch (1.6 ms for 32
// clips, 2.1 ms for 600 with the loads one chunk ahead): the loads run THREE chunks of 8 frames ahead of the recurrence
// and stay raw (complex64) until their chunk is processed, so that nothing in program order waits for them.
// Running the tracker on a side stream beside the transform of channels 1..3 was tried and gains nothing: stft_kernel
// holds the whole register file, the tracker's CTAs only get in as it drains.
template <typename Src>
__global__ void __launch_bounds__(256) tracker_kernel(Src src, uint32_t* __restrict__ mask, int n_frames, int first_bin, int n_bins,
                                                      TrackerConsts c) {
    using Raw = typename Src::Raw;
    const int clip = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int word = b >> 5, n_words = (n_bins + 31) >> 5;
    const bool live = b >= first_bin && b < n_bins;
    const int bb = live ? b : first_bin;
    uint32_t* mrow = mask + (long long)clip * n_frames * n_words + word;
    auto at = [&](int t) -> double {   // wrapped frame access (np.pad 'wrap', :43)
        t %= n_frames;
        if (t < 0) t += n_frames;
        return live ? Src::power(src.load(clip, t, bb)) : 0.0;
    };
    // initial floor: 0.5 * mean(sig[0:5]), sig = sqrt(mean of three powers)   (:53-58)
    const int n_init = min(c.n_init_frames, n_frames);
    double acc = 0.0;
    for (int t = 0; t < n_init; ++t) acc += sqrt(((at(t) + at(t - 1)) + at(t - 2)) / 3.0);
    double nf = 0.5 * (acc / (double)n_init);
    int cd = c.n_sig_frames;
    double a1 = at(-1), a2 = at(-2);
    // The frame loop compares SQUARES: sig > floor  <=>  sig^2 = (a0+a1+a2)/3 > floor^2.  That removes
    // the float64 sqrt and division (about 90 % of this kernel's float64 work) and moves the two
    // comparisons by a few 1e-16 relative, far inside what hypot()'s last bit already leaves open.
    const double third = 1.0 / 3.0;
    const double snr2 = c.snr_ratio * c.snr_ratio;
    constexpr int kChunk = 8;
    auto fetch = [&](Raw (&r)[kChunk], int t0) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) r[i] = src.load(clip, min(t0 + i, n_frames - 1), bb);     // clamped: always a valid address
    };
    auto process = [&](const Raw (&r)[kChunk], int t0) {
        double pw[kChunk];
#pragma unroll
        for (int i = 0; i < kChunk; ++i) pw[i] = live ? Src::power(r[i]) : 0.0;
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            if (t0 + i < n_frames) {
                const double a0 = pw[i];
                const double q = ((a0 + a1) + a2) * third;
                a2 = a1;
                a1 = a0;
                const bool above = q > nf * nf;
                if (above) {
                    cd -= 1;
                    nf *= (cd < 0) ? c.floor_up_slow : c.floor_up;
                } else {
                    cd = c.n_sig_frames;
                    nf *= c.floor_down;
                }
                if (nf < c.floor_min) nf = c.floor_min;
                const bool sel = live && (q > snr2 * (nf * nf));
                const uint32_t bits = __ballot_sync(0xffffffffu, sel);
                // `word < n_words`: a block's surplus warps (n_words not a multiple of the warps per block) take part in
                // nothing but the ballot
                if ((threadIdx.x & 31) == 0 && word < n_words) mrow[(long long)(t0 + i) * n_words] = bits;
            }
        }
    };
    Raw r0[kChunk], r1[kChunk], r2[kChunk];
    fetch(r0, 0);
    fetch(r1, kChunk);
    fetch(r2, 2 * kChunk);
    for (int t0 = 0; t0 < n_frames; t0 += 3 * kChunk) {
        process(r0, t0);
        fetch(r0, t0 + 3 * kChunk);
        process(r1, t0 + kChunk);
        fetch(r1, t0 + 4 * kChunk);
        process(r2, t0 + 2 * kChunk);
        fetch(r2, t0 + 5 * kChunk);
    }
}

// ------------------------------------------------------------------------------------------------
// from_reference_kernel: the reference hands extract_normalized_eigenvector a complex128 array laid
// out (n_bins, n_frames, n_chans) (:20); convert one clip of it to the internal complex64
// [frame][ch][bin] layout and emit |X[:, :, 0]|^2 in float64 for the tracker.
// ------------------------------------------------------------------------------------------------
__global__ void from_reference_kernel(const double2* __restrict__ Xref, float2* __restrict__ X,
                                      double* __restrict__ power0, int n_bins, int n_frames) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_bins * n_frames) return;
    const int t = (int)(i / n_bins), b = (int)(i % n_bins);
    const double2* src = Xref + ((long long)b * n_frames + t) * 4;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const double2 v = src[ch];
        X[((long long)t * 4 + ch) * n_bins + b] = make_float2((float)v.x, (float)v.y);
        if (ch == 0 && power0) {
            const double m = hypot(v.x, v.y);
            power0[(long long)t * n_bins + b] = m * m;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// eig_kernel: X resident in HBM ([clip][frame][ch][bin] complex64); thread per (frame, bin).
// Used by the op-level seam extract_normalized_eigenvector (arbitrary caller-provided X).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eig_kernel(const float2* __restrict__ X, const uint32_t* __restrict__ mask,
                                                  float* __restrict__ out, int n_frames, int n_bins, EigArgs e) {
    const int clip = blockIdx.z;
    const int t = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bins) return;
    const int n_words = (n_bins + 31) >> 5;
    bool sel = true;
    if (mask) sel = (mask[((long long)clip * n_frames + t) * n_words + (b >> 5)] >> (b & 31)) & 1u;
    float o[3] = {0.0f, 0.0f, 0.0f};
    if (sel) {
        const float2* base = X + (long long)clip * n_frames * 4 * n_bins + b;
        const float2* fp[kWin];
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            int tt = (t - kHop + k) % n_frames;              // wrap padding of the frame axis (:43)
            if (tt < 0) tt += n_frames;
            fp[k] = base + (long long)tt * 4 * n_bins;
        }
        auto load = [&](int k, int ch) -> float2 { return __ldg(fp[k] + ch * n_bins); };
        eig_bin(load, e, b, o);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) out[(((long long)clip * 3 + i) * n_frames + t) * n_bins + b] = o[i];
}

// ------------------------------------------------------------------------------------------------
// Clip path, split arrangement: the spectrum X lives in HBM in TILES of 32 spatial bins (b = bin - lower_bin),
//     X[clip][frame][tile = b / 32][channel][b % 32]   (complex64; tiles 0 .. (n_bins - 1) / 32)
// so that the 4 channels x 32 bins of one (frame, tile) are 1 KB contiguous -- one TMA bulk copy per frame of an
// eig_tile_kernel CTA -- while stft_kernel still writes all rows of a frame inside one 6 KB block.  A tile is also one
// word of the tracker mask and one aligned 128-byte segment of a feature row.
// ------------------------------------------------------------------------------------------------
// eig_tile_kernel: the eigenvector step.  grid (frame tiles of FT, bin tiles, clips); 256 threads.
//   1  the FT + 6 frames of this bin tile are brought into shared memory by TMA bulk copies (1 KB per frame, one
//      issuing thread per frame, completion on one mbarrier; the wrap padding of the frame axis, :43, is just the
//      source address).  Every X element is then read 7 times from shared memory.  Meanwhile the tracker mask words of
//      the tile are compacted into a dense list of selected (frame, bin) items, so that the eigenvector step runs with
//      full warps whatever the selection looks like;
//   2  one thread per list item: covariance over 7 frames (28 LDS.64 at immediate offsets), float32 eigenvector
//      (FFMA2 complex arithmetic), certified coherence test, normalisation -> staging tile.  Bins whose float32 verdict
//      cannot be certified are marked in `redo` for eig_redo_kernel;
//   3  the 3 x FT row segments go to HBM, zeros where the bin was not selected / not valid; the last bin tile also
//      zero-fills the columns above the last spatial bin (:373-374).
// No float64 code in this kernel: 64 registers, 4 CTAs (32 warps) per SM.  (Walking several frame tiles per CTA with the
// next tile's copy issued behind the row output was measured: no gain, 10.3 vs 10.0 ms -- with 4 CTAs per SM the copy
// latency is already covered by the other CTAs.)
// ------------------------------------------------------------------------------------------------
struct EigTileArgs {
    const float2* X;         // tiled spectrum (see above)
    const uint32_t* mask;    // [clip][n_frames][n_tiles] tracker selection, or null (is_tracking = false)
    uint32_t* redo;          // [clip][n_frames][n_tiles], every word is written
    float* feature;          // [clip][7][n_frames][feat_dim] (feat_dim a multiple of 4); channels 4..6 are written
    int n_frames, n_bins, n_tiles, feat_dim;
    EigArgs eig;
};

template <int FT>
__host__ __device__ constexpr size_t eig_tile_smem_bytes() {
    return (size_t)(FT + 2 * kHop) * kTileFrameElems * sizeof(float2) + (size_t)3 * FT * kTileBins * sizeof(float) +
           (size_t)FT * kTileBins * sizeof(uint16_t) + 2 * FT * sizeof(uint32_t) + 16;
}

// One tile: frames t0 .. t0 + nt - 1 (nt <= FT) of bin tile bt of clip `clip`; `phase` is the parity of the mbarrier
// phase this copy completes.
template <int FT, int NSQ, int WARPS>
__device__ __forceinline__ void eig_tile_body(const EigTileArgs& a, unsigned char* smem_raw, int clip, int bt, int t0, int nt,
                                              uint32_t phase, int warp, int lane) {
    constexpr int BB = kTileBins, R = FT + 2 * kHop, NT = WARPS * 32;
    float2* xs = reinterpret_cast<float2*>(smem_raw);                                     // [R][4][BB]
    float* stage = reinterpret_cast<float*>(xs + R * kTileFrameElems);                    // [3][FT][BB]
    uint16_t* list = reinterpret_cast<uint16_t*>(stage + 3 * FT * BB);                    // [FT * BB]
    uint32_t* smask = reinterpret_cast<uint32_t*>(list + FT * BB);                        // [FT]
    uint32_t* sredo = smask + FT;                                                          // [FT]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sredo + FT);
    int* n_items = reinterpret_cast<int*>(bar + 1);
    if (threadIdx.x < FT) sredo[threadIdx.x] = 0u;
    // ---- 1
    const int n_rows = nt + 2 * kHop;
    // (the phase cannot complete before thread 0's arrival, whatever the order in which the copies below land)
    if (threadIdx.x == 0) tc::mbar_expect_tx(bar, (uint32_t)(n_rows * kTileFrameElems * sizeof(float2)));
    if (threadIdx.x < n_rows) {
        int f = (t0 - kHop + (int)threadIdx.x) % a.n_frames;             // wrap padding of the frame axis (:43)
        if (f < 0) f += a.n_frames;
        const float2* src = a.X + (((long long)clip * a.n_frames + f) * a.n_tiles + bt) * kTileFrameElems;
        tc::bulk_load_1d(xs + threadIdx.x * kTileFrameElems, src, kTileFrameElems * sizeof(float2), bar);
    }
    {
        // every warp scans the popcounts of the tile's 32 mask words (lane = frame) and expands its own frames at the
        // offsets the scan gives: no atomics, and the list is sorted by (frame, bin)
        uint32_t bits = 0u;
        if (lane < nt) {
            bits = a.mask ? __ldg(a.mask + ((long long)clip * a.n_frames + t0 + lane) * a.n_tiles + bt) : 0xffffffffu;
            if (bt == a.n_tiles - 1 && (a.n_bins & 31)) bits &= (1u << (a.n_bins & 31)) - 1u;
        }
        int incl = __popc(bits);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (warp == 0) {
            smask[lane] = bits;
            if (lane == 31) *n_items = incl;
        }
        const int excl = incl - __popc(bits);
#pragma unroll
        for (int q = 0; q < FT / WARPS; ++q) {
            const int tl = warp + WARPS * q;                 // frames of this warp
            const uint32_t w = __shfl_sync(0xffffffffu, bits, tl);
            const int base = __shfl_sync(0xffffffffu, excl, tl);
            if ((w >> lane) & 1u) list[base + __popc(w & ((1u << lane) - 1u))] = (uint16_t)((tl << 5) | lane);
        }
    }
    if (warp == 0) tc::mbar_wait(bar, phase);      // one warp polls the mbarrier, the others sleep at the barrier below
    __syncthreads();
    // ---- 2
    const int count = *n_items;
    const uint32_t xs_addr = (uint32_t)__cvta_generic_to_shared(xs);
    for (int item = threadIdx.x; item < count; item += NT) {
        const int code = list[item];
        const int tl = code >> 5, bl = code & 31;
        const uint32_t base = xs_addr + (uint32_t)((tl * kTileFrameElems + bl) * sizeof(float2));   // X[t - 3][ch 0][bin]
        float o[3];
        auto load = [&](int k, int ch) -> float2 { return lds_f2(base + (uint32_t)((k * 4 + ch) * BB * sizeof(float2))); };
        const int verdict = eig_bin_f32<NSQ>(load, a.eig, bt * BB + bl, o);
        if (verdict == kEigAmbiguous) atomicOr(&sredo[tl], 1u << bl);
#pragma unroll
        for (int i = 0; i < 3; ++i) stage[(i * FT + tl) * BB + bl] = o[i];
    }
    __syncthreads();
    // ---- 3
    if (threadIdx.x < nt) a.redo[((long long)clip * a.n_frames + t0 + threadIdx.x) * a.n_tiles + bt] = sredo[threadIdx.x];
    const long long chan_stride = (long long)a.n_frames * a.feat_dim;
    float* clip_feat = a.feature + (long long)clip * 7 * chan_stride;
    const int c0 = bt * BB;
    // row segment s = channel * FT + frame; a warp writes four 128-byte segments per step (8 lanes x float4 each)
    for (int s = (threadIdx.x >> 3); s < 3 * FT; s += NT / 8) {
        const int i = s / FT, tl = s % FT;
        if (tl >= nt) continue;
        const int k = (threadIdx.x & 7) * 4;
        if (c0 + k >= a.feat_dim) continue;        // the last bin tile may reach past the feature row (n_bins <= feat_dim < 32 n_tiles)
        const uint32_t bits = smask[tl] >> k;
        const float4 sv = *reinterpret_cast<const float4*>(stage + s * BB + k);
        float4 v;
        v.x = (bits & 1u) ? sv.x : 0.0f;
        v.y = (bits & 2u) ? sv.y : 0.0f;
        v.z = (bits & 4u) ? sv.z : 0.0f;
        v.w = (bits & 8u) ? sv.w : 0.0f;
        *reinterpret_cast<float4*>(clip_feat + (4 + i) * chan_stride + (long long)(t0 + tl) * a.feat_dim + c0 + k) = v;
    }
    if (bt == a.n_tiles - 1) {
        // the last bin tile zero-fills the columns above the last spatial bin (:373-374)
        const int extra = (a.feat_dim - c0 - BB) >> 2;               // float4 per row
        for (int g = threadIdx.x; g < 3 * nt * extra; g += NT) {
            const int r = g / extra, q = g - r * extra;
            const int i = r / nt, tl = r - i * nt;
            *reinterpret_cast<float4*>(clip_feat + (4 + i) * chan_stride + (long long)(t0 + tl) * a.feat_dim + c0 + BB + 4 * q) =
                make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
}

// grid (frame tiles of FT, bin tiles, clips)
template <int FT, int MINB, int NSQ, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, MINB) eig_tile_kernel(EigTileArgs a) {
    static_assert(FT % WARPS == 0 && WARPS * 32 >= FT + 2 * kHop, "frames dealt to the warps; one copy-issuing thread per frame");
    static_assert(FT == 32, "one mask word per lane in the compaction scan (24 frames per tile measured 7 % slower)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + eig_tile_smem_bytes<FT>() - 16);
    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int t0 = blockIdx.x * FT;
    eig_tile_body<FT, NSQ, WARPS>(a, smem_raw, blockIdx.z, blockIdx.y, t0, min(FT, a.n_frames - t0), 0u, warp, lane);
}

// eig_redo_kernel: float64 re-evaluation of the bins eig_tile_kernel marked (a few per 100 000); one thread per
// mask word, valid results overwrite the zeros eig_tile_kernel left in the feature rows.
__global__ void __launch_bounds__(128) eig_redo_kernel(EigTileArgs a, long long n_words_total) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words_total) return;
    uint32_t bits = __ldg(a.redo + w);
    if (!bits) return;
    const int bt = (int)(w % a.n_tiles);
    const long long ft = w / a.n_tiles;
    const int t = (int)(ft % a.n_frames), clip = (int)(ft / a.n_frames);
    const float2* tile_x = a.X + ((long long)clip * a.n_frames * a.n_tiles + bt) * kTileFrameElems;
    const long long chan_stride = (long long)a.n_frames * a.feat_dim;
    float* clip_feat = a.feature + (long long)clip * 7 * chan_stride;
    while (bits) {
        const int bl = __ffs(bits) - 1;
        bits &= bits - 1;
        const float2* fp[kWin];
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            int tt = (t - kHop + k) % a.n_frames;
            if (tt < 0) tt += a.n_frames;
            fp[k] = tile_x + (long long)tt * a.n_tiles * kTileFrameElems + bl;
        }
        auto load = [&](int k, int ch) -> float2 { return __ldg(fp[k] + ch * kTileBins); };
        const int b = bt * kTileBins + bl;                    // spatial bin = feature column
        float o[3] = {0.0f, 0.0f, 0.0f};
        if (eig_bin_f64(load, a.eig, o, b) == kEigPass) {
#pragma unroll
            for (int i = 0; i < 3; ++i) clip_feat[(4 + i) * chan_stride + (long long)t * a.feat_dim + b] = o[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// salsa_fused_kernel: grid (segments, clips).  A CTA walks `seg_len` consecutive frames of one clip
// in steps of FT frames.  Per step:
//   phase 1  FT*4 warp-FFTs (new frames only) into a ring of FT+6 frames of X in shared memory
//            ([slot][ch][bin] complex64); log-spectrogram rows go straight to HBM.  Meanwhile the
//            tracker mask words of the step are compacted into a dense list of selected (frame, bin)
//            items, so that phase 2 runs with full warps whatever the selection looks like.
//   phase 2  one thread per list item: covariance over 7 ring frames, eigenvector, coherence test,
//            normalisation -> staging tile in shared memory (aliases the FFT scratch).
//   phase 3  the 3 x FT spatial rows are written to HBM as whole rows (float4), zeros where the
//            bin was not selected / not valid / above the last spatial bin (:373-374).
// HBM traffic per clip = audio once (+ halo re-reads served by L2) + feature once + mask bits.
// ------------------------------------------------------------------------------------------------
struct FusedArgs {
    const float* audio;
    float* feature;          // [clip][7][n_frames][feat_dim]
    const uint32_t* mask;    // tracker selection or null (is_tracking = false)
    int n_samples;
    int hop;
    int n_frames;
    int lower, upper;
    int nbp;                 // ring row length in bins (n_bins rounded up to 32, <= 256)
    int seg_len;
    BandLayout bands;
    EigArgs eig;
};

template <int FT>
__host__ __device__ constexpr int fused_max_words() { return FT * 8; }

// shared memory after FftSmem<T>: ring, item list, mask words, item count
template <typename T, int FT>
__host__ __device__ inline size_t fused_smem_bytes(int nbp) {
    return sizeof(FftSmem<T>) + (size_t)(FT + 2 * kHop) * 4 * nbp * sizeof(float2) + (size_t)FT * nbp * sizeof(uint16_t) +
           fused_max_words<FT>() * sizeof(uint32_t) + 16;
}

template <typename T, int FT>
__global__ void __launch_bounds__(kThreads, 2) salsa_fused_kernel(FusedArgs a, FftTables<T> tb) {
    constexpr int R = FT + 2 * kHop;
    static_assert(sizeof(Cx<T>) * kScratchElems * kWarps >= 3 * FT * 256 * sizeof(float), "staging tile must fit in the FFT scratch");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FftSmem<T>& s = *reinterpret_cast<FftSmem<T>*>(smem_raw);
    float2* ring = reinterpret_cast<float2*>(smem_raw + sizeof(FftSmem<T>));                 // [R][4][nbp]
    uint16_t* list = reinterpret_cast<uint16_t*>(ring + (size_t)R * 4 * a.nbp);               // [FT * nbp]
    uint32_t* smask = reinterpret_cast<uint32_t*>(list + (size_t)FT * a.nbp);                 // [FT][n_words]
    int* n_items = reinterpret_cast<int*>(smask + fused_max_words<FT>());
    float* stage = reinterpret_cast<float*>(&s.scratch[0][0]);                                 // [3][FT][nbp]
    load_fft_smem(s, tb);
    if (threadIdx.x == 0) *n_items = 0;
    __syncthreads();

    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    const LaneTwiddles<T> tw = lane_twiddles(tb, lane);
    const int clip = blockIdx.y;
    const int s0 = blockIdx.x * a.seg_len;
    const int s1 = min(a.n_frames, s0 + a.seg_len);
    const int n_bins = a.upper - a.lower;
    const int n_words = (n_bins + 31) >> 5;
    const int feat_dim = a.bands.n_out;
    const long long chan_stride = (long long)a.n_frames * feat_dim;
    const float* clip_audio = a.audio + (long long)clip * 4 * a.n_samples;
    float* clip_feat = a.feature + (long long)clip * 7 * chan_stride;
    Cx<T>* scratch = s.scratch[warp];
    const int row = 4 * a.nbp;    // float2 per ring slot
    const uint32_t tail_bits = (n_bins & 31) ? ((1u << (n_bins & 31)) - 1u) : 0xffffffffu;
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t ch_bytes = (uint32_t)(a.nbp * sizeof(float2));

    // compacts the selected (frame, bin) items of frames [t0, t0 + nt) into `list`
    auto compact = [&](int t0, int nt) {
        for (int w = warp; w < nt * n_words; w += kWarps) {
            const int tl = w / n_words, wi = w - tl * n_words;
            uint32_t bits = a.mask ? a.mask[((long long)clip * a.n_frames + t0 + tl) * n_words + wi] : 0xffffffffu;
            if (wi == n_words - 1) bits &= tail_bits;
            int base = 0;
            if (lane == 0) {
                smask[tl * n_words + wi] = bits;
                base = atomicAdd(n_items, __popc(bits));
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((bits >> lane) & 1u) list[base + __popc(bits & ((1u << lane) - 1u))] = (uint16_t)((tl << 8) | (wi * 32 + lane));
        }
    };

    // transforms frames [fa, fb) (un-wrapped indices relative to the clip) into the ring; the samples
    // of a warp's next (frame, channel) item are requested before the current one is transformed
    auto frame_start = [&](int f) {
        f %= a.n_frames;                                   // wrap padding of the frame axis (:43)
        if (f < 0) f += a.n_frames;
        return f * a.hop - kNfft / 2;
    };
    auto transform = [&](int fa, int fb, int ct0, int cnt) {
        const int n_items = (fb - fa) * 4;
        float2 raw[8];
        if (warp < n_items) load_frame(clip_audio + (long long)(warp & 3) * a.n_samples, a.n_samples,
                                       frame_start(fa + (warp >> 2)), lane, raw);
        if (cnt > 0) compact(ct0, cnt);                    // overlaps the latency of the loads above
        for (int item = warp; item < n_items; item += kWarps) {
            const int f = fa + (item >> 2), ch = item & 3;
            float2 cur[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cur[i] = raw[i];
            const int nxt = item + kWarps;
            if (nxt < n_items) load_frame(clip_audio + (long long)(nxt & 3) * a.n_samples, a.n_samples,
                                          frame_start(fa + (nxt >> 2)), lane, raw);
            Cx<T> X[8];
            T nyq;
            warp_rfft512<T>(cur, tb.window ? s.win : nullptr, tw, scratch, lane, X, nyq);
            const int slot = (f - (s0 - kHop)) % R;
            float2* dst = ring + slot * row + ch * a.nbp;
            float p[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = lane + 32 * j;
                const float re = (float)X[j].re, im = (float)X[j].im;
                p[j] = power_f32(re, im);
                if (k >= a.lower && k < a.upper) dst[k - a.lower] = make_float2(re, im);
            }
            const float p_nyq = power_f32((float)nyq, 0.0f);
            if (f >= s0 && f < s1) write_logspec_row(p, p_nyq, clip_feat + ch * chan_stride + (long long)f * feat_dim, a.bands, lane);
        }
    };

    transform(s0 - kHop, s0 + kHop, 0, 0);
    for (int t0 = s0; t0 < s1; t0 += FT) {
        const int nt = min(FT, s1 - t0);
        // ---- phase 1
        transform(t0 + kHop, t0 + nt + kHop, t0, nt);
        __syncthreads();
        // ---- phase 2
        const int count = *n_items;
        for (int item = threadIdx.x; item < count; item += kThreads) {
            const int code = list[item];
            const int tl = code >> 8, b = code & 255;
            // shared-memory addresses of X[t - 3 + k][ch = 0][b], k = 0..6 (ring slots wrap at R)
            int slot = (t0 + tl - s0) % R;                 // slot of frame t - 3
            uint32_t fa[kWin];
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                fa[k] = ring_addr + (uint32_t)((slot * row + b) * sizeof(float2));
                slot = slot + 1 == R ? 0 : slot + 1;
            }
            float o[3];
            auto load = [&](int k, int ch) -> float2 { return lds_f2(fa[k] + ch * ch_bytes); };
            eig_bin(load, a.eig, b, o);
#pragma unroll
            for (int i = 0; i < 3; ++i) stage[(i * FT + tl) * a.nbp + b] = o[i];
        }
        __syncthreads();
        // ---- phase 3
        if (threadIdx.x == 0) *n_items = 0;
        if ((feat_dim & 3) == 0) {
            const int groups = feat_dim >> 2;
            for (int g = threadIdx.x; g < 3 * nt * groups; g += kThreads) {
                const int r = g / groups, k = (g - r * groups) * 4;      // r = channel * nt + frame
                const int i = r / nt, tl = r - i * nt;
                float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (k < n_bins) {
                    const uint32_t bits = smask[tl * n_words + (k >> 5)] >> (k & 31);
                    const float* sp = stage + (i * FT + tl) * a.nbp + k;
                    if (bits & 1u) v.x = sp[0];
                    if (bits & 2u) v.y = sp[1];
                    if (bits & 4u) v.z = sp[2];
                    if (bits & 8u) v.w = sp[3];
                }
                *reinterpret_cast<float4*>(clip_feat + (4 + i) * chan_stride + (long long)(t0 + tl) * feat_dim + k) = v;
            }
        } else {
            for (int g = threadIdx.x; g < 3 * nt * feat_dim; g += kThreads) {
                const int r = g / feat_dim, k = g - r * feat_dim;
                const int i = r / nt, tl = r - i * nt;
                float v = 0.0f;
                if (k < n_bins && ((smask[tl * n_words + (k >> 5)] >> (k & 31)) & 1u)) v = stage[(i * FT + tl) * a.nbp + k];
                clip_feat[(4 + i) * chan_stride + (long long)(t0 + tl) * feat_dim + k] = v;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// lite_kernel: SALSA-Lite / SALSA-IPD (salsa_lite_feature_extraction.py:94-123).
// One warp per frame: the four channel transforms run back to back, channel 0 stays in registers.
// ------------------------------------------------------------------------------------------------
struct LiteArgs {
    const float* audio;
    float* feature;      // [clip][7][n_frames][cutoff - lower]
    int n_samples;
    int hop;
    int n_frames;
    int lower, cutoff;   // spectrogram bins lower..cutoff-1
    int upper_cropped;   // spatial values at cropped index >= upper_cropped are zero (:120)
    int mode;            // SALSA_LITE_*
    double inv_delta;    // 1 / delta
    int frames_per_block;
};

// NJ: number of 32-bin register groups of channel 0 that are kept for the phase differences (bins < 32 NJ reach the
// spatial channels): 2 covers the reference's fmax_doa = 2000 Hz (upper_bin 42), 8 is the general case.
template <typename T, int NJ>
__global__ void __launch_bounds__(kThreads, 2) lite_kernel(LiteArgs a, FftTables<T> tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FftSmem<T>& s = *reinterpret_cast<FftSmem<T>*>(smem_raw);
    load_fft_smem(s, tb);
    __syncthreads();
    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    LaneTwiddles<T> tw = lane_twiddles(tb, lane);
    use_twiddle_tables(tw, s.twiddles, lane);
    const int clip = blockIdx.y;
    const int f0 = blockIdx.x * a.frames_per_block;
    const int f1 = min(a.n_frames, f0 + a.frames_per_block);
    const int width = a.cutoff - a.lower;
    const long long chan_stride = (long long)a.n_frames * width;
    const float* clip_audio = a.audio + (long long)clip * 4 * a.n_samples;
    float* clip_feat = a.feature + (long long)clip * 7 * chan_stride;
    Cx<T>* scratch = s.scratch[warp];
    const T* win = tb.window ? s.win : nullptr;
    const float inv_pi = 0.318309886183790671538f;
    const int rl = (32 - lane) & 31;             // position of this lane's bins in the upper groups (reversed order, fft.cuh)
    float2 raw[8];
    if (f0 + warp < f1) load_frame(clip_audio, a.n_samples, (f0 + warp) * a.hop - kNfft / 2, lane, raw);
    for (int t = f0 + warp; t < f1; t += kWarps) {
        float2 x0[NJ];
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {      // one copy of the transform: unrolling the channels spills
            Cx<T> v[8];
            window_frame<T>(raw, win, tw, lane, v);
            // `raw` is consumed: request the next channel of this frame (or channel 0 of this warp's next frame)
            if (ch < 3) load_frame(clip_audio + (long long)(ch + 1) * a.n_samples, a.n_samples, t * a.hop - kNfft / 2, lane, raw);
            else if (t + kWarps < f1) load_frame(clip_audio, a.n_samples, (t + kWarps) * a.hop - kNfft / 2, lane, raw);
            Cx<T> lo[4], hr[4], x128;
            warp_fft_passes_paired<T, STFT_TAB>(v, tw, scratch, lane, lo, hr, x128);
            float2 hf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) hf[j] = make_float2((float)hr[j].re, (float)hr[j].im);
            const float2 x128f = make_float2((float)x128.re, (float)x128.im);
            float* srow = clip_feat + ch * chan_stride + (long long)t * width - a.lower;         // indexed by the bin
            float* prow = clip_feat + (3 + ch) * chan_stride + (long long)t * width - a.lower;   // used for ch >= 1
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int k = 32 * g + (g < 4 ? lane : rl);
                const float2 x = g < 4 ? make_float2((float)lo[g & 3].re, (float)lo[g & 3].im) : upper_group(hf, x128f, g, lane);
                const float re = x.x, im = x.y;
                if (ch == 0 && g < NJ) x0[g < NJ ? g : 0] = x;
                if ((unsigned)(k - a.lower) < (unsigned)width) {
                    srow[k] = power_db(power_f32(re, im));     // (np.abs(stft) ** 2).T -> power_to_db (:104-105)
                    if (ch > 0) {
                        float ph = 0.0f;
                        if (g < NJ && k - a.lower < a.upper_cropped) {
                            // X_ch conj(X_0) with exact float64 products (:111), angle in float32
                            const float2 r = x0[g < NJ ? g : 0];
                            const double pr = (double)re * r.x + (double)im * r.y;
                            const double pi = (double)im * r.x - (double)re * r.y;
                            const float ang = atan2f((float)pi, (float)pr);
                            ph = a.mode == SALSA_LITE_IPD ? ang * inv_pi
                                                          : ang * (float)(a.inv_delta / (double)max(k, 1));
                        }
                        prow[k] = ph;
                    }
                }
            }
        }
    }
}

// lite256_kernel: the same features at the reference's second transform size (n_fft = 256; the SALSA-Lite script takes n_fft
// from the config, salsa_lite_feature_extraction.py:40-66).  The plain form, like stft256_kernel: the 256 real samples of a
// frame through the 256-point complex transform with zero imaginary parts, window (full-length Hann, :97-98) from the table.
template <typename T>
__global__ void __launch_bounds__(kThreads, 2) lite256_kernel(LiteArgs a, FftTables<T> tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FftSmem<T>& s = *reinterpret_cast<FftSmem<T>*>(smem_raw);
    load_fft_smem(s, tb);
    __syncthreads();
    const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
    const LaneTwiddles<T> tw = lane_twiddles(tb, lane);
    const int clip = blockIdx.y;
    const int f0 = blockIdx.x * a.frames_per_block;
    const int f1 = min(a.n_frames, f0 + a.frames_per_block);
    const int width = a.cutoff - a.lower;
    const long long chan_stride = (long long)a.n_frames * width;
    const float* clip_audio = a.audio + (long long)clip * 4 * a.n_samples;
    float* clip_feat = a.feature + (long long)clip * 7 * chan_stride;
    Cx<T>* scratch = s.scratch[warp];
    const float inv_pi = 0.318309886183790671538f;
    for (int t = f0 + warp; t < f1; t += kWarps) {
        // spectrum of one channel of this frame, rounded to complex64 like librosa's: xf[g] = bin lane + 32 g
        auto spectrum = [&](int ch, float2 (&xf)[4]) {
            const float* x = clip_audio + (long long)ch * a.n_samples;
            const int start = t * a.hop - kNfft256 / 2;
            Cx<T> v[8], z[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const int m = lane + 32 * n1;
                v[n1] = {(T)x[reflect_index(start + m, a.n_samples)] * s.win[m], (T)0};
            }
            warp_fft_core<T, 0>(v, tw, scratch, lane, z);
#pragma unroll
            for (int g = 0; g < 4; ++g) xf[g] = make_float2((float)z[g].re, (float)z[g].im);
        };
        float2 x0[4];
        spectrum(0, x0);
        {
            float* srow = clip_feat + (long long)t * width - a.lower;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int k = lane + 32 * g;
                if ((unsigned)(k - a.lower) < (unsigned)width) srow[k] = power_db(power_f32(x0[g].x, x0[g].y));
            }
        }
#pragma unroll 1
        for (int ch = 1; ch < 4; ++ch) {
            float2 xc[4];
            spectrum(ch, xc);
            float* srow = clip_feat + ch * chan_stride + (long long)t * width - a.lower;
            float* prow = clip_feat + (3 + ch) * chan_stride + (long long)t * width - a.lower;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int k = lane + 32 * g;
                const float re = xc[g].x, im = xc[g].y;
                if ((unsigned)(k - a.lower) < (unsigned)width) {
                    srow[k] = power_db(power_f32(re, im));
                    float ph = 0.0f;
                    if (k - a.lower < a.upper_cropped) {
                        // X_ch conj(X_0) with exact float64 products (:111), angle in float32
                        const float2 r = x0[g];
                        const double pr = (double)re * r.x + (double)im * r.y;
                        const double pi = (double)im * r.x - (double)re * r.y;
                        const float ang = atan2f((float)pi, (float)pr);
                        ph = a.mode == SALSA_LITE_IPD ? ang * inv_pi : ang * (float)(a.inv_delta / (double)max(k, 1));
                    }
                    prow[k] = ph;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// iv_kernel: the intensity-vector channels of LinSpecIvExtractor.extract (dataset/feature_extraction.py:342-351), FOA:
//     IV_c = Re(conj(X_0) X_c), c = 1..3;  normal = sqrt(IVx^2 + IVy^2 + IVz^2) + 1e-8;  feature = W (IV_c / normal)
// in float32 like the reference's complex64 arithmetic, with W the log-linear band matrix of the spectrogram (:273-300:
// 192 single bins, then 8 bands of 8 bins -- the last of 7 -- weighted 1/8).  grid (frames, clips), thread = bin - 1;
// X rows [clip][frame][4][x_pitch] hold bins 1 .. x_pitch (written by stft_kernel's row layout).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) iv_kernel(const float2* __restrict__ X, float* __restrict__ feature, int n_frames, int x_pitch,
                                                 BandLayout bands) {
    __shared__ float s_iv[3][256];
    const int t = blockIdx.x, clip = blockIdx.y, k = threadIdx.x;          // k = bin - 1
    const float2* row = X + ((long long)clip * n_frames + t) * 4 * x_pitch;
    float iv[3] = {0.0f, 0.0f, 0.0f};
    if (k < x_pitch) {
        const float2 x0 = __ldg(row + k);
        float sq = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float2 xc = __ldg(row + (c + 1) * x_pitch + k);
            iv[c] = fmaf(x0.x, xc.x, x0.y * xc.y);
            sq = fmaf(iv[c], iv[c], sq);
        }
        const float normal = sqrtf(sq) + 1e-8f;
#pragma unroll
        for (int c = 0; c < 3; ++c) iv[c] = iv[c] / normal;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) s_iv[c][k] = iv[c];
    __syncthreads();
    const long long chan_stride = (long long)n_frames * bands.n_out;
    float* out = feature + (long long)clip * 7 * chan_stride + 4 * chan_stride + (long long)t * bands.n_out;
    if (k < bands.n_out) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v;
            if (k < bands.n_lin) {
                v = s_iv[c][k];                                   // band k = bin k + 1
            } else {
                const int first = bands.n_lin + 8 * (k - bands.n_lin);      // bins first + 1 .. first + 8, clipped below the Nyquist bin
                v = 0.0f;
                for (int j = 0; j < 8; ++j)
                    if (first + j < kHalf - 1) v += 0.125f * s_iv[c][first + j];
            }
            out[c * chan_stride + k] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GCC-PHAT of LogSpecGccExtractor (dataset/feature_extraction.py:411-445): per channel pair the inverse real FFT of the
// unit phasors of the cross spectrum of a 1024-point STFT (the 512-sample window zero-padded to 1024, librosa's
// win_length < n_fft), of which the 200 lags -100 .. 99 are kept.
//   * the 1024-point spectrum of a zero-padded 512-sample frame is two 512-point transforms: even bins 2m = F[m] with
//     F = DFT512(w x); odd bins 2m+1 = S[m] + i C[m] (up to a sign common to all channels) with C / S = DFT512(w x cos(pi j/512))
//     / DFT512(w x sin(pi j/512)): three passes of stft_kernel with three windows (XE bins 0..256, XC / XS bins 0..255);
//   * gcc_unit_kernel: U[k] = R / |R|, R = X_sig[k] conj(X_ref[k]) (= exp(i angle(R)), 1 where R = 0) for the six pairs,
//     written as the bf16x2-split rows of a GEMM operand: row (clip, pair, frame), columns [Re U0, Re U512, Re U1, Im U1, ...];
//   * the inverse transform restricted to 200 lags is that row times a constant 1024 x 200 cosine / sine table: crnn_gemm
//     (tcgen05, two bf16 planes per operand = float32-grade), then gcc_scatter_kernel puts the rows into the feature layout.
// ------------------------------------------------------------------------------------------------
constexpr int kGccK = 1024;          // GEMM K: 2 real + 2 x 511 (re, im) values per row
constexpr int kGccLags = 200, kGccLagsPad = 256;

__global__ void __launch_bounds__(256) gcc_unit_kernel(const float2* __restrict__ XE, const float2* __restrict__ XC,
                                                       const float2* __restrict__ XS, __nv_bfloat16* __restrict__ A, int n_frames) {
    const int t = blockIdx.x, clip = blockIdx.y;
    const long long ft = (long long)clip * n_frames + t;
    const float2* e = XE + ft * 4 * (kHalf + 1);
    const float2* c = XC + ft * 4 * kHalf;
    const float2* sn = XS + ft * 4 * kHalf;
    for (int k = threadIdx.x; k <= 2 * kHalf; k += blockDim.x) {
        float2 x[4];
        const int m = k >> 1;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            if (k & 1) {
                const float2 cc = __ldg(c + ch * kHalf + m), ss = __ldg(sn + ch * kHalf + m);
                x[ch] = make_float2(ss.x - cc.y, ss.y + cc.x);           // S + i C
            } else {
                x[ch] = __ldg(e + ch * (kHalf + 1) + m);
            }
        }
        int pair = 0;
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int mm = n + 1; mm < 4; ++mm, ++pair) {
                // R = X_sig conj(X_ref), sig = mm, ref = n (:135-136, :431)
                const float re = fmaf(x[mm].x, x[n].x, x[mm].y * x[n].y), im = fmaf(x[mm].y, x[n].x, -x[mm].x * x[n].y);
                const float mag2 = fmaf(re, re, im * im);
                float ur = 1.0f, ui = 0.0f;                               // angle(0) = 0
                if (mag2 > 0.0f) {
                    const float inv = rsqrtf(mag2);
                    ur = re * inv;
                    ui = im * inv;
                }
                __nv_bfloat16* row = A + (((long long)clip * 6 + pair) * n_frames + t) * (2 * kGccK);
                auto put = [&](int col, float v) {
                    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                    row[col] = hi;
                    row[kGccK + col] = __float2bfloat16_rn(v - __bfloat162float(hi));
                };
                if (k == 0) put(0, ur);
                else if (k == 2 * kHalf) put(1, ur);
                else {
                    put(2 * k, ur);
                    put(2 * k + 1, ui);
                }
            }
    }
}

// G [rows = (clip, pair, frame)][256] fp32 -> feature [clip][10][T][200], channels 4 + pair
__global__ void __launch_bounds__(256) gcc_scatter_kernel(const float* __restrict__ G, float* __restrict__ feature, int n_frames, int feat_dim) {
    const int t = blockIdx.x, pair = blockIdx.y, clip = blockIdx.z;
    const float* src = G + (((long long)clip * 6 + pair) * n_frames + t) * kGccLagsPad;
    float* dst = feature + (((long long)clip * 10 + 4 + pair) * n_frames + t) * feat_dim;
    for (int i = threadIdx.x; i < feat_dim; i += blockDim.x) dst[i] = i < kGccLags ? src[i] : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// pcm16_to_float_kernel: 16-bit PCM samples -> float32, sample / 32768 (exact), as soundfile / librosa.load hand the wav
// files of the dataset to the reference (salsa_feature_extraction.py:353).  8 samples per thread and step.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pcm16_to_float_kernel(const int16_t* __restrict__ pcm, float* __restrict__ audio, long long n) {
    const long long groups = n >> 3;
    const float s = 1.0f / 32768.0f;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(pcm) + g);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float4 lo, hi;
        lo.x = (float)(short)(w[0] & 0xffff) * s; lo.y = (float)(w[0] >> 16) * s;
        lo.z = (float)(short)(w[1] & 0xffff) * s; lo.w = (float)(w[1] >> 16) * s;
        hi.x = (float)(short)(w[2] & 0xffff) * s; hi.y = (float)(w[2] >> 16) * s;
        hi.z = (float)(short)(w[3] & 0xffff) * s; hi.w = (float)(w[3] >> 16) * s;
        reinterpret_cast<float4*>(audio)[2 * g] = lo;
        reinterpret_cast<float4*>(audio)[2 * g + 1] = hi;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const long long i = (groups << 3) + threadIdx.x;
        audio[i] = (float)pcm[i] * s;
    }
}

// ------------------------------------------------------------------------------------------------
// scaler_accumulate_kernel: per-channel (0..3), per-frequency sum and sum of squares over all frames of
// all clips, float64 -- the statistics compute_scaler() gathers with StandardScaler.partial_fit
// (salsa_feature_extraction.py:204-262).  grid (frame chunks, 4 channels, clips), thread = frequency.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scaler_accumulate_kernel(const float* __restrict__ feature, int n_frames, int feat_dim,
                                                                int n_feat_chans, int frames_per_block, double* __restrict__ sums) {
    const int f = threadIdx.x;
    if (f >= feat_dim) return;
    const int ch = blockIdx.y, clip = blockIdx.z;
    const int t0 = blockIdx.x * frames_per_block, t1 = min(n_frames, t0 + frames_per_block);
    const float* p = feature + (((long long)clip * n_feat_chans + ch) * n_frames + t0) * feat_dim + f;
    double s1 = 0.0, s2 = 0.0;
    for (int t = t0; t < t1; ++t, p += feat_dim) {
        const double v = (double)__ldg(p);
        s1 += v;
        s2 = fma(v, v, s2);
    }
    atomicAdd(sums + ((long long)ch * feat_dim + f) * 2, s1);
    atomicAdd(sums + ((long long)ch * feat_dim + f) * 2 + 1, s2);
}

}  // namespace salsa
