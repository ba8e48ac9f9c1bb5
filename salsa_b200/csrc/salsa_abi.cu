// C ABI of libsalsa_b200.so: feature-extraction entry points (see include/salsa_b200.h).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "salsa_crnn.h"
#include "salsa_kernels.cuh"

namespace salsa {

// ---------------------------------------------------------------------------------------------
// errors, launch counter
// ---------------------------------------------------------------------------------------------
static thread_local std::string t_error;
static thread_local uint64_t t_launches = 0;

void set_error(const std::string& msg) { t_error = msg; }
int fail(int code, const std::string& msg) {
    set_error(msg);
    return code;
}
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return SALSA_OK;
    return fail(SALSA_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
void count_launch(int n) { t_launches += (uint64_t)n; }

static int check_launch(const char* name) {
    count_launch();
    return check_cuda(cudaGetLastError(), name);
}

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
struct ProfRecord {
    const char* name;
    cudaEvent_t start, stop;
};
static thread_local bool t_profiling = false;
static thread_local std::vector<ProfRecord> t_prof;

struct ProfScope {
    cudaStream_t st;
    bool on;
    ProfRecord rec;
    ProfScope(const char* name, cudaStream_t s) : st(s), on(t_profiling) {
        if (!on) return;
        rec.name = name;
        cudaEventCreate(&rec.start);
        cudaEventCreate(&rec.stop);
        cudaEventRecord(rec.start, st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(rec.stop, st);
        t_prof.push_back(rec);
    }
};

// ---------------------------------------------------------------------------------------------
// per-device tables (twiddles once per device, windows cached by value)
// ---------------------------------------------------------------------------------------------
namespace {
struct WindowEntry {
    std::vector<double> values;
    double* d;
    float* f;
};
struct DeviceState {
    bool ready = false;
    Cx<double>*tw_a_d = nullptr, *tw_b_d = nullptr, *tw_r_d = nullptr;
    Cx<float>*tw_a_f = nullptr, *tw_b_f = nullptr, *tw_r_f = nullptr;
    std::vector<WindowEntry> windows;
};
std::mutex g_mutex;
DeviceState g_state[64];

template <typename T>
int upload(const std::vector<Cx<long double>>& src, Cx<T>** dst) {
    std::vector<Cx<T>> h(src.size());
    for (size_t i = 0; i < src.size(); ++i) h[i] = {(T)src[i].re, (T)src[i].im};
    SALSA_CUDA(cudaMalloc((void**)dst, h.size() * sizeof(Cx<T>)));
    SALSA_CUDA(cudaMemcpy(*dst, h.data(), h.size() * sizeof(Cx<T>), cudaMemcpyHostToDevice));
    return SALSA_OK;
}

Cx<long double> unit_root(long long num, long long den) {   // exp(-2 pi i num / den)
    const long double two_pi = 6.283185307179586476925286766559005768L;
    num %= den;
    const long double ang = two_pi * (long double)num / (long double)den;
    return {cosl(ang), -sinl(ang)};
}
}  // namespace

int get_tables(const double* window_host, DeviceTables* out) {
    int dev = 0;
    SALSA_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(SALSA_EINVAL, "device index out of range");
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState& st = g_state[dev];
    if (!st.ready) {
        std::vector<Cx<long double>> ta(8 * 32), tb(8 * 32), tr(kHalf);
        for (int k1 = 0; k1 < 8; ++k1)
            for (int lane = 0; lane < 32; ++lane) {
                ta[k1 * 32 + lane] = unit_root((long long)lane * k1, 256);
                tb[k1 * 32 + lane] = unit_root((long long)(lane & 3) * k1, 32);
            }
        for (int k = 0; k < kHalf; ++k) tr[k] = unit_root(k, 512);
        int rc;
        if ((rc = upload(ta, &st.tw_a_d)) || (rc = upload(tb, &st.tw_b_d)) || (rc = upload(tr, &st.tw_r_d)) ||
            (rc = upload(ta, &st.tw_a_f)) || (rc = upload(tb, &st.tw_b_f)) || (rc = upload(tr, &st.tw_r_f)))
            return rc;
        st.ready = true;
    }
    if (!window_host) {       // built-in periodic Hann of full length: computed in the kernels, no table
        out->d = {st.tw_a_d, st.tw_b_d, st.tw_r_d, nullptr};
        out->f = {st.tw_a_f, st.tw_b_f, st.tw_r_f, nullptr};
        return SALSA_OK;
    }
    const WindowEntry* hit = nullptr;
    for (const WindowEntry& w : st.windows)
        if (memcmp(w.values.data(), window_host, kNfft * sizeof(double)) == 0) hit = &w;
    if (!hit) {
        WindowEntry w;
        w.values.assign(window_host, window_host + kNfft);
        std::vector<float> wf(kNfft);
        for (int i = 0; i < kNfft; ++i) wf[i] = (float)window_host[i];
        SALSA_CUDA(cudaMalloc((void**)&w.d, kNfft * sizeof(double)));
        SALSA_CUDA(cudaMalloc((void**)&w.f, kNfft * sizeof(float)));
        SALSA_CUDA(cudaMemcpy(w.d, window_host, kNfft * sizeof(double), cudaMemcpyHostToDevice));
        SALSA_CUDA(cudaMemcpy(w.f, wf.data(), kNfft * sizeof(float), cudaMemcpyHostToDevice));
        st.windows.push_back(w);
        hit = &st.windows.back();
    }
    out->d = {st.tw_a_d, st.tw_b_d, st.tw_r_d, hit->d};
    out->f = {st.tw_a_f, st.tw_b_f, st.tw_r_f, hit->f};
    return SALSA_OK;
}

// `out` has kNfft entries; a shorter transform (n_fft = 256) uses the first n_fft of them, the rest is zero
void host_window(const salsa_params_t* p, double* out) {
    for (int i = 0; i < kNfft; ++i) out[i] = 0.0;
    if (p->window) {
        memcpy(out, p->window, p->n_fft * sizeof(double));
        return;
    }
    // scipy.signal.get_window('hann', win_len, fftbins=True), centred in n_fft (librosa pad_center)
    const double two_pi = 6.283185307179586476925286766559;
    const int lpad = (p->n_fft - p->win_len) / 2;
    for (int i = 0; i < p->n_fft; ++i) out[i] = 0.0;
    for (int i = 0; i < p->win_len; ++i) out[lpad + i] = 0.5 - 0.5 * cos(two_pi * (double)i / (double)p->win_len);
}

int validate_params(const salsa_params_t* p, bool with_stft) {
    if (!p) return fail(SALSA_EINVAL, "params is NULL");
    if (p->n_chans != 4) return fail(SALSA_EINVAL, "n_chans must be 4");
    if (p->n_clips < 0) return fail(SALSA_EINVAL, "n_clips is negative");
    if (p->fs <= 0 || p->n_fft <= 0) return fail(SALSA_EINVAL, "fs and n_fft must be positive");
    if (with_stft) {
        if (p->n_fft != 512 && p->n_fft != 256) return fail(SALSA_EINVAL, "nfft is not 512 or 256");
        if (p->hop_len <= 0) return fail(SALSA_EINVAL, "hop_len must be positive");
        if (p->win_len <= 0 || p->win_len > p->n_fft)
            return fail(SALSA_EINVAL, "Windown length is greater than nfft!");
        if (p->n_samples <= p->n_fft / 2)
            return fail(SALSA_EINVAL, "n_samples must exceed n_fft/2 (reflect padding)");
        if (p->upper_bin > p->n_fft / 2) return fail(SALSA_EINVAL, "upper_bin exceeds n_fft/2");
    }
    if (p->lower_bin < 1 || p->upper_bin <= p->lower_bin)
        return fail(SALSA_EINVAL, "need 1 <= lower_bin < upper_bin");
    if (p->audio_format != SALSA_FORMAT_FOA && p->audio_format != SALSA_FORMAT_MIC)
        return fail(SALSA_EINVAL, "audio format is not valid");
    if (p->n_hopframes != kHop) return fail(SALSA_EINVAL, "n_hopframes must be 3");
    if (p->stft_precision != 32 && p->stft_precision != 64)
        return fail(SALSA_EINVAL, "stft_precision must be 32 or 64");
    return SALSA_OK;
}

static BandLayout band_layout(const salsa_params_t* p) {
    const int half = p->n_fft / 2;
    if (!p->is_compress_high_freq) return {half, half};
    return {half * 3 / 4, half * 3 / 4 + half / 32};   // 512: 192 linear + 8 compressed = 200 (:153-162); 256: 96 + 4 = 100 (:163-170)
}

static EigArgs eig_args(const salsa_params_t* p) {
    EigArgs e;
    e.format = p->audio_format;
    e.test = p->is_tracking ? 1 : 0;
    e.cond = (float)p->cond_num;
    e.cond_d = p->cond_num;
    // n squarings and m (1 or 2) products of the squared matrix with its dominant column give the exponent
    // (1 + m) 2^n; a bin is only kept when lambda1 > cond * lambda2, so choose the cheapest (n, m) with
    // cond^-exponent < 1e-7 (a product costs half a squaring): cond = 5 -> n = 2, m = 2, exponent 12, 4e-9.
    // Without a usable gap (no test, cond <= 1) iterate longer.
    int n_sq = 10, n_mv = 1;
    if (e.test && p->cond_num > 1.0) {
#ifndef EIG_CONTAMINATION
#define EIG_CONTAMINATION 1e-7
#endif
        const double need = log(1.0 / EIG_CONTAMINATION) / log(p->cond_num);
        n_sq = (int)ceil(log2(need)) - 1;
        n_sq = std::max(2, std::min(10, n_sq));
        if (n_sq > 2 && 3.0 * ldexp(1.0, n_sq - 1) >= need) {
            n_sq -= 1;
            n_mv = 2;
        }
    }
    e.n_sq = n_sq;
    e.n_mv = n_mv;
    const double delta = 2.0 * M_PI * (double)p->fs / ((double)p->n_fft * 343.0);
    e.inv_delta = 1.0 / delta;
    e.lower = p->lower_bin;
    return e;
}

static TrackerConsts tracker_consts() {
    const double alpha = 0.02, slow_scale = 0.1;    // salsa_feature_extraction.py:30-31
    TrackerConsts c;
    c.floor_up = 1 + alpha;
    c.floor_up_slow = 1 + slow_scale * alpha;
    c.floor_down = 1 - alpha;
    c.snr_ratio = 1.5;
    c.floor_min = 1e-6;
    c.n_sig_frames = 3;
    c.n_init_frames = 5;
    return c;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                      "cudaFuncSetAttribute");
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
template <typename T, int CH, int OUT>
static int launch_stft_t(const StftArgs& a, const FftTables<T>& tb, dim3 grid, cudaStream_t st) {
    const size_t smem = sizeof(FftSmemW<T, kStftWarps>);
    int rc = set_smem(stft_kernel<T, CH, OUT>, smem);
    if (rc) return rc;
    stft_kernel<T, CH, OUT><<<grid, kStftThreads, smem, st>>>(a, tb);
    return SALSA_OK;
}

static int launch_stft(const salsa_params_t* p, const DeviceTables& tb, const float* audio, float2* X, int x_pitch, int x_tiles,
                       float* spec, long long spec_clip_stride, double* power0, int ch_count, cudaStream_t st) {
    if (p->n_clips == 0) return SALSA_OK;
    const int n_frames = salsa_n_frames(p->n_samples, p->hop_len);
    StftArgs a;
    a.audio = audio;
    a.n_chans = p->n_chans;
    a.n_samples = p->n_samples;
    a.hop = p->hop_len;
    a.n_frames = n_frames;
    a.lower = p->lower_bin;
    a.upper = p->upper_bin;
    a.ch_count = ch_count;
    // every warp of a CTA transforms 16 (frame, channel) items: 32 frames per CTA with 8 warps and 4 channels
    a.frames_per_block = 16 * kStftWarps / ch_count;
    a.bands = band_layout(p);
    a.X = X;
    a.x_pitch = x_pitch;
    a.x_tiles = x_tiles;
    a.spec = spec;
    a.spec_clip_stride = spec_clip_stride;
    a.spec_chan_stride = (long long)n_frames * a.bands.n_out;
    a.power0 = power0;
    dim3 grid((n_frames + a.frames_per_block - 1) / a.frames_per_block, p->n_clips);
    ProfScope prof("stft_kernel", st);
    int rc;
    if (ch_count != 1 && ch_count != 4) return fail(SALSA_EINVAL, "stft: 1 or 4 channels");
    const bool d = p->stft_precision == 64;
    if (p->n_fft == 256) {                                   // the plain kernel of the second transform size
        if (!tb.d.window) return fail(SALSA_EINVAL, "stft: n_fft = 256 takes its window from a table");
        const size_t smem = d ? sizeof(FftSmemW<double, kStftWarps>) : sizeof(FftSmemW<float, kStftWarps>);
        if ((rc = d ? set_smem(stft256_kernel<double>, smem) : set_smem(stft256_kernel<float>, smem))) return rc;
        if (d) stft256_kernel<double><<<grid, kStftThreads, smem, st>>>(a, tb.d);
        else stft256_kernel<float><<<grid, kStftThreads, smem, st>>>(a, tb.f);
        return check_launch("stft256_kernel");
    }
    if (x_tiles > 0 && (ch_count != 4 || !X || power0)) return fail(SALSA_EINVAL, "stft: tiled X is the clip path's layout");
    if (x_tiles > 0 && !spec)                               // clip path with a separate spectrogram window: X only
        rc = p->lower_bin < kTileBins
                 ? (d ? launch_stft_t<double, 4, kStftX | kStftTiled>(a, tb.d, grid, st)
                      : launch_stft_t<float, 4, kStftX | kStftTiled>(a, tb.f, grid, st))
                 : (d ? launch_stft_t<double, 4, kStftX | kStftTiledAny>(a, tb.d, grid, st)
                      : launch_stft_t<float, 4, kStftX | kStftTiledAny>(a, tb.f, grid, st));
    else if (x_tiles > 0 && p->lower_bin < kTileBins)       // clip path, split arrangement
        rc = d ? launch_stft_t<double, 4, kStftX | kStftSpec | kStftTiled>(a, tb.d, grid, st)
               : launch_stft_t<float, 4, kStftX | kStftSpec | kStftTiled>(a, tb.f, grid, st);
    else if (x_tiles > 0)
        rc = d ? launch_stft_t<double, 4, kStftX | kStftSpec | kStftTiledAny>(a, tb.d, grid, st)
               : launch_stft_t<float, 4, kStftX | kStftSpec | kStftTiledAny>(a, tb.f, grid, st);
    else if (ch_count == 1 && !X && !spec && power0)        // clip path, fused arrangement (pass A)
        rc = d ? launch_stft_t<double, 1, kStftPower0>(a, tb.d, grid, st) : launch_stft_t<float, 1, kStftPower0>(a, tb.f, grid, st);
    else if (ch_count == 4)
        rc = d ? launch_stft_t<double, 4, kStftAny>(a, tb.d, grid, st) : launch_stft_t<float, 4, kStftAny>(a, tb.f, grid, st);
    else
        rc = d ? launch_stft_t<double, 1, kStftAny>(a, tb.d, grid, st) : launch_stft_t<float, 1, kStftAny>(a, tb.f, grid, st);
    if (rc) return rc;
    return check_launch("stft_kernel");
}

template <typename Src>
static int launch_tracker(Src src, uint32_t* mask, int n_clips, int n_frames, int first_bin, int n_bins, cudaStream_t st) {
    if (n_clips == 0) return SALSA_OK;
    const int n_words = (n_bins + 31) / 32;
    const int wpb = std::min(n_words, 8);                       // warps per block
    dim3 grid((n_words + wpb - 1) / wpb, n_clips);
    ProfScope prof("tracker_kernel", st);
    tracker_kernel<Src><<<grid, wpb * 32, 0, st>>>(src, mask, n_frames, first_bin, n_bins, tracker_consts());
    return check_launch("tracker_kernel");
}

static TrackerPower0 tracker_power0(const double* power0, int n_frames, int n_bins) {
    return {power0, (long long)n_frames * n_bins, (long long)n_bins};
}

constexpr int kFusedFT = 4;   // new frames per step of salsa_fused_kernel

static int choose_seg_len(int n_clips, int n_frames) {
    // aim for >= 4 CTAs per SM slot (2 resident CTAs x 148 SMs) while keeping the 6-frame halo small
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = 4LL * 2 * sms;
    long long segs_per_clip = (want + n_clips - 1) / std::max(1, n_clips);
    segs_per_clip = std::max(1LL, std::min<long long>(segs_per_clip, (n_frames + 31) / 32));
    int seg = (int)((n_frames + segs_per_clip - 1) / segs_per_clip);
    seg = std::min(seg, 240);
    seg = (seg + kFusedFT - 1) / kFusedFT * kFusedFT;
    return std::max(seg, kFusedFT);
}

static int launch_fused(const salsa_params_t* p, const DeviceTables& tb, const float* audio, float* feature,
                        const uint32_t* mask, cudaStream_t st) {
    if (p->n_clips == 0) return SALSA_OK;
    FusedArgs a;
    a.audio = audio;
    a.feature = feature;
    a.mask = mask;
    a.n_samples = p->n_samples;
    a.hop = p->hop_len;
    a.n_frames = salsa_n_frames(p->n_samples, p->hop_len);
    a.lower = p->lower_bin;
    a.upper = p->upper_bin;
    a.nbp = (p->upper_bin - p->lower_bin + 31) / 32 * 32;
    a.seg_len = choose_seg_len(p->n_clips, a.n_frames);
    a.bands = band_layout(p);
    a.eig = eig_args(p);
    dim3 grid((a.n_frames + a.seg_len - 1) / a.seg_len, p->n_clips);
    if (a.nbp > 256) return fail(SALSA_EINVAL, "more than 256 spatial bins");
    ProfScope prof("salsa_fused_kernel", st);
    if (p->stft_precision == 64) {
        const size_t smem = fused_smem_bytes<double, kFusedFT>(a.nbp);
        int rc = set_smem(salsa_fused_kernel<double, kFusedFT>, smem);
        if (rc) return rc;
        salsa_fused_kernel<double, kFusedFT><<<grid, kThreads, smem, st>>>(a, tb.d);
    } else {
        const size_t smem = fused_smem_bytes<float, kFusedFT>(a.nbp);
        int rc = set_smem(salsa_fused_kernel<float, kFusedFT>, smem);
        if (rc) return rc;
        salsa_fused_kernel<float, kFusedFT><<<grid, kThreads, smem, st>>>(a, tb.f);
    }
    return check_launch("salsa_fused_kernel");
}

// ---------------------------------------------------------------------------------------------
// The clip path exists in two arrangements of the same arithmetic (same selection bit for bit, values to float32
// rounding: tests/test_gpu_features_fullsize.py):
//   split  stft_kernel (all channels: X -> HBM in tiles, log-spectrogram rows) | tracker_kernel (on channel 0 of X) |
//          eig_tile_kernel (+ eig_redo_kernel for the few bins that need float64).  X costs 59 MB of
//          HBM traffic per clip on top of the 50 MB of algorithmic bytes, but every kernel runs at its own register
//          budget / occupancy and the channel-0 transform is not done twice: 25.8 against 40.7 ms per 600 clips.  Default.
//   fused  stft_kernel (channel 0 only -> |X0|^2) | tracker_kernel | salsa_fused_kernel (X lives in a shared-memory
//          ring).  Minimal HBM traffic; selected with SALSA_B200_PIPELINE=fused.
// ---------------------------------------------------------------------------------------------
enum Pipeline { kPipelineSplit = 0, kPipelineFused = 1 };

static Pipeline pipeline_choice() {
    const char* e = getenv("SALSA_B200_PIPELINE");
    if (e && strcmp(e, "fused") == 0) return kPipelineFused;
    return kPipelineSplit;
}

struct Workspace {
    double* power0;    // fused
    float2* X;         // split: tiled spectrum
    uint32_t* mask;    // [clip][frame][words of 32 spatial bins]
    uint32_t* redo;    // split
    int n_tiles;       // split
    size_t bytes;
};

static size_t round256(size_t v) { return (v + 255) / 256 * 256; }

static Workspace carve_workspace(const salsa_params_t* p, void* base, Pipeline pl) {
    const size_t n_frames = (size_t)salsa_n_frames(p->n_samples, p->hop_len);
    const size_t n_bins = (size_t)(p->upper_bin - p->lower_bin);
    char* at = reinterpret_cast<char*>(base);
    Workspace w = {};
    if (pl == kPipelineSplit) {
        w.n_tiles = (int)((n_bins + kTileBins - 1) / kTileBins);
        const size_t xb = round256((size_t)p->n_clips * w.n_tiles * n_frames * kTileFrameElems * sizeof(float2));
        const size_t mk = round256((size_t)p->n_clips * n_frames * w.n_tiles * sizeof(uint32_t));
        w.X = reinterpret_cast<float2*>(at);
        w.mask = reinterpret_cast<uint32_t*>(at + xb);
        w.redo = reinterpret_cast<uint32_t*>(at + xb + mk);
        w.bytes = xb + 2 * mk;
    } else {
        const size_t pw = round256((size_t)p->n_clips * n_frames * n_bins * sizeof(double));
        const size_t mk = round256((size_t)p->n_clips * n_frames * ((n_bins + 31) / 32) * sizeof(uint32_t));
        w.power0 = reinterpret_cast<double*>(at);
        w.mask = reinterpret_cast<uint32_t*>(at + pw);
        w.bytes = pw + mk;
        if (!p->is_tracking) w.bytes = 256;
    }
    return w;
}

template <int FT, int MINB, int NSQ, int WARPS>
static int launch_eig_tile_t(const EigTileArgs& a, cudaStream_t st, int n_clips) {
    constexpr size_t smem = eig_tile_smem_bytes<FT>();
    int rc = set_smem(eig_tile_kernel<FT, MINB, NSQ, WARPS>, smem);
    if (rc) return rc;
    dim3 grid((a.n_frames + FT - 1) / FT, a.n_tiles, n_clips);
    eig_tile_kernel<FT, MINB, NSQ, WARPS><<<grid, WARPS * 32, smem, st>>>(a);
    return SALSA_OK;
}

static int launch_eig_tile(const salsa_params_t* p, const Workspace& w, const uint32_t* mask, float* feature, cudaStream_t st) {
    if (p->n_clips == 0) return SALSA_OK;
    EigTileArgs a;
    a.X = w.X;
    a.mask = mask;
    a.redo = w.redo;
    a.feature = feature;
    a.n_frames = salsa_n_frames(p->n_samples, p->hop_len);
    a.n_bins = p->upper_bin - p->lower_bin;
    a.n_tiles = w.n_tiles;
    a.feat_dim = band_layout(p).n_out;
    a.eig = eig_args(p);
    if (a.feat_dim & 3) return fail(SALSA_EINVAL, "feature width must be a multiple of 4");
    // the reference places (upper_bin - lower_bin) columns into a (3, T, freq_dim) array (:373-374): more bins than
    // columns is its broadcast error
    if (a.n_bins > a.feat_dim) return fail(SALSA_EINVAL, "upper_bin - lower_bin exceeds the feature width");
    int rc;
    {
        ProfScope prof("eig_tile_kernel", st);
        // 8 warps per CTA, 4 CTAs per SM.  (4 warps per CTA: 94 % instead of 78 % of the thread slots of a tile's rounds are
        // used, but 16 warps per SM: 10.4 against 9.95 ms per 600 clips; 2 warps: 13.5 ms.)
        if (a.eig.n_sq == 2)       // the default (cond_num = 5): squarings unrolled at compile time
            rc = launch_eig_tile_t<32, 4, 2, 8>(a, st, p->n_clips);
        else
            rc = launch_eig_tile_t<32, 3, 0, 8>(a, st, p->n_clips);
        if (rc) return rc;
        if ((rc = check_launch("eig_tile_kernel"))) return rc;
    }
    if (!a.eig.test) return SALSA_OK;      // without the coherence test no verdict is ever ambiguous
    const long long n_words_total = (long long)p->n_clips * a.n_frames * a.n_tiles;
    ProfScope prof("eig_redo_kernel", st);
    eig_redo_kernel<<<(unsigned)((n_words_total + 127) / 128), 128, 0, st>>>(a, n_words_total);
    return check_launch("eig_redo_kernel");
}

}  // namespace salsa

// =============================================================================================
// extern "C"
// =============================================================================================
using namespace salsa;

extern "C" {

const char* salsa_last_error(void) { return t_error.c_str(); }
const char* salsa_version(void) { return "salsa_b200 0.1 (sm_100a)"; }

int32_t salsa_n_frames(int32_t n_samples, int32_t hop_len) { return hop_len > 0 ? 1 + n_samples / hop_len : 0; }

int32_t salsa_feat_dim(const salsa_params_t* p) { return p ? band_layout(p).n_out : 0; }

uint64_t salsa_launch_count(int reset) {
    const uint64_t v = t_launches;
    if (reset) t_launches = 0;
    return v;
}

int salsa_profile_enable(int on) {
    t_profiling = on != 0;
    return SALSA_OK;
}

int salsa_profile_read(int32_t max_entries, char* names, double* total_ms, int64_t* launches) {
    int n = 0;
    for (const ProfRecord& r : t_prof) {
        float ms = 0.0f;
        cudaEventSynchronize(r.stop);
        cudaEventElapsedTime(&ms, r.start, r.stop);
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
        int slot = -1;
        for (int i = 0; i < n; ++i)
            if (strncmp(names + 32 * i, r.name, 31) == 0) slot = i;
        if (slot < 0) {
            if (n >= max_entries) continue;
            slot = n++;
            strncpy(names + 32 * slot, r.name, 31);
            names[32 * slot + 31] = 0;
            total_ms[slot] = 0.0;
            launches[slot] = 0;
        }
        total_ms[slot] += ms;
        launches[slot] += 1;
    }
    t_prof.clear();
    return n;
}

int salsa_stft(const salsa_params_t* p, const float* audio, float* X, float* logspec, double* power0, void* stream) {
    int rc = validate_params(p);
    if (rc) return rc;
    if (p->n_clips == 0) return SALSA_OK;
    if (!audio) return fail(SALSA_EINVAL, "audio is NULL");
    double win[kNfft];
    const bool builtin_hann = !p->window && p->win_len == p->n_fft && p->n_fft == kNfft;      // computed in the kernels, no table
    if (!builtin_hann) host_window(p, win);
    DeviceTables tb;
    if ((rc = get_tables(builtin_hann ? nullptr : win, &tb))) return rc;
    const int n_frames = salsa_n_frames(p->n_samples, p->hop_len);
    const long long clip_stride = (long long)p->n_chans * n_frames * band_layout(p).n_out;
    const int ch_count = (X || logspec) ? p->n_chans : 1;
    return launch_stft(p, tb, audio, reinterpret_cast<float2*>(X), p->upper_bin - p->lower_bin, 0, logspec, clip_stride,
                       power0, ch_count, (cudaStream_t)stream);
}

int salsa_tracker(const double* power0, uint32_t* mask, int32_t n_clips, int32_t n_frames, int32_t n_bins,
                  void* stream) {
    if (!power0 || !mask) return fail(SALSA_EINVAL, "power0 / mask is NULL");
    if (n_clips < 0 || n_frames <= 0 || n_bins <= 0) return fail(SALSA_EINVAL, "bad tracker dimensions");
    return launch_tracker(tracker_power0(power0, n_frames, n_bins), mask, n_clips, n_frames, 0, n_bins, (cudaStream_t)stream);
}

int salsa_spectrum_from_reference(const double* X_ref, float* X, double* power0, int32_t n_bins, int32_t n_frames,
                                  int32_t n_chans, void* stream) {
    if (!X_ref || !X) return fail(SALSA_EINVAL, "X_ref / X is NULL");
    if (n_chans != 4) return fail(SALSA_EINVAL, "n_chans must be 4");
    if (n_bins <= 0 || n_frames <= 0) return fail(SALSA_EINVAL, "bad spectrum dimensions");
    const long long total = (long long)n_bins * n_frames;
    from_reference_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(X_ref), reinterpret_cast<float2*>(X), power0, n_bins, n_frames);
    return check_launch("from_reference_kernel");
}

int salsa_eigenvector(const salsa_params_t* p, const float* X, const uint32_t* mask, float* eig, int32_t n_frames,
                      void* stream) {
    int rc = validate_params(p, false);
    if (rc) return rc;
    if (!X || !eig) return fail(SALSA_EINVAL, "X / eig is NULL");
    if (n_frames <= 0) return fail(SALSA_EINVAL, "n_frames must be positive");
    if (p->is_tracking && !mask) return fail(SALSA_EINVAL, "is_tracking needs the tracker mask");
    if (p->n_clips == 0) return SALSA_OK;
    const int n_bins = p->upper_bin - p->lower_bin;
    dim3 grid((n_bins + 255) / 256, n_frames, p->n_clips);
    eig_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(X),
                                                       p->is_tracking ? mask : nullptr, eig, n_frames, n_bins,
                                                       eig_args(p));
    return check_launch("eig_kernel");
}

size_t salsa_workspace_bytes(const salsa_params_t* p) {
    if (validate_params(p)) return 0;
    return carve_workspace(p, nullptr, pipeline_choice()).bytes;
}

int salsa_extract(const salsa_params_t* p, const float* audio, float* feature, void* workspace, size_t workspace_bytes,
                  void* stream) {
    int rc = validate_params(p);
    if (rc) return rc;
    if (p->n_clips == 0) return SALSA_OK;
    if (!audio || !feature) return fail(SALSA_EINVAL, "audio / feature is NULL");
    // the reference places (upper_bin - lower_bin) columns into a (3, T, freq_dim) array (:373-374): more bins than
    // columns is its broadcast error
    if (p->upper_bin - p->lower_bin > band_layout(p).n_out) return fail(SALSA_EINVAL, "upper_bin - lower_bin exceeds the feature width");
    const Pipeline pl = pipeline_choice();
    const Workspace w = carve_workspace(p, workspace, pl);
    if ((p->is_tracking || pl == kPipelineSplit) && (!workspace || workspace_bytes < w.bytes))
        return fail(SALSA_ENOMEM, "workspace smaller than salsa_workspace_bytes()");
    double win[kNfft], win_full[kNfft];
    const bool default_window = !p->window && p->win_len == p->n_fft;
    const bool builtin_hann = default_window && p->n_fft == kNfft;        // computed in the kernels, no table
    if (!builtin_hann) host_window(p, win);
    // win_len / window configure MagStftExtractor only (:324-325, :184-192); the spectrum that feeds the eigenvector step
    // is librosa's default full-length Hann whatever they are (:359-361)
    DeviceTables tb, tb_hann;
    if ((rc = get_tables(builtin_hann ? nullptr : win, &tb))) return rc;
    if (p->n_fft == kNfft) {
        if ((rc = get_tables(nullptr, &tb_hann))) return rc;
    } else {                                                               // n_fft = 256: the full-length Hann as a table too
        salsa_params_t ph = *p;
        ph.window = nullptr;
        ph.win_len = p->n_fft;
        host_window(&ph, win_full);
        if ((rc = get_tables(win_full, &tb_hann))) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_frames = salsa_n_frames(p->n_samples, p->hop_len);
    const int n_bins = p->upper_bin - p->lower_bin;
    const uint32_t* mask = nullptr;
    if (pl == kPipelineFused && !builtin_hann)
        return fail(SALSA_EINVAL, "SALSA_B200_PIPELINE=fused supports n_fft = 512 with the full-length Hann window only");
    if (pl == kPipelineSplit) {
        const long long clip_stride = 7LL * n_frames * band_layout(p).n_out;
        if (default_window) {
            if ((rc = launch_stft(p, tb, audio, w.X, 0, w.n_tiles, feature, clip_stride, nullptr, p->n_chans, st))) return rc;
        } else {
            if ((rc = launch_stft(p, tb, audio, nullptr, 0, 0, feature, clip_stride, nullptr, p->n_chans, st))) return rc;
            if ((rc = launch_stft(p, tb_hann, audio, w.X, 0, w.n_tiles, nullptr, 0, nullptr, p->n_chans, st))) return rc;
        }
        if (p->is_tracking) {
            // The tracker is a sequential recurrence over the whole clip in float64 on |X0|^2 of the complex64
            // spectrum (what the reference computes, :53-55); with stft_precision = 32 the selection follows that
            // spectrum, as the reference's would.
            const TrackerTiles src = {w.X, w.n_tiles, n_frames};
            if ((rc = launch_tracker(src, w.mask, p->n_clips, n_frames, 0, n_bins, st))) return rc;
            mask = w.mask;
        }
        return launch_eig_tile(p, w, mask, feature, st);
    }
    if (p->is_tracking) {
        // pass A: channel-0 spectrum in float64 -> tracker (a sequential recurrence over the whole clip,
        // so it has to finish before any bin can be selected).  Always float64: one flipped
        // comparison would shift the floor of that bin for the rest of the clip.
        salsa_params_t pa = *p;
        pa.stft_precision = 64;
        if ((rc = launch_stft(&pa, tb, audio, nullptr, 0, 0, nullptr, 0, w.power0, 1, st))) return rc;
        if ((rc = launch_tracker(tracker_power0(w.power0, n_frames, n_bins), w.mask, p->n_clips, n_frames, 0, n_bins, st))) return rc;
        mask = w.mask;
    }
    return launch_fused(p, tb, audio, feature, mask, st);
}

int salsa_lite_extract(const salsa_params_t* p, int32_t cutoff_bin, int32_t mode, const float* audio, float* feature,
                       void* stream) {
    int rc = validate_params(p);
    if (rc) return rc;
    if (mode != SALSA_LITE_NIPD && mode != SALSA_LITE_IPD) return fail(SALSA_EINVAL, "Invalid feature type");
    if (p->upper_bin > cutoff_bin || cutoff_bin > p->n_fft / 2)
        return fail(SALSA_EINVAL, "Upper bin for spatial feature is higher than cutoff bin for spectrogram!");
    if (p->n_clips == 0) return SALSA_OK;
    if (!audio || !feature) return fail(SALSA_EINVAL, "audio / feature is NULL");
    // the reference reads win_len from the config but never passes it on (salsa_lite_feature_extraction.py:44, :97-98):
    // every SALSA-Lite transform uses librosa's default full-length Hann, so win_len / window are ignored here too
    DeviceTables tb;
    if (p->n_fft == kNfft) {
        if ((rc = get_tables(nullptr, &tb))) return rc;
    } else {                                   // n_fft = 256: the full-length Hann as a table
        double win_full[kNfft];
        salsa_params_t ph = *p;
        ph.window = nullptr;
        ph.win_len = p->n_fft;
        host_window(&ph, win_full);
        if ((rc = get_tables(win_full, &tb))) return rc;
    }
    LiteArgs a;
    a.audio = audio;
    a.feature = feature;
    a.n_samples = p->n_samples;
    a.hop = p->hop_len;
    a.n_frames = salsa_n_frames(p->n_samples, p->hop_len);
    a.lower = p->lower_bin;
    a.cutoff = cutoff_bin;
    a.upper_cropped = p->upper_bin;     // applied to the CROPPED axis, as salsa_lite_feature_extraction.py:120 does
    a.mode = mode;
    a.inv_delta = eig_args(p).inv_delta;
    a.frames_per_block = 32;
    dim3 grid((a.n_frames + a.frames_per_block - 1) / a.frames_per_block, p->n_clips);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof("lite_kernel", st);
    // bins below 32 * NJ can carry a phase difference: cropped index < upper_cropped <=> bin < upper_bin + lower_bin
    const bool narrow = p->upper_bin + p->lower_bin <= 64;
    if (p->n_fft == 256) {
        const bool d = p->stft_precision == 64;
        const size_t smem = d ? sizeof(FftSmem<double>) : sizeof(FftSmem<float>);
        if ((rc = d ? set_smem(lite256_kernel<double>, smem) : set_smem(lite256_kernel<float>, smem))) return rc;
        if (d) lite256_kernel<double><<<grid, kThreads, smem, st>>>(a, tb.d);
        else lite256_kernel<float><<<grid, kThreads, smem, st>>>(a, tb.f);
        return check_launch("lite256_kernel");
    }
    if (p->stft_precision == 64) {
        const size_t smem = sizeof(FftSmem<double>);
        if (narrow) {
            if ((rc = set_smem(lite_kernel<double, 2>, smem))) return rc;
            lite_kernel<double, 2><<<grid, kThreads, smem, st>>>(a, tb.d);
        } else {
            if ((rc = set_smem(lite_kernel<double, 8>, smem))) return rc;
            lite_kernel<double, 8><<<grid, kThreads, smem, st>>>(a, tb.d);
        }
    } else {
        const size_t smem = sizeof(FftSmem<float>);
        if (narrow) {
            if ((rc = set_smem(lite_kernel<float, 2>, smem))) return rc;
            lite_kernel<float, 2><<<grid, kThreads, smem, st>>>(a, tb.f);
        } else {
            if ((rc = set_smem(lite_kernel<float, 8>, smem))) return rc;
            lite_kernel<float, 8><<<grid, kThreads, smem, st>>>(a, tb.f);
        }
    }
    return check_launch("lite_kernel");
}

// ---- LogSpecGccExtractor (dataset/feature_extraction.py:362-482) --------------------------------------------------
namespace {
// bf16 bits, round to nearest even
uint16_t bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float bf16_val(uint16_t b) {
    const uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
std::mutex g_gcc_mutex;
void* g_gcc_table[64] = {};          // per device: the inverse-transform table, bf16x2 planes [256][2][1024]

// row n = lag n - 100 (cc[-100:] then cc[:100], :439): cc[lag] = (1/N) (U0 + (-1)^lag U512 + 2 sum_k Re(U_k e^{2 pi i k lag / N}))
int gcc_table(void** out) {
    int dev = 0;
    SALSA_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_gcc_mutex);
    if (!g_gcc_table[dev]) {
        const int N = 1024;
        std::vector<uint16_t> h((size_t)kGccLagsPad * 2 * kGccK, 0);
        for (int n = 0; n < kGccLags; ++n) {
            const int lag = n - kGccLags / 2;
            for (int col = 0; col < kGccK; ++col) {
                double v;
                if (col == 0) v = 1.0 / N;
                else if (col == 1) v = ((lag & 1) ? -1.0 : 1.0) / N;
                else {
                    const int k = col >> 1;
                    const double ang = 2.0 * M_PI * (double)((long long)k * lag % N) / N;
                    v = (col & 1) ? -2.0 * sin(ang) / N : 2.0 * cos(ang) / N;
                }
                const uint16_t hi = bf16_rne((float)v);
                h[((size_t)n * 2) * kGccK + col] = hi;
                h[((size_t)n * 2 + 1) * kGccK + col] = bf16_rne((float)(v - (double)bf16_val(hi)));
            }
        }
        void* d = nullptr;
        SALSA_CUDA(cudaMalloc(&d, h.size() * sizeof(uint16_t)));
        SALSA_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        g_gcc_table[dev] = d;
    }
    *out = g_gcc_table[dev];
    return SALSA_OK;
}

struct GccWorkspace {
    float2 *XE, *XC, *XS;
    void* A;          // bf16 [rows_pad][2][1024]
    float* G;         // fp32 [rows_pad][256]
    size_t bytes;
};
GccWorkspace carve_gcc(const salsa_params_t* p, void* base) {
    const size_t T = (size_t)salsa_n_frames(p->n_samples, p->hop_len), B = (size_t)p->n_clips;
    const size_t rows = (B * 6 * T + 7) / 8 * 8;
    char* at = reinterpret_cast<char*>(base);
    GccWorkspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* ptr = at + off;
        off += round256(bytes);
        return ptr;
    };
    w.XE = reinterpret_cast<float2*>(take(B * T * 4 * (kHalf + 1) * sizeof(float2)));
    w.XC = reinterpret_cast<float2*>(take(B * T * 4 * kHalf * sizeof(float2)));
    w.XS = reinterpret_cast<float2*>(take(B * T * 4 * kHalf * sizeof(float2)));
    w.A = take(rows * 2 * kGccK * 2);
    w.G = reinterpret_cast<float*>(take(rows * kGccLagsPad * sizeof(float)));
    w.bytes = off + 256;
    return w;
}
}  // namespace

size_t salsa_logspec_gcc_workspace_bytes(const salsa_params_t* p) {
    if (!p || p->n_clips < 0 || p->hop_len <= 0) return 0;
    return carve_gcc(p, nullptr).bytes;
}

int salsa_logspec_gcc(const salsa_params_t* p, const float* audio, float* feature, void* workspace, size_t workspace_bytes, void* stream) {
    if (!p) return fail(SALSA_EINVAL, "params is NULL");
    salsa_params_t q = *p;
    q.lower_bin = 1;
    q.upper_bin = kHalf;
    q.audio_format = SALSA_FORMAT_MIC;
    int rc = validate_params(&q);
    if (rc) return rc;
    if (!q.is_compress_high_freq) return fail(SALSA_EINVAL, "logspec_gcc: only the compressed 200-band layout is implemented");
    if (q.window) return fail(SALSA_EINVAL, "logspec_gcc: a Hann window of win_len samples is built in");
    if (q.n_fft != kNfft || q.win_len != q.n_fft) return fail(SALSA_EINVAL, "logspec_gcc: win_len must equal n_fft (512)");
    if (q.n_clips == 0) return SALSA_OK;
    if (!audio || !feature) return fail(SALSA_EINVAL, "audio / feature is NULL");
    const GccWorkspace w = carve_gcc(&q, workspace);
    if (!workspace || workspace_bytes < w.bytes) return fail(SALSA_ENOMEM, "workspace smaller than salsa_logspec_gcc_workspace_bytes()");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_frames = salsa_n_frames(q.n_samples, q.hop_len);
    const BandLayout bands = band_layout(&q);
    // three transforms per (frame, channel): the Hann window w (even bins of the 1024-point spectrum, and the log-linear
    // spectrogram channels 0..3), w cos(pi j / 512) and w sin(pi j / 512) (odd bins)
    double win[3][kNfft];
    for (int j = 0; j < kNfft; ++j) {
        const double wj = 0.5 - 0.5 * cos(2.0 * M_PI * (double)j / kNfft);
        win[0][j] = wj;
        win[1][j] = wj * cos(M_PI * (double)j / kNfft);
        win[2][j] = wj * sin(M_PI * (double)j / kNfft);
    }
    salsa_params_t qe = q;
    qe.lower_bin = 0;
    qe.upper_bin = kHalf + 1;                  // bins 0 .. 256
    salsa_params_t qo = q;
    qo.lower_bin = 0;
    qo.upper_bin = kHalf;                      // bins 0 .. 255
    DeviceTables tb;
    if ((rc = get_tables(nullptr, &tb))) return rc;
    const long long spec_stride = 10LL * n_frames * bands.n_out;
    if ((rc = launch_stft(&qe, tb, audio, w.XE, kHalf + 1, 0, feature, spec_stride, nullptr, 4, st))) return rc;
    if ((rc = get_tables(win[1], &tb))) return rc;
    if ((rc = launch_stft(&qo, tb, audio, w.XC, kHalf, 0, nullptr, 0, nullptr, 4, st))) return rc;
    if ((rc = get_tables(win[2], &tb))) return rc;
    if ((rc = launch_stft(&qo, tb, audio, w.XS, kHalf, 0, nullptr, 0, nullptr, 4, st))) return rc;
    {
        ProfScope prof("gcc_unit_kernel", st);
        gcc_unit_kernel<<<dim3(n_frames, q.n_clips), 256, 0, st>>>(w.XE, w.XC, w.XS, reinterpret_cast<__nv_bfloat16*>(w.A), n_frames);
        if ((rc = check_launch("gcc_unit_kernel"))) return rc;
    }
    void* table = nullptr;
    if ((rc = gcc_table(&table))) return rc;
    const long long rows = (long long)q.n_clips * 6 * n_frames;
    if (rows > 0x7fffffffLL) return fail(SALSA_EINVAL, "logspec_gcc: too many rows for one call, split the batch");
    if ((rc = crnn_gemm(w.A, table, nullptr, nullptr, w.G, (int32_t)rows, kGccLagsPad, kGccK, 0, 2, stream))) return rc;
    ProfScope prof("gcc_scatter_kernel", st);
    gcc_scatter_kernel<<<dim3(n_frames, 6, q.n_clips), 256, 0, st>>>(w.G, feature, n_frames, bands.n_out);
    return check_launch("gcc_scatter_kernel");
}

size_t salsa_linspec_iv_workspace_bytes(const salsa_params_t* p) {
    if (!p || p->n_clips < 0 || p->hop_len <= 0) return 0;
    return round256((size_t)p->n_clips * salsa_n_frames(p->n_samples, p->hop_len) * 4 * (kHalf - 1) * sizeof(float2)) + 256;
}

int salsa_linspec_iv(const salsa_params_t* p, const float* audio, float* feature, void* workspace, size_t workspace_bytes, void* stream) {
    if (!p) return fail(SALSA_EINVAL, "params is NULL");
    salsa_params_t q = *p;
    q.lower_bin = 1;                       // every bin the band matrix touches: 1 .. n_fft/2 - 1
    q.upper_bin = kHalf;
    q.audio_format = SALSA_FORMAT_FOA;
    int rc = validate_params(&q);
    if (rc) return rc;
    if (q.n_fft != kNfft) return fail(SALSA_EINVAL, "linspec_iv: only n_fft = 512 is implemented");
    if (!q.is_compress_high_freq)
        return fail(SALSA_EINVAL, "linspec_iv: only the compressed 200-band layout is implemented (the uncompressed one needs the Nyquist bin)");
    if (q.n_clips == 0) return SALSA_OK;
    if (!audio || !feature) return fail(SALSA_EINVAL, "audio / feature is NULL");
    if (!workspace || workspace_bytes < salsa_linspec_iv_workspace_bytes(&q)) return fail(SALSA_ENOMEM, "workspace smaller than salsa_linspec_iv_workspace_bytes()");
    double win[kNfft];
    const bool builtin_hann = !q.window && q.win_len == q.n_fft;
    if (!builtin_hann) host_window(&q, win);
    DeviceTables tb;
    if ((rc = get_tables(builtin_hann ? nullptr : win, &tb))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_frames = salsa_n_frames(q.n_samples, q.hop_len);
    const BandLayout bands = band_layout(&q);
    float2* X = reinterpret_cast<float2*>(workspace);
    // one transform per (frame, channel): log-linear spectrogram rows into channels 0..3, the complex64 spectrum (one
    // window for both, as the reference: :325-340) into the workspace
    if ((rc = launch_stft(&q, tb, audio, X, kHalf - 1, 0, feature, 7LL * n_frames * bands.n_out, nullptr, 4, st))) return rc;
    ProfScope prof("iv_kernel", st);
    iv_kernel<<<dim3(n_frames, q.n_clips), 256, 0, st>>>(X, feature, n_frames, kHalf - 1, bands);
    return check_launch("iv_kernel");
}

int salsa_pcm16_to_float(const int16_t* pcm, float* audio, int64_t n, void* stream) {
    if (!pcm || !audio) return fail(SALSA_EINVAL, "pcm16_to_float: null pointer");
    if (n <= 0) return SALSA_OK;
    if ((reinterpret_cast<uintptr_t>(pcm) | reinterpret_cast<uintptr_t>(audio)) & 15) return fail(SALSA_EINVAL, "pcm16_to_float: pointers must be 16-byte aligned");
    const long long groups = (n + 7) / 8;
    const int blocks = (int)std::min<long long>((groups + 255) / 256, 148LL * 16);
    pcm16_to_float_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pcm, audio, n);
    return check_launch("pcm16_to_float_kernel");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// host-buffer variants: 3-stage pipeline over chunks of clips (H2D | kernels | D2H)
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kPipeBufs = 3;

// Streams, events and device staging buffers of the host-buffer entry points.  Cached per host thread
// and grown on demand: allocating them per call cost about 10 % of a 120-clip call.
struct HostPipeline {
    int device = -1;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t in_done[kPipeBufs] = {}, run_done[kPipeBufs] = {}, out_done[kPipeBufs] = {};
    float* d_audio[kPipeBufs] = {};
    float* d_feat[kPipeBufs] = {};
    int16_t* d_pcm[kPipeBufs] = {};     // 16-bit input staging (salsa_extract_host_pcm16)
    void* d_work = nullptr;
    size_t audio_bytes = 0, feat_bytes = 0, work_bytes = 0, pcm_bytes = 0;

    void release() {
        for (int i = 0; i < kPipeBufs; ++i) {
            if (d_audio[i]) cudaFree(d_audio[i]);
            if (d_feat[i]) cudaFree(d_feat[i]);
            if (d_pcm[i]) cudaFree(d_pcm[i]);
            if (in_done[i]) cudaEventDestroy(in_done[i]);
            if (run_done[i]) cudaEventDestroy(run_done[i]);
            if (out_done[i]) cudaEventDestroy(out_done[i]);
            d_audio[i] = d_feat[i] = nullptr;
            d_pcm[i] = nullptr;
            in_done[i] = run_done[i] = out_done[i] = nullptr;
        }
        if (d_work) cudaFree(d_work);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_out) cudaStreamDestroy(s_out);
        d_work = nullptr;
        s_in = s_run = s_out = nullptr;
        audio_bytes = feat_bytes = work_bytes = pcm_bytes = 0;
        device = -1;
    }
    int ensure(size_t ab, size_t fb, size_t wb, size_t pb) {
        int dev = 0;
        SALSA_CUDA(cudaGetDevice(&dev));
        if (dev != device) {
            release();
            device = dev;
            SALSA_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
            SALSA_CUDA(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
            SALSA_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
            for (int i = 0; i < kPipeBufs; ++i) {
                SALSA_CUDA(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
                SALSA_CUDA(cudaEventCreateWithFlags(&run_done[i], cudaEventDisableTiming));
                SALSA_CUDA(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
            }
        }
        if (ab > audio_bytes) {
            for (int i = 0; i < kPipeBufs; ++i) {
                if (d_audio[i]) cudaFree(d_audio[i]);
                d_audio[i] = nullptr;
                SALSA_CUDA(cudaMalloc((void**)&d_audio[i], ab));
            }
            audio_bytes = ab;
        }
        if (fb > feat_bytes) {
            for (int i = 0; i < kPipeBufs; ++i) {
                if (d_feat[i]) cudaFree(d_feat[i]);
                d_feat[i] = nullptr;
                SALSA_CUDA(cudaMalloc((void**)&d_feat[i], fb));
            }
            feat_bytes = fb;
        }
        if (pb > pcm_bytes) {
            for (int i = 0; i < kPipeBufs; ++i) {
                if (d_pcm[i]) cudaFree(d_pcm[i]);
                d_pcm[i] = nullptr;
                SALSA_CUDA(cudaMalloc((void**)&d_pcm[i], pb));
            }
            pcm_bytes = pb;
        }
        if (wb > work_bytes) {
            if (d_work) cudaFree(d_work);
            d_work = nullptr;
            SALSA_CUDA(cudaMalloc(&d_work, wb));
            work_bytes = wb;
        }
        return SALSA_OK;
    }
};
thread_local HostPipeline t_pipeline;

// In: float (the audio as librosa.load returns it) or int16_t (the 16-bit PCM of the wav files themselves,
// salsa_feature_extraction.py:353: half the host-to-device bytes; converted on the device, sample / 32768 as soundfile does)
template <typename In, typename Run>
int run_host_pipeline(const salsa_params_t* p, size_t feat_elems_per_clip, size_t work_bytes, const In* audio_host,
                      float* feature_host, int clips_per_chunk, Run run) {
    constexpr bool pcm = sizeof(In) == 2;
    if (p->n_clips == 0) return SALSA_OK;
    if (clips_per_chunk <= 0) clips_per_chunk = 16;
    clips_per_chunk = std::min(clips_per_chunk, p->n_clips);
    const size_t audio_elems = (size_t)p->n_chans * p->n_samples;
    HostPipeline& hp = t_pipeline;
    int rc0 = hp.ensure(clips_per_chunk * audio_elems * sizeof(float), clips_per_chunk * feat_elems_per_clip * sizeof(float), work_bytes,
                        pcm ? clips_per_chunk * audio_elems * sizeof(int16_t) : 0);
    if (rc0) return rc0;
    const int n_chunks = (p->n_clips + clips_per_chunk - 1) / clips_per_chunk;
    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c % kPipeBufs;
        const int first = c * clips_per_chunk;
        const int n = std::min(clips_per_chunk, p->n_clips - first);
        // the audio buffer is free once the kernels of chunk c - kPipeBufs are done, the feature buffer once
        // its copy-out is done (events of a previous call have completed: every call ends synchronised)
        if (c >= kPipeBufs) {
            SALSA_CUDA(cudaStreamWaitEvent(hp.s_in, hp.run_done[buf], 0));
            SALSA_CUDA(cudaStreamWaitEvent(hp.s_run, hp.out_done[buf], 0));
        }
        void* d_in = pcm ? (void*)hp.d_pcm[buf] : (void*)hp.d_audio[buf];
        SALSA_CUDA(cudaMemcpyAsync(d_in, audio_host + (size_t)first * audio_elems, (size_t)n * audio_elems * sizeof(In),
                                   cudaMemcpyHostToDevice, hp.s_in));
        SALSA_CUDA(cudaEventRecord(hp.in_done[buf], hp.s_in));
        SALSA_CUDA(cudaStreamWaitEvent(hp.s_run, hp.in_done[buf], 0));
        if (pcm) {
            int rcp = salsa_pcm16_to_float(hp.d_pcm[buf], hp.d_audio[buf], (int64_t)((size_t)n * audio_elems), (void*)hp.s_run);
            if (rcp) return rcp;
        }
        salsa_params_t pc = *p;
        pc.n_clips = n;
        int rc = run(&pc, hp.d_audio[buf], hp.d_feat[buf], hp.d_work, hp.work_bytes, hp.s_run);
        if (rc) {
            // copies of earlier chunks may still be in flight into the caller's buffers: every call ends synchronised
            cudaStreamSynchronize(hp.s_in);
            cudaStreamSynchronize(hp.s_run);
            cudaStreamSynchronize(hp.s_out);
            return rc;
        }
        SALSA_CUDA(cudaEventRecord(hp.run_done[buf], hp.s_run));
        SALSA_CUDA(cudaStreamWaitEvent(hp.s_out, hp.run_done[buf], 0));
        SALSA_CUDA(cudaMemcpyAsync(feature_host + (size_t)first * feat_elems_per_clip, hp.d_feat[buf],
                                   (size_t)n * feat_elems_per_clip * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out));
        SALSA_CUDA(cudaEventRecord(hp.out_done[buf], hp.s_out));
    }
    SALSA_CUDA(cudaStreamSynchronize(hp.s_out));
    SALSA_CUDA(cudaStreamSynchronize(hp.s_run));
    SALSA_CUDA(cudaStreamSynchronize(hp.s_in));
    return SALSA_OK;
}
}  // namespace

extern "C" {

int salsa_extract_host_pcm16(const salsa_params_t* p, const int16_t* audio_host, float* feature_host, int32_t clips_per_chunk) {
    int rc = validate_params(p);
    if (rc) return rc;
    if (!audio_host || !feature_host) return fail(SALSA_EINVAL, "audio / feature is NULL");
    const size_t n_frames = (size_t)salsa_n_frames(p->n_samples, p->hop_len);
    const size_t feat = 7 * n_frames * (size_t)band_layout(p).n_out;
    salsa_params_t pc = *p;
    pc.n_clips = std::min(p->n_clips, clips_per_chunk > 0 ? clips_per_chunk : 16);
    const size_t work = salsa_workspace_bytes(&pc);
    return run_host_pipeline(p, feat, work, audio_host, feature_host, clips_per_chunk,
                             [](const salsa_params_t* q, const float* a, float* f, void* w, size_t wb, cudaStream_t s) {
                                 return salsa_extract(q, a, f, w, wb, (void*)s);
                             });
}

int salsa_extract_host(const salsa_params_t* p, const float* audio_host, float* feature_host,
                       int32_t clips_per_chunk) {
    int rc = validate_params(p);
    if (rc) return rc;
    if (!audio_host || !feature_host) return fail(SALSA_EINVAL, "audio / feature is NULL");
    const size_t n_frames = (size_t)salsa_n_frames(p->n_samples, p->hop_len);
    const size_t feat = 7 * n_frames * (size_t)band_layout(p).n_out;
    salsa_params_t pc = *p;
    pc.n_clips = std::min(p->n_clips, clips_per_chunk > 0 ? clips_per_chunk : 16);
    const size_t work = salsa_workspace_bytes(&pc);
    return run_host_pipeline(p, feat, work, audio_host, feature_host, clips_per_chunk,
                             [](const salsa_params_t* q, const float* a, float* f, void* w, size_t wb, cudaStream_t s) {
                                 return salsa_extract(q, a, f, w, wb, (void*)s);
                             });
}

int salsa_scaler_accumulate(const float* feature, int32_t n_clips, int32_t n_feat_chans, int32_t n_frames, int32_t feat_dim,
                            double* sums, void* stream) {
    if (!feature || !sums) return fail(SALSA_EINVAL, "scaler: null pointer");
    if (n_clips < 0 || n_feat_chans < 4 || n_frames <= 0 || feat_dim <= 0 || feat_dim > 256)
        return fail(SALSA_EINVAL, "scaler: bad dimensions");
    if (n_clips == 0) return SALSA_OK;
    const int frames_per_block = 512;
    dim3 grid((n_frames + frames_per_block - 1) / frames_per_block, 4, n_clips);
    scaler_accumulate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feature, n_frames, feat_dim, n_feat_chans, frames_per_block, sums);
    return check_launch("scaler_accumulate_kernel");
}

int salsa_host_release(void) {
    t_pipeline.release();
    return SALSA_OK;
}

int salsa_lite_extract_host(const salsa_params_t* p, int32_t cutoff_bin, int32_t mode, const float* audio_host,
                            float* feature_host, int32_t clips_per_chunk) {
    int rc = validate_params(p);
    if (rc) return rc;
    if (!audio_host || !feature_host) return fail(SALSA_EINVAL, "audio / feature is NULL");
    if (cutoff_bin <= p->lower_bin) return fail(SALSA_EINVAL, "cutoff_bin must exceed lower_bin");
    const size_t n_frames = (size_t)salsa_n_frames(p->n_samples, p->hop_len);
    const size_t feat = 7 * n_frames * (size_t)(cutoff_bin - p->lower_bin);
    return run_host_pipeline(
        p, feat, 0, audio_host, feature_host, clips_per_chunk,
        [=](const salsa_params_t* q, const float* a, float* f, void*, size_t, cudaStream_t s) {
            return salsa_lite_extract(q, cutoff_bin, mode, a, f, (void*)s);
        });
}

}  // extern "C"
