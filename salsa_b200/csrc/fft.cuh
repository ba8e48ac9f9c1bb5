// Warp-level 512-point real FFT (one warp = one frame of one channel), sm_100a.
//
// Replaces the librosa.stft call sites of the reference
// (dataset/salsa_feature_extraction.py:186-192, :360-361; salsa_lite_feature_extraction.py:97-98):
// periodic Hann window (float64 in the reference), centre padding by reflection, rfft.
//
// The 512 real samples are packed as 256 complex points z[n] = x[2n] + i x[2n+1] and transformed
// with a 8 x 8 x 4 decimation: every lane owns 8 complex values, the two exchanges between the
// three butterfly passes go through a per-warp shared-memory scratch, and a final split (in
// registers, the mirrored partner comes by warp shuffle) turns Z[k] into the 257 bins of the real
// transform.  `T` is the arithmetic type: double reproduces the
// reference's float64 transform (results rounded to float32 afterwards, as librosa stores complex64),
// float is the fast variant.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace salsa {

constexpr int kNfft = 512;
constexpr int kHalf = 256;        // complex points of the packed transform
constexpr int kScratchPad = 36;   // row stride (complex elements) of the 8 x 32 buffer between pass 1 and 2
constexpr int kUStrideK = 39;     // strides of the [k1][m2][j1] buffer between pass 2 and 3
constexpr int kUStrideM = 10;
constexpr int kScratchElems = 8 * kUStrideK;   // complex elements of per-warp scratch (312)

template <typename T>
struct alignas(2 * sizeof(T)) Cx {
    T re, im;
};

template <typename T> __host__ __device__ __forceinline__ Cx<T> cadd(Cx<T> a, Cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <typename T> __host__ __device__ __forceinline__ Cx<T> csub(Cx<T> a, Cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <typename T> __host__ __device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
// multiply by -i (forward-transform quarter turn)
template <typename T> __host__ __device__ __forceinline__ Cx<T> mul_mi(Cx<T> a) { return {a.im, -a.re}; }

// In-place forward 8-point DFT: out[k] = sum_n in[n] exp(-2 pi i n k / 8).
template <typename T>
__device__ __forceinline__ void dft8(Cx<T> (&v)[8]) {
    const T h = (T)0.70710678118654752440084436210485;
    Cx<T> a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]);
    Cx<T> a2 = cadd(v[2], v[6]), a3 = mul_mi(csub(v[2], v[6]));
    Cx<T> a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
    Cx<T> a6 = cadd(v[3], v[7]), a7 = mul_mi(csub(v[3], v[7]));
    Cx<T> b0 = cadd(a0, a2), b2 = csub(a0, a2);          // even half, 4-point
    Cx<T> b1 = cadd(a1, a3), b3 = csub(a1, a3);
    Cx<T> c0 = cadd(a4, a6), c2 = mul_mi(csub(a4, a6));  // odd half, 4-point then W8 twiddles
    Cx<T> c1 = cadd(a5, a7), c3 = csub(a5, a7);
    c1 = {(c1.re + c1.im) * h, (c1.im - c1.re) * h};     // * exp(-i pi/4)
    c3 = {(c3.im - c3.re) * h, -(c3.re + c3.im) * h};    // * exp(-3 i pi/4)
    v[0] = cadd(b0, c0); v[4] = csub(b0, c0);
    v[1] = cadd(b1, c1); v[5] = csub(b1, c1);
    v[2] = cadd(b2, c2); v[6] = csub(b2, c2);
    v[3] = cadd(b3, c3); v[7] = csub(b3, c3);
}

// Forward 4-point DFT.
template <typename T>
__device__ __forceinline__ void dft4(Cx<T> (&v)[4]) {
    Cx<T> a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
    Cx<T> a2 = cadd(v[1], v[3]), a3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(a0, a2); v[2] = csub(a0, a2);
    v[1] = cadd(a1, a3); v[3] = csub(a1, a3);
}

// Twiddle tables, built once by the host in long double and rounded (salsa_abi.cu).  The kernels read
// three values per lane from them (lane_twiddles) and derive everything else in registers.
//   tw_a[k1][lane] = exp(-2 pi i * lane*k1 / 256)          k1 = 0..7   (row 1 is used)
//   tw_b[j1][lane] = exp(-2 pi i * (lane&3)*j1 / 32)       j1 = 0..7   (row 1 is used)
//   tw_r[k]        = exp(-2 pi i * k / 512)                k = 0..255  (k = lane is used)
template <typename T>
struct FftTables {
    const Cx<T>* tw_a;
    const Cx<T>* tw_b;
    const Cx<T>* tw_r;
    const T* window;   // n_fft entries
};

// Reflect-padded sample index (np.pad mode='reflect'): valid for -n < i < 2n-1.
__device__ __forceinline__ int reflect_index(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

// One warp: loads the 512 samples of the frame starting at sample `start` (in un-padded coordinates,
// may be negative / run past the end -> reflection) of channel signal `x` of length n.
// raw[n1] = (x[2m], x[2m+1]) for m = lane + 32 n1.  Split from the transform so that callers can
// issue the loads of the next frame before transforming the current one.
__device__ __forceinline__ void load_frame(const float* __restrict__ x, int n, int start, int lane, float2 (&raw)[8]) {
    const bool interior = (start >= 0) && (start + kNfft <= n) &&
                          ((reinterpret_cast<uintptr_t>(x + start) & 7) == 0);
    if (interior) {
        const float2* xp = reinterpret_cast<const float2*>(x + start);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) raw[n1] = __ldg(xp + lane + 32 * n1);
    } else {
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int m = lane + 32 * n1;
            raw[n1] = make_float2(x[reflect_index(start + 2 * m, n)], x[reflect_index(start + 2 * m + 1, n)]);
        }
    }
}

// Per-lane twiddle constants, loaded once per kernel and kept in registers.  Every other twiddle of
// the transform is a product of one of these with itself or with a compile-time constant: shared
// memory carries only the data exchanges (it is the pipe this transform is bound by).
template <typename T>
struct LaneTwiddles {
    Cx<T> w256;   // W256^lane          pass 1 -> 2
    Cx<T> w32;    // W32^(lane & 3)     pass 2 -> 3
    Cx<T> w512h;  // W512^lane / 2      real split (the 1/2 of the odd part folded in)
    // Optional tables in shared memory (null: the powers are multiplied out in registers).
    //   tab_a + 32 k1 = W256^(lane k1), k1 = 1..7: one conflict-free 16-byte load per power
    //   tab_b +  4 j1 = W32^((lane & 3) j1), j1 = 1..7: four distinct addresses inside one 64-byte line per load
    // Which of them are used is a template parameter of the transform (TAB: bit 0 = tab_a, bit 1 = tab_b).
    const Cx<T>* tab_a;
    const Cx<T>* tab_b;
};

constexpr int kTwiddleTabElems = 8 * 32 + 8 * 4;   // complex values of FftSmem's twiddle tables

template <typename T>
__device__ __forceinline__ LaneTwiddles<T> lane_twiddles(const FftTables<T>& tb, int lane) {
    const Cx<T> r = tb.tw_r[lane];
    return {tb.tw_a[32 + lane], tb.tw_b[32 + lane], {(T)0.5 * r.re, (T)0.5 * r.im}, nullptr, nullptr};
}

// Fills the shared-memory twiddle tables (`tab`: kTwiddleTabElems values; call by the whole CTA, barrier afterwards) ...
template <typename T>
__device__ __forceinline__ void load_twiddle_tables(Cx<T>* tab, const FftTables<T>& tb) {
    for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) tab[i] = tb.tw_a[i];                      // [k1][lane]
    for (int i = threadIdx.x; i < 8 * 4; i += blockDim.x) tab[8 * 32 + i] = tb.tw_b[(i >> 2) * 32 + (i & 3)];   // [j1][m2]
}
// ... and points a lane's constants at them.
template <typename T>
__device__ __forceinline__ void use_twiddle_tables(LaneTwiddles<T>& tw, const Cx<T>* tab, int lane) {
    tw.tab_a = tab + lane;
    tw.tab_b = tab + 8 * 32 + (lane & 3);
}

template <typename T> __device__ __forceinline__ Cx<T> shfl_cx(Cx<T> v, int src) {
    return {__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src)};
}

// One warp: z[m] = x[2m] w[2m] + i x[2m+1] w[2m+1], m = lane + 32 n1, from the samples in `raw` (see load_frame).
// `win` points to the window in shared memory (null: the periodic Hann window of length 512, computed from the lane
// constants).  Split from the passes so that callers can refill `raw` with the next frame as soon as it is consumed.
template <typename T>
__device__ __forceinline__ void window_frame(const float2 (&raw)[8], const T* __restrict__ win, const LaneTwiddles<T>& tw, int lane,
                                             Cx<T> (&v)[8]) {
    if (win) {
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int m = lane + 32 * n1;
            v[n1] = {(T)raw[n1].x * win[2 * m], (T)raw[n1].y * win[2 * m + 1]};
        }
    } else {
        // periodic Hann of length 512 without a table: w[n] = 1/2 - 1/2 cos(2 pi n / 512) = 1/2 - 1/2 Re W512^n, and
        // W512^(2m) = W256^m = w256 * W8^n1, W512^(2m+1) = W256^m * W512 (W8^n1 and W512 are constants)
        const T h = (T)0.70710678118654752440;
        const T c512 = (T)0.99992470183914454092, s512 = (T)0.01227153828571992608;    // cos, sin of 2 pi / 512
        const Cx<T> w8[8] = {{(T)1, (T)0}, {h, -h}, {(T)0, (T)-1}, {-h, -h}, {(T)-1, (T)0}, {-h, h}, {(T)0, (T)1}, {h, h}};
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const Cx<T> pw = n1 == 0 ? tw.w256 : cmul(tw.w256, w8[n1]);
            const T we = (T)0.5 - (T)0.5 * pw.re;
            const T wo = (T)0.5 - (T)0.5 * (pw.re * c512 + pw.im * s512);
            v[n1] = {(T)raw[n1].x * we, (T)raw[n1].y * wo};
        }
    }
}

// One warp: passes 1-3 of the 256-point complex transform of the windowed, packed frame in v (see window_frame).
// z[i] = Z[lane + 32 i].  `scratch` points to kScratchElems complex values of per-warp shared memory used by the two
// exchanges between the three butterfly passes.
template <typename T, int TAB>
__device__ __forceinline__ void warp_fft_core(Cx<T> (&v)[8], const LaneTwiddles<T>& tw, Cx<T>* scratch, int lane, Cx<T> (&z)[8]) {
    // ---- pass 1: 8-point DFT over n1 (stride 32), twiddle W256^(lane*k1) = w256^k1
    dft8(v);
    if (TAB & 1) {
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) v[k1] = cmul(v[k1], tw.tab_a[32 * k1]);
    } else {
        const Cx<T> w1 = tw.w256, w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
        v[1] = cmul(v[1], w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        v[4] = cmul(v[4], w4);
        v[5] = cmul(v[5], cmul(w4, w1));
        v[6] = cmul(v[6], cmul(w4, w2));
        v[7] = cmul(v[7], cmul(w4, w3));
    }
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) scratch[k1 * kScratchPad + lane] = v[k1];
    __syncwarp();
    // ---- pass 2: lane = (k1, m2); 8-point DFT over m1 of y[k1][4 m1 + m2], twiddle W32^(m2*j1) = w32^j1
    const int k1 = lane >> 2, m2 = lane & 3;
#pragma unroll
    for (int m1 = 0; m1 < 8; ++m1) v[m1] = scratch[k1 * kScratchPad + 4 * m1 + m2];
    __syncwarp();
    dft8(v);
    if (TAB & 2) {
#pragma unroll
        for (int j1 = 1; j1 < 8; ++j1) v[j1] = cmul(v[j1], tw.tab_b[4 * j1]);
    } else {
        const Cx<T> u1 = tw.w32, u2 = cmul(u1, u1), u3 = cmul(u2, u1), u4 = cmul(u2, u2);
        v[1] = cmul(v[1], u1);
        v[2] = cmul(v[2], u2);
        v[3] = cmul(v[3], u3);
        v[4] = cmul(v[4], u4);
        v[5] = cmul(v[5], cmul(u4, u1));
        v[6] = cmul(v[6], cmul(u4, u2));
        v[7] = cmul(v[7], cmul(u4, u3));
    }
    // u[k1][m2][j1] at k1*kUStrideK + m2*kUStrideM + j1  (strides chosen bank-conflict free)
#pragma unroll
    for (int j1 = 0; j1 < 8; ++j1) scratch[k1 * kUStrideK + m2 * kUStrideM + j1] = v[j1];
    __syncwarp();
    // ---- pass 3: lane owns (k1 = lane&7, j1 = (lane>>3) + 4p), p = 0,1; 4-point DFT over m2
    //      -> z[p + 2 j2] = Z[k1 + 8 j1 + 64 j2] = Z[lane + 32 (p + 2 j2)]
    const int pk1 = lane & 7;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int pj1 = (lane >> 3) + 4 * p;
        Cx<T> w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = scratch[pk1 * kUStrideK + q * kUStrideM + pj1];
        dft4(w);
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) z[p + 2 * j2] = w[j2];
    }
    __syncwarp();      // scratch may be reused by the caller / the next transform
}

// One warp: the 512-point real transform of the windowed, packed frame in v (see window_frame).
// X[j] = bin lane + 32 j (j = 0..7) of this lane; `nyq` = bin 256 (valid in lane 0 only).
template <typename T, int TAB = 0>
__device__ __forceinline__ void warp_fft_passes(Cx<T> (&v)[8], const LaneTwiddles<T>& tw, Cx<T>* scratch, int lane, Cx<T> (&X)[8],
                                                T& nyq) {
    Cx<T> z[8];
    warp_fft_core<T, TAB>(v, tw, scratch, lane, z);
    // ---- real split, in registers: X[k] = E + W512^k O with E = (Z[k] + conj Z[256-k]) / 2,
    //      O = -i (Z[k] - conj Z[256-k]) / 2.  For k = lane + 32 j the partner Z[256 - k] is z[7 - j] of lane
    //      32 - lane (lane 0: its own z[(8 - j) & 7]); W512^k = w512 * W16^j with W16^j a constant.  The factor 1/2
    //      of O lives in the twiddle (tw.w512h = W512^lane / 2), the one of E in the final multiply-add:
    //      10 instead of 14 operations per bin.
    nyq = z[0].re - z[0].im;
    const int partner_lane = (32 - lane) & 31;
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    const Cx<T> w16[8] = {{(T)1, (T)0}, {c1, -s1}, {h, -h}, {s1, -c1}, {(T)0, (T)-1}, {-s1, -c1}, {-h, -h}, {-c1, -s1}};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const Cx<T> other = shfl_cx(z[7 - j], partner_lane);
        const Cx<T> own = z[(8 - j) & 7];
        const Cx<T> a = z[j];
        const Cx<T> b = lane == 0 ? own : other;
        const T sr = a.re + b.re, si = a.im - b.im;          // 2 E
        const T dr = a.im + b.im, di = b.re - a.re;          // 2 O
        const Cx<T> wh = j == 0 ? tw.w512h : cmul(tw.w512h, w16[j]);
        const T tr = dr * wh.re - di * wh.im, ti = dr * wh.im + di * wh.re;
        X[j] = {(T)0.5 * sr + tr, (T)0.5 * si + ti};
    }
}

// The same transform with the real split done on PAIRS of bins: with E, O as above and T = W512^k O,
//     X[k] = E + T          X[256 - k] = conj(E - T),
// so one partner exchange and one twiddle product serve two bins (12 operations per pair instead of 2 x 10, half the
// shuffles).  The price is the order of the upper half of the spectrum:
//     lo[j] = X[lane + 32 j]                 j = 0..3   (bins 0..127, natural lane order)
//     hr[j] = X[256 - lane - 32 j]           j = 0..3   (lanes 1..31: bins 129..255 except 160, 192, 224; lane 0: bins
//                                                        256 (Nyquist), 224, 192, 160)
//     x128  = X[128]                         (valid in lane 0)
// upper_group() below turns (hr, x128) into "group g = bins 32 g .. 32 g + 31 in REVERSED lane order".
template <typename T, int TAB>
__device__ __forceinline__ void warp_fft_passes_paired(Cx<T> (&v)[8], const LaneTwiddles<T>& tw, Cx<T>* scratch, int lane,
                                                       Cx<T> (&lo)[4], Cx<T> (&hr)[4], Cx<T>& x128) {
    Cx<T> z[8];
    warp_fft_core<T, TAB>(v, tw, scratch, lane, z);
    x128 = {z[4].re, -z[4].im};                 // k = 128 pairs with itself: E = Re Z, T = -i Im Z
    const int partner_lane = (32 - lane) & 31;
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    const Cx<T> w16[4] = {{(T)1, (T)0}, {c1, -s1}, {h, -h}, {s1, -c1}};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // partner Z[256 - k] of k = lane + 32 j: z[7 - j] of lane 32 - lane; lane 0 (which reads from itself) pairs
        // with its own z[(8 - j) & 7]
        const Cx<T> send = lane == 0 ? z[(8 - j) & 7] : z[7 - j];
        const Cx<T> b = shfl_cx(send, partner_lane);
        const Cx<T> a = z[j];
        const T sr = a.re + b.re, si = a.im - b.im;          // 2 E
        const T dr = a.im + b.im, di = b.re - a.re;          // 2 O
        const Cx<T> wh = j == 0 ? tw.w512h : cmul(tw.w512h, w16[j]);
        const T tr = dr * wh.re - di * wh.im, ti = dr * wh.im + di * wh.re;
        lo[j] = {(T)0.5 * sr + tr, (T)0.5 * si + ti};
        hr[j] = {(T)0.5 * sr - tr, ti - (T)0.5 * si};
    }
}

// Value this lane holds of upper group g = 4..7 (bin 32 g + ((32 - lane) & 31)) given hr[] / x128 of
// warp_fft_passes_paired, already converted to whatever type V the caller works in.
template <typename V>
__device__ __forceinline__ V upper_group(const V (&hr)[4], V x128, int g, int lane) {
    const V first = g == 4 ? x128 : hr[(8 - g) & 3];       // lane 0: bins 128, 160, 192, 224
    return lane == 0 ? first : hr[7 - g];
}

// window + passes in one call
template <typename T>
__device__ __forceinline__ void warp_rfft512(const float2 (&raw)[8], const T* __restrict__ win, const LaneTwiddles<T>& tw,
                                             Cx<T>* scratch, int lane, Cx<T> (&X)[8], T& nyq) {
    Cx<T> v[8];
    window_frame<T>(raw, win, tw, lane, v);
    warp_fft_passes<T>(v, tw, scratch, lane, X, nyq);
}

// load + transform in one call
template <typename T>
__device__ __forceinline__ void warp_rfft512_frame(const float* __restrict__ x, int n, int start, const T* __restrict__ win,
                                                   const LaneTwiddles<T>& tw, Cx<T>* scratch, int lane, Cx<T> (&X)[8], T& nyq) {
    float2 raw[8];
    load_frame(x, n, start, lane, raw);
    warp_rfft512<T>(raw, win, tw, scratch, lane, X, nyq);
}

}  // namespace salsa
