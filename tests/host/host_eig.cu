// CPU build of the per-bin arithmetic of salsa_b200/csrc/eig.cuh (the functions are __host__ __device__), so that the
// certified coherence test and the eigenvector can be checked against the oracle over whole clips without a GPU.
// Test infrastructure: compiled on demand by tests/test_host_eig.py, never loaded by the product.
#include <stdint.h>
#include <string.h>

#include "eig.cuh"

using namespace salsa;

// X: complex64 [n_frames][4][pitch]; mask: [n_frames][(n_bins+31)/32] or NULL; out: [3][n_frames][n_bins] (zeroed here).
// stats[0] = bins evaluated, [1] = float32 verdict pass, [2] = fail, [3] = ambiguous (re-done in float64),
// [4] = ambiguous bins that pass in float64.
extern "C" int host_eig_clip(const float* Xf, const uint32_t* mask, int n_frames, int n_bins, int pitch, int format, int test,
                             double cond, int n_sq, int n_mv, int lower, double inv_delta, float* out, long long* stats) {
    const float2* X = reinterpret_cast<const float2*>(Xf);
    const int n_words = (n_bins + 31) / 32;
    memset(out, 0, sizeof(float) * 3 * (size_t)n_frames * n_bins);
    memset(stats, 0, sizeof(long long) * 5);
    EigArgs e;
    e.format = format;
    e.test = test;
    e.n_sq = n_sq;
    e.n_mv = n_mv;
    e.cond = (float)cond;
    e.cond_d = cond;
    e.inv_delta = inv_delta;
    e.lower = lower;
    for (int t = 0; t < n_frames; ++t)
        for (int b = 0; b < n_bins; ++b) {
            if (mask && !((mask[(size_t)t * n_words + (b >> 5)] >> (b & 31)) & 1u)) continue;
            const float2* fp[kWin];
            for (int k = 0; k < kWin; ++k) {
                int tt = (t - kHop + k) % n_frames;
                if (tt < 0) tt += n_frames;
                fp[k] = X + (size_t)tt * 4 * pitch + b;
            }
            auto load = [&](int k, int ch) -> float2 { return fp[k][ch * pitch]; };
            float o[3];
            int verdict = n_sq == 2 ? eig_bin_f32<2>(load, e, b, o) : eig_bin_f32<0>(load, e, b, o);
            stats[0]++;
            if (verdict == kEigAmbiguous) {
                stats[3]++;
                o[0] = o[1] = o[2] = 0.0f;
                verdict = eig_bin_f64(load, e, o, b);
                if (verdict == kEigPass) stats[4]++;
            } else {
                stats[verdict == kEigPass ? 1 : 2]++;
            }
            if (verdict == kEigPass)
                for (int i = 0; i < 3; ++i) out[((size_t)i * n_frames + t) * n_bins + b] = o[i];
        }
    return 0;
}
