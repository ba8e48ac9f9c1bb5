// Weight gradient of the 3x3 / pad 1 / stride 1 convolution on the Blackwell tensor cores (training step).
//
//   dW[tap][co][ci] = sum over pixels p of  dY[p][co] * X[p + offset(tap)][ci]          (zero padding outside the image)
//
// i.e. per tap a GEMM whose reduction runs over the PIXELS.  Both operands are NHWC activations, pixel-major in memory:
// exactly what tcgen05 calls "MN-major" operands (the M / N index is the contiguous one).  A TMA box of [64 channels] x
// [16 x 8 pixels] with 128-byte swizzle lands in shared memory as 128 rows (pixels = K) of 128 bytes (64 channels = M or N),
// which is the canonical MN-major SWIZZLE_128B layout: 8-row groups 1024 bytes apart (SBO), one MMA (K = 16) per two groups.
//
// One CTA owns one (64 output channels) x (64 input channels) block of all nine taps and a strided share of the pixel tiles
// (split-K over the grid's x dimension):
//   warp 0   TMA producer: per pixel tile the dY box and three column-shifted 18-row halo boxes of X (the same loads as
//            the forward kernel's: the nine taps are row offsets of 1024 bytes into them; out-of-image pixels arrive as zeros)
//   warp 1   MMA issuer: 9 taps x 8 tcgen05.mma (M = 64, N = 64, K = 16) per pixel tile into nine accumulators that live in
//            TMEM for the whole kernel: 64-row accumulators use 16 lanes of each TMEM quarter, so taps 2j and 2j+1 share the
//            columns 64 j .. 64 j + 63 at lane offsets 0 and 16 (5 x 64 = 320 of the 512 columns)
//   warp 2   TMEM allocation
//   warps 4-7 epilogue, once: tcgen05.ld -> red.global.add.f32 into dW (fp32 [9][Cout][Cin], zeroed by the host)
#pragma once
#include "crnn_conv.cuh"

namespace salsa {
namespace crnn {

constexpr int kWgStages = 3;
constexpr int kWgGyBytes = kTileH * kTileW * 128;                 // dY tile: 128 pixels x 64 channels
constexpr int kWgStageBytes = kWgGyBytes + 3 * kHaloBytes;
constexpr int kWgThreads = 256;
constexpr size_t kWgSmemBytes = 1024 + (size_t)kWgStages * kWgStageBytes + 256;

struct WgradArgs {
    int B, H, W, Cin, Cout;
    int tiles_w, tiles_h, n_ktiles;     // pixel tiles = B * tiles_h * tiles_w
    float* dw;                          // [9][Cout][Cin]
};

// shared-memory descriptor of an MN-major operand tile with 128-byte swizzle: rows (K) of 128 bytes, 8-row groups `sbo` apart
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t start, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((start >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;          // distance between 64-element groups along M / N (one group here)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                   // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands MN-major, M = 64, N = 64
__host__ __device__ constexpr uint32_t idesc_bf16_mn_m64_n64() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_gy, WgradArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
    uint64_t* full = bars;                      // [kWgStages]
    uint64_t* empty = full + kWgStages;         // [kWgStages]
    uint64_t* acc_full = empty + kWgStages;     // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((int)blockIdx.x >= a.n_ktiles) return;              // no pixel tile for this split (uniform for the CTA)
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_x);
        tc::prefetch_tmap(&tm_gy);
        for (int i = 0; i < kWgStages; ++i) {
            tc::mbar_init(full + i, 1);
            tc::mbar_init(empty + i, 1);
        }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512u);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int n_ci = a.Cin / 64;
    const int co0 = ((int)blockIdx.y / n_ci) * 64, ci0 = ((int)blockIdx.y % n_ci) * 64;

    if (warp == 0) {
        if (tc::elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int kt = blockIdx.x; kt < a.n_ktiles; kt += gridDim.x) {
                const int tw = kt % a.tiles_w, th = (kt / a.tiles_w) % a.tiles_h, b = kt / (a.tiles_w * a.tiles_h);
                const int h0 = th * kTileH, w0 = tw * kTileW;
                tc::mbar_wait(empty + s, ph ^ 1);
                tc::mbar_expect_tx(full + s, (uint32_t)kWgStageBytes);
                unsigned char* dst = smem + s * kWgStageBytes;
                tc::tma_load_4d(dst, &tm_gy, full + s, co0, w0, h0, b);
                for (int kw = 0; kw < 3; ++kw)
                    tc::tma_load_4d(dst + kWgGyBytes + kw * kHaloBytes, &tm_x, full + s, ci0, w0 - 1 + kw, h0 - 1, b);
                if (++s == kWgStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = idesc_bf16_mn_m64_n64();
            const uint64_t desc_fixed = smem_desc_mn_sw128(0, 1024, 1024);
            int s = 0;
            uint32_t ph = 0, accumulate = 0;
            for (int kt = blockIdx.x; kt < a.n_ktiles; kt += gridDim.x) {
                tc::mbar_wait(full + s, ph);
                tc::fence_after_sync();
                const uint32_t base = tc::smem_u32(smem + s * kWgStageBytes);
                const uint64_t a_desc = desc_fixed + (uint64_t)(base >> 4);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    // tap (dy, dx) = (t / 3, t % 3): X rows start dy * 8 pixels = dy * 1024 bytes into the halo box shifted by dx
                    const uint64_t b_desc = desc_fixed + (uint64_t)((base + kWgGyBytes + (t % 3) * kHaloBytes + (t / 3) * kTileW * 128) >> 4);
                    const uint32_t tmem_d = tmem_base + (uint32_t)((t >> 1) * 64) + ((uint32_t)((t & 1) * 16) << 16);
#pragma unroll
                    for (int k = 0; k < (kTileH * kTileW) / 16; ++k)       // 16 pixels = 16 rows of 128 bytes per MMA
                        tc::mma_bf16(tmem_d, a_desc + (uint64_t)(k * 128), b_desc + (uint64_t)(k * 128), idesc, (k == 0) ? accumulate : 1u);
                }
                accumulate = 1;
                tc::mma_commit(empty + s);
                if (++s == kWgStages) { s = 0; ph ^= 1; }
            }
            tc::mma_commit(acc_full);
        }
    } else if (warp >= 4) {
        const int q = warp - 4;                               // TMEM lane quarter: rows 16 q .. 16 q + 15 of every accumulator
        tc::mbar_wait(acc_full, 0);
        tc::fence_after_sync();
        const int row = co0 + 16 * q + (lane & 15);
#pragma unroll 1
        for (int j = 0; j < 5; ++j) {
            const int tap = 2 * j + (lane >> 4);              // lanes 0-15: tap 2j, lanes 16-31: tap 2j + 1
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tc::tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(j * 64 + half * 32), r);
                tc::tmem_ld_wait();
                if (tap < 9) {
                    float* dst = a.dw + ((size_t)tap * a.Cout + row) * a.Cin + ci0 + half * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(dst + i, __uint_as_float(r[i]));
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 512u);
}

}  // namespace crnn
}  // namespace salsa
