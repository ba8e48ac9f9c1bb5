// Weight gradient of the 3x3 / pad 1 / stride 1 convolution on the Blackwell tensor cores (training step).
//
//   dW[tap][co][ci] = sum over pixels p of  dY[p][co] * X[p + offset(tap)][ci]          (zero padding outside the image)
//                   = sum over pixels q of  dY[q - offset(tap)][co] * X[q][ci]
//
// i.e. per tap a GEMM whose reduction runs over the PIXELS.  Both operands are NHWC activations, pixel-major in memory:
// exactly what tcgen05 calls "MN-major" operands (the M / N index is the contiguous one).  A TMA box of [64 channels] x
// [16 x 8 pixels] with 128-byte swizzle lands in shared memory as 128 rows (pixels = K) of 128 bytes (64 channels = M or N),
// which is the canonical MN-major SWIZZLE_128B layout: 8-row groups 1024 bytes apart (SBO), one MMA (K = 16) per two groups.
//
// The halo is taken on dY (second form above): per pixel tile one plain X box and ONE 18-row x 16-column box of dY, in
// which tap (dy, dx) starts at box row 2 - dy, column 2 - dx (the start address is then not 1024-byte aligned: fine with the
// base-offset field 0, see crnn_conv.cuh CONV_SINGLE_HALO) and its 8-pixel rows are one box row = 2048 bytes apart.  A 64 x 64
// block of one tap would be an M = 64 MMA, which runs at half the tensor rate; an M = 128 MN-major operand is TWO 64-channel
// groups a "leading byte offset" apart, and nothing says the second group has to be other channels: with LBO = one box row it
// is the same 64 output channels one pixel row further down, i.e. the neighbouring tap.  So TWO TAPS are stacked into one
// M = 128 MMA (rows 0-63 of the accumulator = one tap, rows 64-127 = the other): (dy 2, dy 1) at each of the three columns,
// then (column 0, dy 0) with (column 1, dy 0) at LBO = one pixel = 128 bytes; the ninth tap stays an M = 64 MMA.  5 instead
// of 9 MMAs per 16 pixels at the same cycles each.
//
// One CTA owns one (64 output channels) x (64 input channels) block of all nine taps and a strided share of the pixel tiles
// (split-K over the grid's x dimension):
//   warp 0   TMA producer: per pixel tile the X box and the dY halo box (out-of-image pixels arrive as zeros)
//   warp 1   MMA issuer: 8 x (4 MMAs M = 128 + 1 MMA M = 64, N = 64, K = 16) per pixel tile into five accumulators that live
//            in TMEM for the whole kernel (4 x 64 columns on all 128 lanes + 64 columns on 16 lanes of each quarter)
//   warp 2   TMEM allocation
//   warps 4-7 epilogue, once: tcgen05.ld -> red.global.add.v4.f32 into dW (fp32 [9][Cout][Cin], zeroed by the host)
#pragma once
#include "crnn_conv.cuh"

namespace salsa {
namespace crnn {

constexpr int kWgStages = 4;
constexpr int kWgXBytes = kTileH * kTileW * 128;                  // X tile: 128 pixels x 64 channels
constexpr int kWgStageBytes = kWgXBytes + kHalo1Bytes;            // + the 18 x 16 halo box of dY
constexpr int kWgThreads = 256;
constexpr size_t kWgSmemBytes = 1024 + (size_t)kWgStages * kWgStageBytes + 256;

struct WgradArgs {
    int B, H, W, Cin, Cout;
    int tiles_w, tiles_h, n_ktiles;     // pixel tiles = B * tiles_h * tiles_w
    int taps;                           // 9 (3x3, pad 1) or 1 (1x1: only the centre tap = box row 1, column 1, as one M = 64 MMA)
    int cin_valid;                      // input channels that exist (Cin, or 16 for the padded first convolution: TMA zero-fills
                                        // channels 16..63 of the box, and only the first 16 columns of a dW row are written)
    float* dw;                          // [taps][Cout][Cin]
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

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands MN-major, N = 64
__host__ __device__ constexpr uint32_t idesc_bf16_mn_n64(uint32_t m) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
                 "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
                 : "memory");
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
    const int n_ci = (a.Cin + 63) / 64;
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
                tc::tma_load_4d(dst, &tm_x, full + s, ci0, w0, h0, b);
                tc::tma_load_4d(dst + kWgXBytes, &tm_gy, full + s, co0, w0 - 1, h0 - 1, b);
                if (++s == kWgStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc128 = idesc_bf16_mn_n64(128), idesc64 = idesc_bf16_mn_n64(64);
            constexpr uint32_t kRow = kHalo1Cols * 128;                                        // one row of the dY box (2048 bytes)
            const uint64_t desc_x = smem_desc_mn_sw128(0, 1024, 1024);                         // X tile: dense 16 x 8 pixels
            const uint64_t desc_rows = smem_desc_mn_sw128(0, kRow, kRow);                      // second M group: one pixel row down
            const uint64_t desc_cols = smem_desc_mn_sw128(0, 128, kRow);                       // second M group: the next column
            int s = 0;
            uint32_t ph = 0, accumulate = 0;
            for (int kt = blockIdx.x; kt < a.n_ktiles; kt += gridDim.x) {
                tc::mbar_wait(full + s, ph);
                tc::fence_after_sync();
                const uint32_t base = tc::smem_u32(smem + s * kWgStageBytes);
                const uint32_t gy = base + kWgXBytes;
                const uint64_t b_desc = desc_x + (uint64_t)(base >> 4);
                // 16 pixels per MMA = two tile rows: 2048 bytes of the X tile, two box rows of dY
                if (a.taps == 1) {
#pragma unroll
                    for (int k = 0; k < (kTileH * kTileW) / 16; ++k)
                        tc::mma_bf16(tmem_base + 256u, desc_rows + (uint64_t)((gy + kRow + 128) >> 4) + (uint64_t)(k * (2 * kRow >> 4)),
                                     b_desc + (uint64_t)(k * 128), idesc64, (k == 0) ? accumulate : 1u);
                } else
#pragma unroll
                for (int k = 0; k < (kTileH * kTileW) / 16; ++k) {
                    const uint32_t acc = (k == 0) ? accumulate : 1u;
                    const uint64_t ka = (uint64_t)(k * (2 * kRow >> 4)), kb = (uint64_t)(k * 128);
#pragma unroll
                    for (int sh = 0; sh < 3; ++sh)      // column sh (dx = 2 - sh): rows 0-63 <- dy = 2 (box row 0), rows 64-127 <- dy = 1
                        tc::mma_bf16(tmem_base + (uint32_t)(sh * 64), desc_rows + (uint64_t)((gy + sh * 128) >> 4) + ka, b_desc + kb, idesc128, acc);
                    // dy = 0 (box row 2): columns 0 and 1 stacked, column 2 alone
                    tc::mma_bf16(tmem_base + 192u, desc_cols + (uint64_t)((gy + 2 * kRow) >> 4) + ka, b_desc + kb, idesc128, acc);
                    tc::mma_bf16(tmem_base + 256u, desc_rows + (uint64_t)((gy + 2 * kRow + 2 * 128) >> 4) + ka, b_desc + kb, idesc64, acc);
                }
                accumulate = 1;
                tc::mma_commit(empty + s);
                if (++s == kWgStages) { s = 0; ph ^= 1; }
            }
            tc::mma_commit(acc_full);
        }
    } else if (warp >= 4) {
        const int q = warp - 4;                               // TMEM lane quarter
        tc::mbar_wait(acc_full, 0);
        tc::fence_after_sync();
        // M = 128 accumulators: lane = row; rows 0-63 / 64-127 are the two stacked taps
        const int row128 = co0 + ((32 * q + lane) & 63), upper = q >> 1;
#pragma unroll 1
        for (int j = 0; j < (a.taps == 9 ? 4 : 0); ++j) {
            // j < 3: column j, (dy 2 | dy 1), dx = 2 - j;   j = 3: dy 0, (column 0 -> dx 2 | column 1 -> dx 1)
            const int tap = j < 3 ? (upper ? 3 : 6) + (2 - j) : (upper ? 1 : 2);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tc::tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(j * 64 + half * 32), r);
                tc::tmem_ld_wait();
                float* dst = a.dw + ((size_t)tap * a.Cout + row128) * a.Cin + ci0 + half * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    if (half * 32 + i < a.cin_valid) red_add_v4(dst + i, r[i], r[i + 1], r[i + 2], r[i + 3]);
            }
        }
        // M = 64 accumulator (tap 0 = dy 0, dx 0; the only tap of a 1x1): rows 16 q .. 16 q + 15 on lanes 0-15 of this quarter
        const int row64 = co0 + 16 * q + (lane & 15);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(256 + half * 32), r);
            tc::tmem_ld_wait();
            if (lane < 16) {
                float* dst = a.dw + (size_t)row64 * a.Cin + ci0 + half * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    if (half * 32 + i < a.cin_valid) red_add_v4(dst + i, r[i], r[i + 1], r[i + 2], r[i + 3]);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 512u);
}

}  // namespace crnn
}  // namespace salsa
