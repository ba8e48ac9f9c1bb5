// Implicit-GEMM convolution (3x3 pad 1 / 1x1) and plain GEMM on the Blackwell tensor cores.
//
// Replaces the cuDNN / cuBLAS call sites of the reference CRNN: nn.Conv2d 3x3 of ConvBlock and
// _ResnetBasicBlock (models/model_utils.py:192-200, :301-304), the 1x1 downsample convolutions
// (:307-309), and the nn.Linear / GRU input projections of the decoder (models/decoders.py:44-46,
// :75-92) -- the last two as "1x1 convolutions over an image 8 pixels wide".
//
// Data layout: activations NHWC bf16 with C a multiple of 64; weights [tap][Cout][Cin] bf16 with the
// BatchNorm scale folded in; bias = folded BatchNorm shift (fp32).
//
// One CTA computes tiles of 16 x 8 = 128 output pixels (UMMA M) x N_TILE output channels (UMMA N),
// fp32 accumulators in TMEM (two stages, so the epilogue of tile i overlaps the MMAs of tile i+1):
//   warp 0     TMA producer.  Per 64-channel chunk of Cin it loads one halo box (18 rows x 16 pixels x
//              64 ch, 128-byte swizzle; out-of-image pixels are zero-filled by TMA = the convolution's
//              zero padding), from which all nine taps are addressed by a start offset of
//              (dy * 16 + dx) * 128 bytes (see CONV_SINGLE_HALO); per (chunk, tap) it loads the
//              [N_TILE][64] weight tile.
//   warp 1     MMA issuer: 4 x tcgen05.mma (K = 16) per (chunk, tap); tcgen05.commit releases the
//              shared-memory stages and finally publishes the accumulator.
//   warp 2     TMEM allocation.
//   warps 4-11 epilogue (two per TMEM lane quarter): tcgen05.ld -> + bias (+ residual) -> ReLU -> bf16 staging
//              tile -> TMA store (or direct fp32 / three-plane stores).
// Persistent: grid = min(tiles, SMs), static round-robin over tiles.
#pragma once
#include <cuda_bf16.h>

#include "tc_ptx.cuh"

namespace salsa {
namespace crnn {

constexpr int kTileH = 16, kTileW = 8;       // 128 output pixels per tile
constexpr int kKC = 64;                       // channels per K chunk = one 128-byte swizzle row
constexpr int kEpiWarps = 8;                  // two warps per TMEM lane quarter: each takes one 32-column half of a 64-channel group
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kConvThreads = 128 + kEpiThreads;
#ifndef CONV_A_STAGES
#define CONV_A_STAGES 2
#endif
constexpr int kAStages = CONV_A_STAGES;
constexpr int kHaloBytes = (kTileH + 2) * kTileW * 128;    // one column-shifted halo tile (3x3)
constexpr int kPlainBytes = kTileH * kTileW * 128;         // A tile of a 1x1 convolution / GEMM
// CONV_SINGLE_HALO = 2 (default): ONE halo box of 18 rows x 16 columns per 64-channel chunk; tap (dy, dx) starts
// (dy * 16 + dx) * 128 bytes into it with its 8-row groups (one tile row of 8 pixels each) 2048 bytes apart.  For dx > 0 that
// start address is not 1024-byte aligned -- which works because the tensor core applies the 128-byte swizzle to the ABSOLUTE
// shared-memory address bits (like TMA did when it wrote the box), so the descriptor's base-offset field stays 0.
// (= 1 puts the row phase dx into the base-offset field: wrong results, measured -- that was round 1's failed attempt.
//  = 0 is the earlier arrangement, three column-shifted 18 x 8 boxes whose tap offsets are multiples of 1024 bytes: 54 KB of TMA
//  fill per chunk instead of 36 KB; the 64-channel layers, bound by shared-memory bandwidth, run 5-9 % faster with one box.)
#ifndef CONV_SINGLE_HALO
#define CONV_SINGLE_HALO 2
#endif
constexpr int kHalo1Cols = 16;
constexpr int kHalo1Bytes = (kTileH + 2) * kHalo1Cols * 128;

struct ConvArgs {
    int B, H, W, Cin, Cout;
    int taps;                  // 9 (3x3, pad 1) or 1
    int planes;                // 1: bf16 operands.  3: every operand is a sum of three bf16 planes (fp32-grade products)
    int tiles_w, tiles_h;      // spatial tiles per image
    int n_tiles;               // B * tiles_h * tiles_w * (Cout / N_TILE)
    int relu;
    int pool;                  // 2x2 average pooling fused into the TMA-store epilogue: the output is [B][H/2][W/2][Cout]
    int tma_store;             // bf16 output leaves through a swizzled staging tile and TMA stores (planes == 1)
    int resident_b;            // all weight tiles stay in shared memory for the whole kernel (Cin = Cout = 64)
    long long pix_limit;       // pixels (rows of a GEMM) beyond this index are not written
    const float* bias;                 // [Cout] or null
    const __nv_bfloat16* residual;     // NHWC [B][H][W][Cout] or null
    __nv_bfloat16* out;                // NHWC bf16 (or null)
    float* out_f32;                    // NHWC fp32 (or null)
};

template <int N_TILE>
struct ConvSmem {
#ifndef CONV_B_STAGES_256
#define CONV_B_STAGES_256 4        // 4 x 32 KB: the single halo box freed 36 KB (3 stages before); >= 256-channel layers 5-8 % faster
#endif
#ifndef CONV_B_STAGES_128
#define CONV_B_STAGES_128 6
#endif
    static constexpr int kBStages = N_TILE >= 256 ? CONV_B_STAGES_256 : CONV_B_STAGES_128;
    static constexpr int kBBytes = N_TILE * 128;
    __host__ __device__ static constexpr int a_bytes(int taps) { return taps == 9 ? (CONV_SINGLE_HALO ? kHalo1Bytes : 3 * kHaloBytes) : kPlainBytes; }
    __host__ __device__ static constexpr int b_tiles(int taps, int resident) { return resident ? taps : kBStages; }
    static constexpr int kStgBufs = N_TILE >= 256 ? 1 : 2;          // output staging tiles (128 px x 64 ch bf16)
    static constexpr int kStgBytes = 128 * 128;
    static constexpr int kBiasBytes = 2048;                          // up to 512 output channels
    static constexpr size_t total(int taps, int resident) {
        return 1024 /* alignment slack */ + (size_t)kAStages * a_bytes(taps) + (size_t)b_tiles(taps, resident) * kBBytes +
               (size_t)kStgBufs * kStgBytes + kBiasBytes + 256 /* barriers */;
    }
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// Epilogue of one tile for one warp (32 output pixels): TMEM -> registers -> + bias (+ residual) -> ReLU ->
// fp32 and / or bf16 (one or three planes) NHWC stores.  `taddr0` = accumulator stage address of this warp's
// lane quarter; with planes > 1 the correction accumulator sits N_TILE columns further.
template <int N_TILE>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& a, uint32_t taddr0, int n0, size_t pix, bool valid) {
#pragma unroll 1
    for (int j = ((int)threadIdx.x - 128) >> 7; j < N_TILE / 32; j += 2) {      // the two warps of a lane quarter alternate
        uint32_t r[32];
        const uint32_t taddr = taddr0 + (uint32_t)(j * 32);
        tc::tmem_ld32(taddr, r);
        tc::tmem_ld_wait();
        if (a.planes > 1) {
            uint32_t r2[32];
            tc::tmem_ld32(taddr + (uint32_t)N_TILE, r2);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
        }
        if (valid) {
            const int n = n0 + j * 32;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + (a.bias ? __ldg(a.bias + n + i) : 0.0f);
            const size_t ostride = (size_t)a.planes * a.Cout;      // channels per pixel in memory
            if (a.residual) {
                for (int pl = 0; pl < a.planes; ++pl) {
                    const uint4* rp = reinterpret_cast<const uint4*>(a.residual + pix * ostride + (size_t)pl * a.Cout + n);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 u = __ldg(rp + i);
                        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[e]);
                            v[i * 8 + e * 2] += __low2float(h2);
                            v[i * 8 + e * 2 + 1] += __high2float(h2);
                        }
                    }
                }
            }
            if (a.relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
            if (a.out_f32) {
                float4* op = reinterpret_cast<float4*>(a.out_f32 + pix * a.Cout + n);
#pragma unroll
                for (int i = 0; i < 8; ++i) op[i] = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
            }
            if (a.out) {
                for (int pl = 0; pl < a.planes; ++pl) {
                    uint4* op = reinterpret_cast<uint4*>(a.out + pix * ostride + (size_t)pl * a.Cout + n);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                        pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                        v[2 * i] -= __low2float(h2);              // remainder goes to the next plane
                        v[2 * i + 1] -= __high2float(h2);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) op[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                }
            }
        }
    }
}

// Epilogue of one tile through shared memory and TMA (bf16 output, planes == 1): per 64-channel group the eight
// epilogue warps (two per TMEM lane quarter, one 32-column half each) convert their accumulator rows (+ bias from shared memory, + residual, ReLU) to bf16 and write
// them into a 128 x 64 staging tile in the 128-byte-swizzled layout of the output tensor map; one thread then
// issues a TMA store, which writes whole 128-byte lines and clips the tile at the image border.  Two staging
// tiles alternate (one for N_TILE = 256) so the store of group g overlaps the conversion of group g + 1.
//   `row`      tile row = output pixel of this thread (0..127), `stg_count` running group counter (selects the tile)
template <int N_TILE, int STG_BUFS>
__device__ __forceinline__ void epilogue_tile_tma(const ConvArgs& a, const CUtensorMap* tm_out, unsigned char* stg, const float* bias_s,
                                                  uint32_t taddr0, int n0, int b, int h0, int w0, size_t pix, bool valid, int row,
                                                  uint32_t& stg_count, uint64_t* acc_full_bar, uint32_t acc_parity, bool pingpong) {
    // Two ways to use the eight epilogue warps (two per TMEM lane quarter):
    //   split      both warp groups work on the same tile, group 0 on columns 0..31 and group 1 on columns 32..63 of every
    //              64-channel group (N_TILE = 256, where there is room for one staging tile only);
    //   ping-pong  the groups alternate tiles (group = tile parity = accumulator stage), each with its own staging tile and
    //              named barriers, so the fixed latencies of one tile's epilogue (accumulator wait, TMEM loads, proxy
    //              fence, barriers, TMA issue) overlap with the other group's.
    const int group = ((int)threadIdx.x - 128) >> 7;
    const int gtid = ((int)threadIdx.x - 128) & 127;
    const int n_thr = pingpong ? 128 : kEpiThreads;
    const int bar0 = pingpong ? 1 + 4 * group : 1;         // named barrier ids bar0 .. bar0 + 3
    const int h_begin = pingpong ? 0 : group, h_end = pingpong ? 2 : group + 1;
    const bool issuer = pingpong ? gtid == 0 : threadIdx.x == 128;
    const bool has_res = a.residual != nullptr && valid;
    // Residual values of the next 32-channel piece are requested one piece ahead -- the first one before the
    // accumulator is even complete -- so that their global-memory latency hides behind the MMAs / the previous piece.
    const uint4* res_base = reinterpret_cast<const uint4*>(a.residual + pix * a.Cout + n0);
    uint4 res_next[4];
    if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) res_next[i] = __ldg(res_base + h_begin * 4 + i);
    }
    tc::mbar_wait(acc_full_bar, acc_parity);
    tc::fence_after_sync();
#pragma unroll 1
    for (int g = 0; g < N_TILE / 64; ++g) {
        unsigned char* tile = stg + (pingpong ? group : (int)(stg_count % STG_BUFS)) * (128 * 128);
        if (pingpong && issuer) tc::tma_store_wait_read<0>();      // the previous store of this group has read `tile`
        tc::named_barrier(bar0, n_thr);                 // ... and (split mode) the issuer waited right after issuing it
#pragma unroll 1
        for (int half = h_begin; half < h_end; ++half) {
            uint32_t r[32];
            tc::tmem_ld32(taddr0 + (uint32_t)(g * 64 + half * 32), r);
            const int n = n0 + g * 64 + half * 32;
            uint4 res[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) res[i] = res_next[i];
            const int nh = half + 1 < h_end ? half + 1 : h_begin, ng = half + 1 < h_end ? g : g + 1;
            if (has_res && ng < N_TILE / 64) {
#pragma unroll
                for (int i = 0; i < 4; ++i) res_next[i] = __ldg(res_base + (ng * 2 + nh) * 4 + i);
            }
            tc::tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + (bias_s ? bias_s[n + i] : (a.bias ? __ldg(a.bias + n + i) : 0.0f));
            if (has_res) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t w4[4] = {res[i].x, res[i].y, res[i].z, res[i].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[e]);
                        v[i * 8 + e * 2] += __low2float(h2);
                        v[i * 8 + e * 2 + 1] += __high2float(h2);
                    }
                }
            }
            if (a.relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 q4 = make_uint4(pack_bf16(v[i * 8], v[i * 8 + 1]), pack_bf16(v[i * 8 + 2], v[i * 8 + 3]),
                                            pack_bf16(v[i * 8 + 4], v[i * 8 + 5]), pack_bf16(v[i * 8 + 6], v[i * 8 + 7]));
                const int chunk = (half * 4 + i) ^ (row & 7);                 // 128-byte swizzle: 16-byte chunk ^ (row mod 8)
                *reinterpret_cast<uint4*>(tile + row * 128 + chunk * 16) = q4;
            }
        }
        if (a.pool) {
            // F.avg_pool2d(kernel 2x2) of the finished tile, in place: 32 pooled pixels x 8 chunks = 256 tasks over the
            // group's threads.  Averages the bf16-rounded activations in fp32, exactly like avgpool2_kernel.
            tc::named_barrier(bar0 + 1, n_thr);
            const int my = pingpong ? gtid : (int)threadIdx.x - 128;
            const int n_task = 256 / n_thr;
            uint4 pooled[2];
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                if (s2 < n_task) {
                    const int task = my + n_thr * s2;
                    const int pp = task >> 3, c = task & 7;
                    const int r00 = (pp >> 2) * 16 + (pp & 3) * 2;
                    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        const int r = r00 + (d & 1) + (d >> 1) * 8;
                        const uint4 u = *reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) * 16));
                        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[e]);
                            acc[2 * e] += __low2float(h2);
                            acc[2 * e + 1] += __high2float(h2);
                        }
                    }
                    pooled[s2] = make_uint4(pack_bf16(0.25f * acc[0], 0.25f * acc[1]), pack_bf16(0.25f * acc[2], 0.25f * acc[3]),
                                            pack_bf16(0.25f * acc[4], 0.25f * acc[5]), pack_bf16(0.25f * acc[6], 0.25f * acc[7]));
                }
            }
            tc::named_barrier(bar0 + 2, n_thr);         // every read of the full-resolution tile is done
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                if (s2 < n_task) {
                    const int task = my + n_thr * s2;
                    const int pp = task >> 3, c = task & 7;
                    *reinterpret_cast<uint4*>(tile + pp * 128 + ((c ^ (pp & 7)) * 16)) = pooled[s2];
                }
            }
        }
        tc::fence_proxy_async();                        // generic-proxy writes -> visible to the TMA engine
        tc::named_barrier(a.pool ? bar0 + 3 : bar0 + 1, n_thr);
        if (issuer) {
            if (a.pool) tc::tma_store_4d(tm_out, tile, n0 + g * 64, w0 >> 1, h0 >> 1, b);
            else tc::tma_store_4d(tm_out, tile, n0 + g * 64, w0, h0, b);
            tc::tma_store_commit();
            if (!pingpong) tc::tma_store_wait_read<STG_BUFS - 1>();
        }
        ++stg_count;
    }
}

template <int N_TILE>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_wgt,
               const __grid_constant__ CUtensorMap tm_out, ConvArgs a) {
    using S = ConvSmem<N_TILE>;
    constexpr int NB = S::kBStages;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_bytes = S::a_bytes(a.taps);
    unsigned char* a_smem = smem;
    unsigned char* b_smem = smem + kAStages * a_bytes;
    unsigned char* stg_smem = b_smem + S::b_tiles(a.taps, a.resident_b) * S::kBBytes;      // [kStgBufs][128][128 B]
    float* bias_s = reinterpret_cast<float*>(stg_smem + S::kStgBufs * S::kStgBytes);        // [Cout]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_smem + S::kStgBufs * S::kStgBytes + S::kBiasBytes);
    uint64_t* a_full = bars;                  // [kAStages]
    uint64_t* a_empty = a_full + kAStages;    // [kAStages]
    uint64_t* b_full = a_empty + kAStages;    // [NB]
    uint64_t* b_empty = b_full + NB;          // [NB]
    uint64_t* acc_full = b_empty + NB;        // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ping-pong pays when the epilogue has a residual to fetch (measured: 0.28 -> 0.25 ms on the 64-channel residual
    // layers, 0.16 -> 0.20 ms on the plain ones), see epilogue_tile_tma
#ifndef CONV_PINGPONG_ALWAYS
#define CONV_PINGPONG_ALWAYS 0
#endif
    // ping-pong where the epilogue dominates: layers with a residual, and the 1x1 convolutions / GEMMs (4 MMAs per chunk;
    // measured 0.152 -> 0.135 ms on the 64 -> 128 downsample); the 3x3 layers without residual are faster split (0.338 vs 0.367)
    const bool pingpong = a.tma_store && S::kStgBufs == 2 && (CONV_PINGPONG_ALWAYS || a.residual != nullptr || a.taps == 1);
    if (a.Cout * (int)sizeof(float) > S::kBiasBytes) bias_s = nullptr;      // wide GEMMs read the bias from global memory
    if (a.tma_store && bias_s)
        for (int i = threadIdx.x; i < a.Cout; i += kConvThreads) bias_s[i] = a.bias ? a.bias[i] : 0.0f;
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_act);
        tc::prefetch_tmap(&tm_wgt);
        if (a.tma_store) tc::prefetch_tmap(&tm_out);
        for (int i = 0; i < kAStages; ++i) {
            tc::mbar_init(a_full + i, 1);
            tc::mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < NB; ++i) {
            tc::mbar_init(b_full + i, 1);
            tc::mbar_init(b_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(acc_full + i, 1);
            tc::mbar_init(acc_empty + i, pingpong ? kEpiWarps / 2 : kEpiWarps);
        }
        tc::fence_barrier_init();
    }
    // planes > 1 keeps the small correction products in their own accumulator (see the MMA issuer)
    const int acc_cols = a.planes > 1 ? 2 * N_TILE : N_TILE;      // TMEM columns per accumulator stage
    if (warp == 2) tc::tmem_alloc(tmem_slot, (uint32_t)(2 * acc_cols));
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int n_chunks = a.Cin / kKC;
    const int n_nt = a.Cout / N_TILE;
    auto decode = [&](int tile, int& n0, int& b, int& h0, int& w0) {
        const int nt = tile % n_nt;
        int sp = tile / n_nt;
        const int tw = sp % a.tiles_w;
        sp /= a.tiles_w;
        const int th = sp % a.tiles_h;
        b = sp / a.tiles_h;
        n0 = nt * N_TILE;
        h0 = th * kTileH;
        w0 = tw * kTileW;
    };

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (tc::elect_one()) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            if (a.resident_b) {     // every weight tile once, for the whole kernel
                tc::mbar_expect_tx(b_full, (uint32_t)(a.taps * S::kBBytes));
                for (int t = 0; t < a.taps; ++t) tc::tma_load_2d(b_smem + t * S::kBBytes, &tm_wgt, b_full, 0, t * a.Cout);
            }
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                int n0, b, h0, w0;
                decode(tile, n0, b, h0, w0);
                for (int c = 0; c < n_chunks; ++c) {
                    for (int ap = 0; ap < a.planes; ++ap) {
                        tc::mbar_wait(a_empty + sa, pa ^ 1);
                        tc::mbar_expect_tx(a_full + sa, (uint32_t)a_bytes);
                        unsigned char* dst = a_smem + sa * a_bytes;
                        const int ch0 = ap * a.Cin + c * kKC;
                        if (a.taps == 9 && CONV_SINGLE_HALO) {
                            tc::tma_load_4d(dst, &tm_act, a_full + sa, ch0, w0 - 1, h0 - 1, b);
                        } else if (a.taps == 9) {
                            for (int kw = 0; kw < 3; ++kw)
                                tc::tma_load_4d(dst + kw * kHaloBytes, &tm_act, a_full + sa, ch0, w0 - 1 + kw, h0 - 1, b);
                        } else {
                            tc::tma_load_4d(dst, &tm_act, a_full + sa, ch0, w0, h0, b);
                        }
                        if (++sa == kAStages) { sa = 0; pa ^= 1; }
                        // activation plane ap meets weight planes 0 .. planes-1-ap (products below 2^-24 are dropped)
                        for (int wp = 0; wp < a.planes - ap && !a.resident_b; ++wp) {
                            for (int t = 0; t < a.taps; ++t) {
                                tc::mbar_wait(b_empty + sb, pb ^ 1);
                                tc::mbar_expect_tx(b_full + sb, (uint32_t)S::kBBytes);
                                tc::tma_load_2d(b_smem + sb * S::kBBytes, &tm_wgt, b_full + sb, wp * a.Cin + c * kKC, t * a.Cout + n0);
                                if (++sb == NB) { sb = 0; pb ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_bf16_m128(N_TILE);
            int sa = 0, sb = 0, as = 0;
            uint32_t pa = 0, pb = 0, pacc = 0;
            if (a.resident_b) tc::mbar_wait(b_full, 0);
            // One thread issues every MMA of the CTA, so its instruction count per MMA is on the critical path:
            // descriptors are a constant (layout, SBO, version) plus the 16-byte-granular start address, and the
            // tap / K-step offsets are compile-time constants added to the stage's base descriptor.
            const uint64_t desc_fixed = tc::smem_desc_sw128(0, 1024);
            // A operand of a 3x3 tap in the single halo box: the 8-row groups (tile rows) are one box row = 2048 bytes apart
            const uint64_t desc_fixed_a = (a.taps == 9 && CONV_SINGLE_HALO) ? tc::smem_desc_sw128(0, kHalo1Cols * 128) : desc_fixed;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                tc::mbar_wait(acc_empty + as, pacc ^ 1);
                tc::fence_after_sync();
                // The tensor core's fp32 accumulation loses about 2^-24 of the accumulator per instruction, whatever
                // the size of the addend.  In bf16x3 mode the five correction products (2^-8 and below) therefore
                // go to a second accumulator, so that the main one sees one sixth of the instructions; the
                // epilogue adds the two.
                const uint32_t tmem_main = tmem_base + (uint32_t)(as * acc_cols);
                const uint32_t tmem_corr = tmem_main + (uint32_t)N_TILE;
                uint32_t acc_main = 0, acc_corr = 0;      // 0 until the first MMA into that accumulator
                for (int c = 0; c < n_chunks; ++c) {
                    for (int ap = 0; ap < a.planes; ++ap) {
                        tc::mbar_wait(a_full + sa, pa);
                        const uint64_t a_desc = desc_fixed_a + (uint64_t)(tc::smem_u32(a_smem + sa * a_bytes) >> 4);
                        for (int wp = 0; wp < a.planes - ap; ++wp) {
                            const bool main_acc = ap == 0 && wp == 0;
                            const uint32_t tmem_d = main_acc ? tmem_main : tmem_corr;
                            uint32_t accumulate = main_acc ? acc_main : acc_corr;
                            auto issue_tap = [&](int tap_off16) {     // tap_off16: A offset of the tap in 16-byte units
                                if (!a.resident_b) tc::mbar_wait(b_full + sb, pb);
                                tc::fence_after_sync();
                                const uint64_t b_desc = desc_fixed + (uint64_t)(tc::smem_u32(b_smem + sb * S::kBBytes) >> 4);
#pragma unroll
                                for (int k = 0; k < kKC / 16; ++k) {
                                    tc::mma_bf16(tmem_d, a_desc + (uint64_t)(tap_off16 + 2 * k), b_desc + (uint64_t)(2 * k), idesc, accumulate);
                                    accumulate = 1;
                                }
                                if (!a.resident_b) {
                                    tc::mma_commit(b_empty + sb);
                                    if (++sb == NB) { sb = 0; pb ^= 1; }
                                } else {
                                    ++sb;                             // resident: sb is just the tap index
                                }
                            };
                            if (a.taps == 9) {
                                if (a.resident_b) sb = 0;
#pragma unroll
                                for (int t = 0; t < 9; ++t)
                                    issue_tap(CONV_SINGLE_HALO ? (((t / 3) * kHalo1Cols + (t % 3)) * 128) >> 4
                                                               : ((t % 3) * kHaloBytes + (t / 3) * kTileW * 128) >> 4);
                            } else {
                                issue_tap(0);
                            }
                            if (main_acc) acc_main = 1; else acc_corr = 1;
                        }
                        tc::mma_commit(a_empty + sa);
                        if (++sa == kAStages) { sa = 0; pa ^= 1; }
                    }
                }
                tc::mma_commit(acc_full + as);
                if (++as == 2) { as = 0; pacc ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // =========================== epilogue ===========================
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;           // tile row = output pixel
        const int hl = row / kTileW, wl = row % kTileW;
        const int group = (warp - 4) >> 2;
        int as = 0, it = 0;
        uint32_t pacc = 0, stg_count = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            if (!pingpong || (it & 1) == group) {
                int n0, b, h0, w0;
                decode(tile, n0, b, h0, w0);
                const int h = h0 + hl, w = w0 + wl;
                const size_t pix = ((size_t)b * a.H + h) * a.W + w;
                const bool valid = h < a.H && w < a.W && (long long)pix < a.pix_limit;
                const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * acc_cols);
                if (a.tma_store) {
                    epilogue_tile_tma<N_TILE, S::kStgBufs>(a, &tm_out, stg_smem, bias_s, taddr0, n0, b, h0, w0, pix, valid, row, stg_count,
                                                           acc_full + as, pacc, pingpong);
                } else {
                    tc::mbar_wait(acc_full + as, pacc);
                    tc::fence_after_sync();
                    epilogue_tile<N_TILE>(a, taddr0, n0, pix, valid);
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(acc_empty + as);
            }
            if (++as == 2) { as = 0; pacc ^= 1; }
        }
        if (a.tma_store && (threadIdx.x == 128 || threadIdx.x == 256)) tc::tma_store_wait_read<0>();     // staging tiles must outlive their stores
    }
    // teardown
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, (uint32_t)(2 * acc_cols));
}

// ------------------------------------------------------------------------------------------------
// conv_first_kernel: the first convolution of the network (conv_block1.conv1, 7 -> 64 channels, 3x3).
// Its input has 7 channels, padded to 16 = ONE K = 16 step per tap, so the whole 3x3 x 16 filter bank
// (9 x 2 KB per plane) stays resident in shared memory and a tile costs 9 MMAs.  Same tiling, TMA halo
// trick (32-byte swizzle here: one pixel = one 32-byte row), accumulator staging and epilogue as
// conv_tc_kernel; the kernel is bound by writing its 64-channel output.
// ------------------------------------------------------------------------------------------------
constexpr int kC1 = 16;                                        // padded input channels
constexpr int kC1HaloBytes = (kTileH + 2) * kTileW * kC1 * 2;  // 4608: one column-shifted halo tile
constexpr int kC1WBytes = 64 * kC1 * 2;                        // 2048: weight tile of one tap
constexpr int kC1AStages = 4;

__host__ __device__ constexpr size_t conv_first_smem(int planes) {
    return 1024 + (size_t)planes * 9 * kC1WBytes + (size_t)kC1AStages * 3 * kC1HaloBytes + 1024 + 2 * 128 * 128 /* staging */ +
           256 /* bias */ + 256;
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_first_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_wgt,
                  const __grid_constant__ CUtensorMap tm_out, ConvArgs a) {
    constexpr int N_TILE = 64;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* w_smem = smem;                                         // [planes][9][64][16]
    unsigned char* a_smem = w_smem + a.planes * 9 * kC1WBytes;            // [stages][3][18][8][16]
    constexpr int a_bytes = 3 * kC1HaloBytes;
    // the A ring ends at a multiple of 512 bytes; the staging tiles need 1024-byte alignment for the 128-byte swizzle
    unsigned char* stg_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(a_smem + kC1AStages * a_bytes) + 1023) & ~(uintptr_t)1023);
    float* bias_s = reinterpret_cast<float*>(stg_smem + 2 * 128 * 128);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(bias_s) + 256);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kC1AStages;
    uint64_t* w_full = a_empty + kC1AStages;
    uint64_t* acc_full = w_full + 1;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifndef CONV_FIRST_PINGPONG
#define CONV_FIRST_PINGPONG 1
#endif
    // the two epilogue warp groups alternate tiles (1) or split every tile (0): the kernel is bound by its epilogue's fixed
    // latencies (accumulator wait, TMEM load, proxy fence, barriers, TMA issue), which ping-pong overlaps: 0.90 -> 0.66 ms per 16 clips
    const bool pingpong = CONV_FIRST_PINGPONG && a.tma_store;
    if (threadIdx.x < 64) bias_s[threadIdx.x] = a.bias ? a.bias[threadIdx.x] : 0.0f;
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_act);
        tc::prefetch_tmap(&tm_wgt);
        if (a.tma_store) tc::prefetch_tmap(&tm_out);
        for (int i = 0; i < kC1AStages; ++i) {
            tc::mbar_init(a_full + i, 1);
            tc::mbar_init(a_empty + i, 1);
        }
        tc::mbar_init(w_full, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(acc_full + i, 1);
            tc::mbar_init(acc_empty + i, pingpong ? kEpiWarps / 2 : kEpiWarps);
        }
        tc::fence_barrier_init();
    }
    const int acc_cols = a.planes > 1 ? 2 * N_TILE : N_TILE;
    if (warp == 2) tc::tmem_alloc(tmem_slot, (uint32_t)(2 * acc_cols));
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int tile, int& b, int& h0, int& w0) {
        const int tw = tile % a.tiles_w;
        tile /= a.tiles_w;
        const int th = tile % a.tiles_h;
        b = tile / a.tiles_h;
        h0 = th * kTileH;
        w0 = tw * kTileW;
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::mbar_expect_tx(w_full, (uint32_t)(a.planes * 9 * kC1WBytes));
            for (int wp = 0; wp < a.planes; ++wp)
                for (int t = 0; t < 9; ++t)
                    tc::tma_load_2d(w_smem + (wp * 9 + t) * kC1WBytes, &tm_wgt, w_full, wp * kC1, t * a.Cout);
            int sa = 0;
            uint32_t pa = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                int b, h0, w0;
                decode(tile, b, h0, w0);
                for (int ap = 0; ap < a.planes; ++ap) {
                    tc::mbar_wait(a_empty + sa, pa ^ 1);
                    tc::mbar_expect_tx(a_full + sa, (uint32_t)a_bytes);
                    for (int kw = 0; kw < 3; ++kw)
                        tc::tma_load_4d(a_smem + sa * a_bytes + kw * kC1HaloBytes, &tm_act, a_full + sa, ap * kC1, w0 - 1 + kw, h0 - 1, b);
                    if (++sa == kC1AStages) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_bf16_m128(N_TILE);
            int sa = 0, as = 0;
            uint32_t pa = 0, pacc = 0;
            tc::mbar_wait(w_full, 0);
            const uint64_t desc_fixed = tc::smem_desc_sw32(0, 256);
            const uint64_t w_desc = desc_fixed + (uint64_t)(tc::smem_u32(w_smem) >> 4);
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                tc::mbar_wait(acc_empty + as, pacc ^ 1);
                tc::fence_after_sync();
                const uint32_t tmem_main = tmem_base + (uint32_t)(as * acc_cols);
                const uint32_t tmem_corr = tmem_main + (uint32_t)N_TILE;
                uint32_t acc_main = 0, acc_corr = 0;
                for (int ap = 0; ap < a.planes; ++ap) {
                    tc::mbar_wait(a_full + sa, pa);
                    tc::fence_after_sync();
                    const uint64_t a_desc = desc_fixed + (uint64_t)(tc::smem_u32(a_smem + sa * a_bytes) >> 4);
                    for (int wp = 0; wp < a.planes - ap; ++wp) {
                        const bool main_acc = ap == 0 && wp == 0;
                        const uint32_t tmem_d = main_acc ? tmem_main : tmem_corr;
                        uint32_t accumulate = main_acc ? acc_main : acc_corr;
                        const uint64_t wp_desc = w_desc + (uint64_t)((wp * 9 * kC1WBytes) >> 4);
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            tc::mma_bf16(tmem_d, a_desc + (uint64_t)(((t % 3) * kC1HaloBytes + (t / 3) * kTileW * kC1 * 2) >> 4),
                                         wp_desc + (uint64_t)((t * kC1WBytes) >> 4), idesc, accumulate);
                            accumulate = 1;
                        }
                        if (main_acc) acc_main = 1; else acc_corr = 1;
                    }
                    tc::mma_commit(a_empty + sa);
                    if (++sa == kC1AStages) { sa = 0; pa ^= 1; }
                }
                tc::mma_commit(acc_full + as);
                if (++as == 2) { as = 0; pacc ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int hl = row / kTileW, wl = row % kTileW;
        const int group = (warp - 4) >> 2;
        int as = 0, it = 0;
        uint32_t pacc = 0, stg_count = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            if (!pingpong || (it & 1) == group) {
                int b, h0, w0;
                decode(tile, b, h0, w0);
                const int h = h0 + hl, w = w0 + wl;
                const size_t pix = ((size_t)b * a.H + h) * a.W + w;
                const bool valid = h < a.H && w < a.W;
                const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * acc_cols);
                if (a.tma_store) {
                    epilogue_tile_tma<N_TILE, 2>(a, &tm_out, stg_smem, bias_s, taddr0, 0, b, h0, w0, pix, valid, row, stg_count,
                                                 acc_full + as, pacc, pingpong);
                } else {
                    tc::mbar_wait(acc_full + as, pacc);
                    tc::fence_after_sync();
                    epilogue_tile<N_TILE>(a, taddr0, 0, pix, valid);
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(acc_empty + as);
            }
            if (++as == 2) { as = 0; pacc ^= 1; }
        }
        if (a.tma_store && (threadIdx.x == 128 || threadIdx.x == 256)) tc::tma_store_wait_read<0>();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, (uint32_t)(2 * acc_cols));
}

}  // namespace crnn
}  // namespace salsa
