// Element-wise / recurrent kernels around the tensor-core convolutions of the SELD CRNN.
//
//   pack_input_kernel     (B,7,T,F) fp32 NCHW -> (B,T,F,64) bf16 NHWC, zero-padded channels
//   avgpool2_kernel       F.avg_pool2d(kernel 2x2), floor mode   (model_utils.py:220, :349, :476)
//   freq_mean_kernel      torch.mean(x, dim=3)                   (decoders.py:111)
//   gru_layer_kernel      one bidirectional GRU layer, recurrent part (decoders.py:44-46, :126)
//   head_finish_kernel    split of the fused head GEMM, tanh on the DOA part (decoders.py:137-147)
//   gather_time_kernel    interpolate_tensor's index gather      (model_utils.py:57-75)
//   augment_kernel        training-time channel swaps + frequency shift of a feature batch (utilities/transforms.py)
//   seld_loss_*_kernel    BaseModel.compute_loss, reg_xyz (models/interfaces.py:273-355), and its gradient
#pragma once
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace salsa {
namespace crnn {

namespace cg = cooperative_groups;

// A float as the sum of up to three bf16 planes (the "bf16x3" precision mode): plane k holds the bf16
// rounding of what the planes before it left over, so hi + mid + lo reproduces 24 mantissa bits.
__device__ __forceinline__ void split_store8(const float (&v)[8], __nv_bfloat16* dst, int planes, size_t plane_stride) {
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = v[k];
    for (int pl = 0; pl < planes; ++pl) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            h[k] = __floats2bfloat162_rn(r[2 * k], r[2 * k + 1]);
            r[2 * k] -= __low2float(h[k]);
            r[2 * k + 1] -= __high2float(h[k]);
        }
        *reinterpret_cast<uint4*>(dst + pl * plane_stride) = *reinterpret_cast<uint4*>(h);
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int C, int T, int F,
                                  int T_use, int Cpad, int planes, const float* __restrict__ mean, const float* __restrict__ stdv,
                                  int n_scaled) {
    // one thread per output pixel; reads are coalesced along F, each thread writes Cpad bf16
    const long long n_pix = (long long)B * T_use * F;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(p % F);
        const long long bt = p / F;
        const int t = (int)(bt % T_use), b = (int)(bt / T_use);
        const float* src = x + ((long long)b * C * T + t) * F + f;
        __nv_bfloat16* dst = y + p * Cpad * planes;
        for (int c0 = 0; c0 < Cpad; c0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = c0 + i;
                v[i] = (c < C) ? __ldg(src + (long long)c * T * F) : 0.0f;
                // scaler normalisation of the data layer, fused: (x - mean) / std on the spectrogram channels
                // (database.py:196-202); a true float32 division, like the reference
                if (c < n_scaled) v[i] = (v[i] - __ldg(mean + c * F + f)) / __ldg(stdv + c * F + f);
            }
            split_store8(v, dst + c0, planes, Cpad);
        }
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __low2float(h[i]);
        f[2 * i + 1] = __high2float(h[i]);
    }
}

__global__ void avgpool2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H, int W, int C,
                                int planes) {
    const int CP = C * planes;       // channels per pixel in memory
    const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
    const long long total = (long long)B * Ho * Wo * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long p = i / C8;
        const int wo = (int)(p % Wo);
        p /= Wo;
        const int ho = (int)(p % Ho), b = (int)(p / Ho);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int dh = 0; dh < 2; ++dh)
#pragma unroll
            for (int dw = 0; dw < 2; ++dw) {
                const __nv_bfloat16* px = x + (((long long)b * H + 2 * ho + dh) * W + 2 * wo + dw) * CP + c8 * 8;
                for (int pl = 0; pl < planes; ++pl) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(px + pl * C));
                    float f[8];
                    bf16x8_to_float(u, f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += f[k];
                }
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= 0.25f;
        split_store8(acc, y + (((long long)b * Ho + ho) * Wo + wo) * CP + c8 * 8, planes, C);
    }
}

// (B,H,W,C) bf16 -> (B*H, C) bf16, mean over W
__global__ void freq_mean_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int BH, int W, int C,
                                 int planes) {
    const int CP = C * planes;
    const int C8 = C / 8;
    const long long total = (long long)BH * C8;
    const float inv = 1.0f / (float)W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const long long r = i / C8;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int w = 0; w < W; ++w) {
            for (int pl = 0; pl < planes; ++pl) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (r * W + w) * CP + pl * C + c8 * 8));
                float f[8];
                bf16x8_to_float(u, f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= inv;
        split_store8(acc, y + r * CP + c8 * 8, planes, C);
    }
}

// ------------------------------------------------------------------------------------------------
// gru_layer_kernel: recurrent part of one bidirectional GRU layer, hidden size 256 (nn.GRU gate
// order r, z, n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn))).
//
// A thread-block cluster of 8 CTAs owns one (direction, group of 8 clips): CTA c keeps the 96 rows of
// W_hh that belong to hidden units 32c .. 32c+31 (fp32, 96 KB of shared memory, loaded once) and the
// full hidden state of its 8 clips.  Per time step every CTA computes its 32 x 8 new hidden values
// and writes them into the shared memory of all 8 CTAs (DSMEM); one cluster barrier per step.
// The input projections W_ih x + b_ih come from the tensor-core GEMM (xproj, fp32).
// ------------------------------------------------------------------------------------------------
constexpr int kGruHidden = 256;
constexpr int kGruCluster = 8;
constexpr int kGruUnits = kGruHidden / kGruCluster;    // hidden units per CTA (32)
constexpr int kGruClips = 8;                            // clips per cluster
constexpr int kGruThreads = 256;

struct GruArgs {
    const float* xproj;        // [B*T][2*768]   (direction-major: fwd r,z,n | bwd r,z,n)
    const float* w_hh;         // [2][768][256]
    const float* b_hh;         // [2][768]
    __nv_bfloat16* y;          // [B*T][planes][512]  (fwd | bwd)
    int B, T, planes;
};

constexpr size_t kGruSmemBytes = (size_t)(3 * kGruUnits * kGruHidden + 2 * kGruHidden * kGruClips + 3 * kGruUnits * kGruClips) * sizeof(float);
constexpr size_t kGruStageBytes = (size_t)kGruUnits * kGruClips * sizeof(float);       // staging tile of the DSMEM broadcast (gru_layer_kernel)

__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(kGruThreads, 1) gru_layer_kernel(GruArgs a) {
    extern __shared__ __align__(16) float gsm[];
    float* wT = gsm;                                             // [256 k][96 rows]   (k-major: conflict-free)
    float* hbuf = wT + 3 * kGruUnits * kGruHidden;               // [2][256 k][8 clips]
    float* gates = hbuf + 2 * kGruHidden * kGruClips;            // [96 rows][8 clips]
    float* hstage = gates + 3 * kGruUnits * kGruClips;           // [32 own units][8 clips]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / kGruCluster;                    // cluster index
    const int dir = cid & 1, grp = cid >> 1;
    const int b0 = grp * kGruClips;
    const int tid = threadIdx.x;
    constexpr int R = 3 * kGruUnits;                             // 96 rows per CTA

    // W_hh slice: local row lr = g*32 + jl  <-  global row g*256 + rank*32 + jl
    const float* w = a.w_hh + (size_t)dir * 3 * kGruHidden * kGruHidden;
    for (int i = tid; i < R * kGruHidden; i += kGruThreads) {
        const int lr = i / kGruHidden, k = i - lr * kGruHidden;
        const int g = lr / kGruUnits, jl = lr - g * kGruUnits;
        wT[k * R + lr] = w[(size_t)(g * kGruHidden + rank * kGruUnits + jl) * kGruHidden + k];
    }
    for (int i = tid; i < 2 * kGruHidden * kGruClips; i += kGruThreads) hbuf[i] = 0.0f;
    cluster.sync();

    // matvec mapping: threads 0..191: row lr = tid % 96, clip half = tid / 96 (4 clips each)
    const int lr = tid % R, half = tid / R;
    // gate mapping: thread -> (unit jl = tid % 32, clip bl = tid / 32)
    const int jl = tid & 31, bl = tid >> 5;
    const int j = rank * kGruUnits + jl;
    const float* bh = a.b_hh + (size_t)dir * 3 * kGruHidden;
    const float b_hr = bh[j], b_hz = bh[kGruHidden + j], b_hn = bh[2 * kGruHidden + j];
    const int b = b0 + bl;
    const bool live = b < a.B;
    float h_prev = 0.0f;

    for (int s = 0; s < a.T; ++s) {
        const int t = dir ? a.T - 1 - s : s;
        const float* hc = hbuf + (s & 1) * kGruHidden * kGruClips;
        float* hn = hbuf + ((s + 1) & 1) * kGruHidden * kGruClips;
        // the three input projections of this thread's (unit, clip): issued early, used after the matvec
        float xr = 0.0f, xz = 0.0f, xn = 0.0f;
        if (live) {
            const float* xp = a.xproj + ((size_t)b * a.T + t) * (2 * 3 * kGruHidden) + dir * 3 * kGruHidden + j;
            xr = __ldg(xp);
            xz = __ldg(xp + kGruHidden);
            xn = __ldg(xp + 2 * kGruHidden);
        }
        if (tid < 2 * R) {
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const float4* h4 = reinterpret_cast<const float4*>(hc) + half;     // [k][2 halves] of float4
#pragma unroll 8
            for (int k = 0; k < kGruHidden; ++k) {
                const float wv = wT[k * R + lr];
                const float4 hv = h4[k * 2];
                acc[0] = fmaf(wv, hv.x, acc[0]);
                acc[1] = fmaf(wv, hv.y, acc[1]);
                acc[2] = fmaf(wv, hv.z, acc[2]);
                acc[3] = fmaf(wv, hv.w, acc[3]);
            }
            *reinterpret_cast<float4*>(gates + lr * kGruClips + half * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncthreads();
        {
            const float ar = gates[jl * kGruClips + bl];
            const float az = gates[(kGruUnits + jl) * kGruClips + bl];
            const float an = gates[(2 * kGruUnits + jl) * kGruClips + bl];
            const float r = 1.0f / (1.0f + expf(-(xr + ar + b_hr)));
            const float z = 1.0f / (1.0f + expf(-(xz + az + b_hz)));
            const float n = tanhf(xn + r * (an + b_hn));
            const float h_new = (1.0f - z) * n + z * h_prev;
            h_prev = h_new;
            hstage[jl * kGruClips + bl] = h_new;
            if (live) {
                __nv_bfloat16* yp = a.y + ((size_t)b * a.T + t) * (2 * kGruHidden) * a.planes + dir * kGruHidden + j;
                float rem = h_new;
                for (int pl = 0; pl < a.planes; ++pl) {
                    const __nv_bfloat16 hb = __float2bfloat16(rem);
                    yp[pl * 2 * kGruHidden] = hb;
                    rem -= __bfloat162float(hb);
                }
            }
        }
        __syncthreads();
        // publish to every CTA of the cluster (including this one): the CTA's 32 x 8 values are 1 KB contiguous in the
        // destination, sent by 64 threads as 16-byte stores (scalar stores from the (unit, clip) threads are 32 separate
        // sectors per warp; measured in the training kernels: 2.3x faster steps)
        if (tid < kGruUnits * kGruClips / 4) {
            const float4 v = reinterpret_cast<const float4*>(hstage)[tid];
#pragma unroll
            for (int c = 0; c < kGruCluster; ++c)
                reinterpret_cast<float4*>(cluster.map_shared_rank(hn, c))[rank * (kGruUnits * kGruClips / 4) + tid] = v;
        }
        cluster.sync();
    }
}

// ------------------------------------------------------------------------------------------------
// gru_layer_mma_kernel: the same recurrence with the per-step matrix product on the tensor cores
// (mma.sync m16n8k16, bf16 operands, fp32 accumulation) -- the fast ('bf16') mode.  Same cluster layout as
// gru_layer_kernel.  A CTA's 96 x 256 slice of W_hh lives in REGISTERS for the whole sequence as the A
// fragments of six warps (one 16-row tile each, 64 registers per lane); the hidden state of the 8 clips
// is the 256 x 8 B operand, kept in shared memory as bf16 [clip][k] and re-published through DSMEM every
// step; the fp32 hidden state of a (unit, clip) pair stays in the register of the thread that updates it.
// ------------------------------------------------------------------------------------------------
constexpr int kGruHPitch = kGruHidden + 8;      // bf16 elements per clip row of the B operand (bank-conflict-free)
constexpr size_t kGruMmaSmemBytes = (size_t)2 * kGruClips * kGruHPitch * sizeof(__nv_bfloat16) + (size_t)3 * kGruUnits * kGruClips * sizeof(float) +
                                    (size_t)kGruClips * kGruUnits * sizeof(__nv_bfloat16);      // + staging tile of the broadcast

__device__ __forceinline__ void mma_bf16_16x8x16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(kGruThreads, 1) gru_layer_mma_kernel(GruArgs a) {
    extern __shared__ __align__(16) unsigned char gsm_raw[];
    __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(gsm_raw);                       // [2][8 clips][kGruHPitch]
    float* gates = reinterpret_cast<float*>(hb + 2 * kGruClips * kGruHPitch);             // [96 rows][8 clips]
    __nv_bfloat16* hstage = reinterpret_cast<__nv_bfloat16*>(gates + 3 * kGruUnits * kGruClips);   // [8 clips][32 own units]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / kGruCluster;
    const int dir = cid & 1, grp = cid >> 1;
    const int b0c = grp * kGruClips;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;

    // A fragments of this warp's 16-row tile (warps 0..5): local rows 16*warp + {g, g+8}
    uint32_t afrag[kGruHidden / 16][4];
    if (warp < 6) {
        const float* w = a.w_hh + (size_t)dir * 3 * kGruHidden * kGruHidden;
        auto grow = [&](int lr) { return (size_t)((lr / kGruUnits) * kGruHidden + rank * kGruUnits + (lr % kGruUnits)) * kGruHidden; };
        const float* r0 = w + grow(16 * warp + g);
        const float* r1 = w + grow(16 * warp + g + 8);
#pragma unroll
        for (int ks = 0; ks < kGruHidden / 16; ++ks) {
            const int k = 16 * ks + 2 * t4;
            __nv_bfloat162 v;
            v = __floats2bfloat162_rn(r0[k], r0[k + 1]);     afrag[ks][0] = *reinterpret_cast<uint32_t*>(&v);
            v = __floats2bfloat162_rn(r1[k], r1[k + 1]);     afrag[ks][1] = *reinterpret_cast<uint32_t*>(&v);
            v = __floats2bfloat162_rn(r0[k + 8], r0[k + 9]); afrag[ks][2] = *reinterpret_cast<uint32_t*>(&v);
            v = __floats2bfloat162_rn(r1[k + 8], r1[k + 9]); afrag[ks][3] = *reinterpret_cast<uint32_t*>(&v);
        }
    }
    for (int i = tid; i < 2 * kGruClips * kGruHPitch; i += kGruThreads) hb[i] = __float2bfloat16(0.0f);
    cluster.sync();

    const int jl = tid & 31, bl = tid >> 5;               // gate mapping: (unit, clip)
    const int j = rank * kGruUnits + jl;
    const float* bh = a.b_hh + (size_t)dir * 3 * kGruHidden;
    const float b_hr = bh[j], b_hz = bh[kGruHidden + j], b_hn = bh[2 * kGruHidden + j];
    const int b = b0c + bl;
    const bool live = b < a.B;
    float h_prev = 0.0f;

    for (int s = 0; s < a.T; ++s) {
        const int t = dir ? a.T - 1 - s : s;
        const __nv_bfloat16* hc = hb + (s & 1) * kGruClips * kGruHPitch;
        __nv_bfloat16* hn = hb + ((s + 1) & 1) * kGruClips * kGruHPitch;
        float xr = 0.0f, xz = 0.0f, xn = 0.0f;
        if (live) {
            const float* xp = a.xproj + ((size_t)b * a.T + t) * (2 * 3 * kGruHidden) + dir * 3 * kGruHidden + j;
            xr = __ldg(xp);
            xz = __ldg(xp + kGruHidden);
            xn = __ldg(xp + 2 * kGruHidden);
        }
        if (warp < 6) {
            // two independent accumulators (even / odd k steps) halve the dependent MMA chain
            float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t* hrow = reinterpret_cast<const uint32_t*>(hc + g * kGruHPitch) + t4;       // clip g, k = 2*t4
#pragma unroll
            for (int ks = 0; ks < kGruHidden / 16; ks += 2) {
                mma_bf16_16x8x16(d0, afrag[ks], hrow[8 * ks], hrow[8 * ks + 4]);
                mma_bf16_16x8x16(d1, afrag[ks + 1], hrow[8 * ks + 8], hrow[8 * ks + 12]);
            }
            // C fragment: rows g / g+8 of the tile, clips 2*t4, 2*t4+1
            float* g0 = gates + (16 * warp + g) * kGruClips + 2 * t4;
            *reinterpret_cast<float2*>(g0) = make_float2(d0[0] + d1[0], d0[1] + d1[1]);
            *reinterpret_cast<float2*>(g0 + 8 * kGruClips) = make_float2(d0[2] + d1[2], d0[3] + d1[3]);
        }
        __syncthreads();
        {
            const float ar = gates[jl * kGruClips + bl];
            const float az = gates[(kGruUnits + jl) * kGruClips + bl];
            const float an = gates[(2 * kGruUnits + jl) * kGruClips + bl];
            const float r = 1.0f / (1.0f + __expf(-(xr + ar + b_hr)));
            const float z = 1.0f / (1.0f + __expf(-(xz + az + b_hz)));
            const float n = tanhf(xn + r * (an + b_hn));
            const float h_new = (1.0f - z) * n + z * h_prev;
            h_prev = h_new;
            const __nv_bfloat16 hb16 = __float2bfloat16(h_new);
            hstage[bl * kGruUnits + jl] = hb16;
            if (live) a.y[((size_t)b * a.T + t) * (2 * kGruHidden) + dir * kGruHidden + j] = hb16;
        }
        __syncthreads();
        // publish: one warp sends the CTA's 8 x 32 bf16 values as 16-byte stores (8 units each) to every CTA of the cluster
        if (tid < kGruClips * kGruUnits / 8) {
            const uint4 v = reinterpret_cast<const uint4*>(hstage)[tid];
            const int clip = tid >> 2, piece = tid & 3;
#pragma unroll
            for (int c = 0; c < kGruCluster; ++c)
                *reinterpret_cast<uint4*>(cluster.map_shared_rank(hn, c) + clip * kGruHPitch + rank * kGruUnits + piece * 8) = v;
        }
        cluster.sync();
    }
}

// ------------------------------------------------------------------------------------------------
// head GEMM output (rows, 64) fp32: cols 0..11 SED logits, 12..47 x|y|z before tanh
__global__ void head_finish_kernel(const float* __restrict__ z, float* __restrict__ logits, float* __restrict__ doa, int rows,
                                   int n_classes) {
    const int total = rows * 4 * n_classes;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / (4 * n_classes), c = i - r * 4 * n_classes;
        const float v = z[(size_t)r * 64 + c];
        if (c < n_classes) logits[(size_t)r * n_classes + c] = v;
        else doa[(size_t)r * 3 * n_classes + (c - n_classes)] = tanhf(v);
    }
}

// Output decoding of BaseModel.write_classwise_output_to_file (models/interfaces.py:224-246), reg_xyz format:
// active = sigmoid(logit) >= threshold; azimuth / elevation in whole degrees from the (x, y, z) regression,
// np.around (half to even) of the float32 degree value, azimuth 180 -> -180.
// One thread per (row, class); rows = clips x frames.
__global__ void decode_events_kernel(const float* __restrict__ logits, const float* __restrict__ doa, int rows, int n_classes,
                                     float threshold, uint8_t* __restrict__ active, int16_t* __restrict__ azi,
                                     int16_t* __restrict__ ele) {
    const int total = rows * n_classes;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / n_classes, c = i - r * n_classes;
        // torch.sigmoid in float32, then the comparison against a Python float (promoted to float64 by NumPy)
        const float prob = 1.0f / (1.0f + expf(-logits[i]));
        active[i] = (double)prob >= (double)threshold ? 1 : 0;
        const float x = doa[(size_t)r * 3 * n_classes + c];
        const float y = doa[(size_t)r * 3 * n_classes + n_classes + c];
        const float z = doa[(size_t)r * 3 * n_classes + 2 * n_classes + c];
        // float32 arithmetic like NumPy on float32 arrays; the transcendental itself is evaluated in float64 and rounded,
        // i.e. the correctly rounded float32 value
        const float a_rad = (float)atan2((double)y, (double)x);
        const float hyp = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));      // no FMA contraction: NumPy rounds each step
        const float e_rad = (float)atan2((double)z, (double)hyp);
        const float k180 = 180.0f, kpi = 3.14159265358979323846f;
        int a_deg = (int)rintf(__fdiv_rn(__fmul_rn(a_rad, k180), kpi));
        const int e_deg = (int)rintf(__fdiv_rn(__fmul_rn(e_rad, k180), kpi));
        if (a_deg == 180) a_deg = -180;
        azi[i] = (int16_t)a_deg;
        ele[i] = (int16_t)e_deg;
    }
}

// out[b][i][:] = in[b][idx[i]][:]
__global__ void gather_time_kernel(const float* __restrict__ in, const int* __restrict__ idx, float* __restrict__ out, int B,
                                   int n_in, int n_out, int width) {
    const long long total = (long long)B * n_out * width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % width);
        const long long r = i / width;
        const int o = (int)(r % n_out), b = (int)(r / n_out);
        out[i] = in[((long long)b * n_in + idx[o]) * width + c];
    }
}

// ------------------------------------------------------------------------------------------------
// Training-time augmentations that commute with the SALSA feature layout, one pass over the batch:
//   TfmapRandomSwapChannelFoa.apply  (utilities/transforms.py:394-437)   x: W Y Z X | Y Z X
//   TfmapRandomSwapChannelMic.apply  (:470-523)                          x: M1 M2 M3 M4 | p12 p13 p14
//   RandomShiftUpDownNp.apply        (:298-320, mode = 'reflect', all channels)
// The swaps act across the 7 channels of one (t, f) point, the shift is an index map along f, so out[:, t, f] is the
// swap of x[:, t, src(f)].  The float32 operations are the reference's, in its order (one subtraction per value at
// most per step): results are bit-identical.  ops[b] = {format, swap flags (bit i = m[i]), shift_len, direction}.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void swap_channels(float (&v)[7], int format, int m) {
    if (format == 0) {                      // FOA
        if (m & 1) {                        // swap x and y
            float t = v[1]; v[1] = v[3]; v[3] = t;
            t = v[4]; v[4] = v[6]; v[6] = t;
        }
        if (m & 2) v[6] = -v[6];
        if (m & 4) v[4] = -v[4];
        if (m & 8) v[5] = -v[5];
    } else {                                // MIC
        if (m & 1) {                        // swap M2 and M3
            float t = v[1]; v[1] = v[2]; v[2] = t;
            t = v[4]; v[4] = v[5]; v[5] = t;
        }
        if (m & 2) {                        // swap M1 and M4
            const float c4 = v[4], c5 = v[5], c6 = v[6];
            const float t = v[0]; v[0] = v[3]; v[3] = t;
            v[6] = -c6;
            v[5] = c5 - c6;
            v[4] = c4 - c6;
        }
        if (m & 4) {                        // swap M1 and M2, M3 and M4
            const float c4 = v[4], c5 = v[5], c6 = v[6];
            float t = v[0]; v[0] = v[1]; v[1] = t;
            t = v[2]; v[2] = v[3]; v[3] = t;
            v[4] = -c4;
            v[5] = c6 - c4;
            v[6] = c5 - c4;
        }
    }
}

__global__ void augment_kernel(const float* __restrict__ x, float* __restrict__ out, const int4* __restrict__ ops, int B, int T, int F) {
    const long long total = (long long)B * T * F;
    const long long plane = (long long)T * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const long long r = i / F;
        const int t = (int)(r % T), b = (int)(r / T);
        const int4 op = ops[b];
        int fs = f;
        if (op.z > 0) {
            // np.pad(..., mode='reflect') then crop: 'up' pads shift_len values in front, 'down' behind
            if (op.w == 0) fs = f < op.z ? op.z - f : f - op.z;
            else fs = f + op.z < F ? f + op.z : 2 * (F - 1) - (f + op.z);
        }
        const float* src = x + (long long)b * 7 * plane + (long long)t * F + fs;
        float v[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) v[c] = src[c * plane];
        swap_channels(v, op.x, op.y);
        float* dst = out + (long long)b * 7 * plane + (long long)t * F + f;
#pragma unroll
        for (int c = 0; c < 7; ++c) dst[c * plane] = v[c];
    }
}

// ------------------------------------------------------------------------------------------------
// Train-mode BatchNorm2d on NHWC bf16 activations (nn.BatchNorm2d with batch statistics, models/model_utils.py:202-203, :216,
// :356 in the training step), fused with what surrounds it in ConvBlock / _ResnetBasicBlock: the residual add and the ReLU.
//   forward   bn_stats_kernel (per-channel sum, sum of squares) -> bn_finalize_kernel (mean, 1/std, running statistics)
//             -> bn_apply_kernel: z = relu?(gamma (y - mean) / std + beta (+ residual))
//   backward  bn_bwd_reduce_kernel: g = dz * (z > 0) (the mask read from z, or recomputed from y when there was no residual);
//             dbeta = sum g, dgamma = sum g * xhat
//             -> bn_bwd_apply_kernel: dy = gamma / std * (g - dbeta / N - xhat * dgamma / N); the residual branch receives g
// Thread = 8 consecutive channels (one 16-byte load) of a strided set of pixels; C is a multiple of 8, at most 512.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[e]);
        v[2 * e] = __low2float(h2);
        v[2 * e + 1] = __high2float(h2);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint32_t w4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        w4[e] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    return make_uint4(w4[0], w4[1], w4[2], w4[3]);
}

// Block-level end of a per-channel reduction (thread = channel group threadIdx % groups, pixel lane threadIdx / groups, its
// 2 x 8 partial sums in a / b; s_part: 256 x 16 floats of shared memory).  Combine the pixel lanes of a channel group: every thread parks its 2 x 8 partial sums, then one thread per channel adds
// that channel's `lanes` values in float64 (one barrier; the loop-and-barrier-per-channel form cost ~8 us per launch,
// as much as streaming a 40 MB tensor)
__device__ __forceinline__ void channel_reduce_tail(const float (&a)[8], const float (&b)[8], bool active, int C, double* __restrict__ out,
                                                    float* s_part) {
    const int groups = C >> 3, lanes = blockDim.x / groups;
    {
        float* mine = s_part + threadIdx.x * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            mine[i] = active ? a[i] : 0.0f;
            mine[8 + i] = active ? b[i] : 0.0f;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c >> 3, i = c & 7;
        double sa = 0.0, sb = 0.0;
        for (int l = 0; l < lanes; ++l) {
            const float* p = s_part + (l * groups + g) * 16;
            sa += (double)p[i];
            sb += (double)p[8 + i];
        }
        atomicAdd(out + (size_t)c * 2, sa);
        atomicAdd(out + (size_t)c * 2 + 1, sb);
    }
}

// Per-channel reduction of two quantities over the pixels: every thread accumulates its channels over its pixels in fp32
// (at most a few hundred values), threads that share a channel group are combined through shared memory in fp64, one
// atomicAdd(double) per channel and block.  `load(pix, c8, q)` fetches the NT 16-byte words of pixel `pix`, channels
// 8 c8 .. 8 c8 + 7; `acc(q, a, b)` adds their contributions.  Four pixels' loads are issued before the first is consumed
// (the reductions are pure streaming reads: bytes in flight are what the bandwidth depends on).
template <int NT, typename L, typename A>
__device__ __forceinline__ void channel_reduce2(long long n_pix, int C, double* __restrict__ out /* [C][2] */, L load, A acc) {
    __shared__ float s_part[256 * 16];
    constexpr int U = 4;
    const int groups = C >> 3;                        // channel groups of 8
    const int lanes = blockDim.x / groups;            // threads per channel group (pixel lanes)
    const int c8 = threadIdx.x % groups, pl = threadIdx.x / groups;
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = b[i] = 0.0f;
    if (pl < lanes) {
        const long long stride = (long long)gridDim.x * lanes;
        long long p = (long long)blockIdx.x * lanes + pl;
        for (; p + (U - 1) * stride < n_pix; p += U * stride) {
            uint4 q[U][NT];
#pragma unroll
            for (int u = 0; u < U; ++u) load(p + u * stride, c8, q[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) acc(q[u], a, b);
        }
        for (; p < n_pix; p += stride) {
            uint4 q[NT];
            load(p, c8, q);
            acc(q, a, b);
        }
    }
    channel_reduce_tail(a, b, pl < lanes, C, out, s_part);
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ y, long long n_pix, int C, double* __restrict__ sums) {
    channel_reduce2<1>(n_pix, C, sums,
        [&](long long p, int c8, uint4 (&q)[1]) { q[0] = __ldg(reinterpret_cast<const uint4*>(y + p * C) + c8); },
        [&](const uint4 (&q)[1], float (&a)[8], float (&b)[8]) {
            float v[8];
            unpack8(q[0], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] += v[i];
                b[i] = fmaf(v[i], v[i], b[i]);
            }
        });
}

// sums [C][2] -> stat [C][2] = (mean, 1 / sqrt(var + eps)) with the biased batch variance; running statistics updated like
// nn.BatchNorm2d (momentum on the mean and on the UNBIASED variance)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long n_pix, int C, float eps, float momentum,
                                   float* __restrict__ stat, float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double n = (double)n_pix;
    const double mean = sums[2 * c] / n;
    const double var = fmax(sums[2 * c + 1] / n - mean * mean, 0.0);
    stat[2 * c] = (float)mean;
    stat[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(n > 1.0 ? var * n / (n - 1.0) : var);
    }
}

// Dropout fused behind the ReLU (nn.Dropout2d is NOT what the reference uses: models/model_utils.py:356 is element-wise
// nn.Dropout(p = 0.1) on relu(bn1(conv1(x)))).  The keep decisions are a pure function of (seed, salt, element index): 16 bits
// per element from two splitmix64 values per group of 8 channels, so the backward pass recomputes them instead of storing a
// mask.  The seed lives in DEVICE memory (a CUDA graph of the step replays with whatever the counter holds then).
struct DropArgs {
    const unsigned long long* seed;   // device scalar, or null: no dropout
    unsigned int salt;                // distinguishes the layers that share the seed
    unsigned int threshold;           // round(p * 65536): an element is dropped when its 16 bits are below it
    float scale;                      // 1 / (1 - p)
};
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// 32-bit finaliser (two multiplies): the per-element work of the dropout is four of these per group of 8 channels, cheap
// enough that the memory-bound BatchNorm kernels stay memory-bound (two splitmix64 per group cost 15 % of their time)
__device__ __forceinline__ unsigned int mix32(unsigned int x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    return x ^ (x >> 16);
}
// keep bits of the 8 channels of 16-byte group `i` (bit k = element k survives): 16 bits per element
__device__ __forceinline__ unsigned int drop_keep8(unsigned long long key, long long i, unsigned int threshold) {
    const unsigned int lo = (unsigned int)key, hi = (unsigned int)(key >> 32);
    const unsigned int base = ((unsigned int)i * 4u) ^ lo, salt = hi + (unsigned int)((unsigned long long)i >> 30) * 0x9E3779B9u;
    unsigned int keep = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const unsigned int h = mix32((base + (unsigned int)q) ^ salt);
        keep |= ((h & 0xffffu) >= threshold ? 1u : 0u) << (2 * q);
        keep |= ((h >> 16) >= threshold ? 1u : 0u) << (2 * q + 1);
    }
    return keep;
}
__device__ __forceinline__ unsigned long long drop_key(const DropArgs& d) {
    return splitmix64(*d.seed ^ ((unsigned long long)d.salt << 32));
}

// z = dropout?(relu?(scale y + shift (+ residual))) with scale = gamma / std, shift = beta - mean scale.  blockDim (256) is a
// multiple of the channel groups, so a thread keeps its 8 channels for the whole grid-stride loop: coefficients live in registers.
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ stat,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ z,
                                                       long long n_pix, int C, int relu, DropArgs drop) {
    const int groups = C >> 3;
    const long long total = n_pix * groups;
    const int c0 = (int)(threadIdx.x % groups) * 8;
    float scale[8], shift[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        scale[k] = gamma[c0 + k] * stat[2 * (c0 + k) + 1];
        shift[k] = fmaf(-stat[2 * (c0 + k)], scale[k], beta[c0 + k]);
    }
    const unsigned long long key = drop.seed ? drop_key(drop) : 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float v[8], r[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), v);
        if (residual) unpack8(__ldg(reinterpret_cast<const uint4*>(residual) + i), r);
        const unsigned int keep = drop.seed ? drop_keep8(key, i, drop.threshold) : 0xffu;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float o = fmaf(v[k], scale[k], shift[k]);
            if (residual) o += r[k];
            o = relu ? fmaxf(o, 0.0f) : o;
            if (drop.seed) o = ((keep >> k) & 1u) ? o * drop.scale : 0.0f;
            v[k] = o;
        }
        reinterpret_cast<uint4*>(z)[i] = pack8(v);
    }
}

// ReLU mask of the backward pass.  relu = 0: none; 1: z > 0 read from the forward output; 2: the forward had no residual, so
// z = relu(scale y + shift) is a function of y alone and the mask is RECOMPUTED from y with bn_apply_kernel's own expression
// (bit-identical to z > 0; one tensor less to read).
struct BnMaskCoef {
    float scale[8], shift[8];
};
__device__ __forceinline__ void bn_mask_coef(const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta,
                                             int c0, BnMaskCoef& m) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        m.scale[k] = gamma[c0 + k] * stat[2 * (c0 + k) + 1];
        m.shift[k] = fmaf(-stat[2 * (c0 + k)], m.scale[k], beta[c0 + k]);
    }
}
__device__ __forceinline__ bool bn_mask_from_y(float y, float scale, float shift) {
    // bf16(relu(o)) > 0  <=>  o rounds to a positive bf16: every positive float32 above half the smallest bf16 subnormal
    // (2^-134, bit pattern 0x00008000; the tie rounds to even = 0) does.  As signed integers the negative floats compare below.
    return __float_as_int(fmaf(y, scale, shift)) > 0x00008000;
}

// sums [C][2] += (sum g, sum g * xhat) with g = dz * mask, xhat = (y - mean) / std
template <int RELU>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ z,
                                                            const __nv_bfloat16* __restrict__ y, const float* __restrict__ stat,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            long long n_pix, int C, double* __restrict__ sums, DropArgs drop) {
    // a thread keeps its channel group: mean / (1 / std) of its 8 channels live in registers
    const int c0 = (int)(threadIdx.x % (C >> 3)) * 8;
    float mean[8], inv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mean[k] = stat[2 * (c0 + k)];
        inv[k] = stat[2 * (c0 + k) + 1];
    }
    BnMaskCoef mc;
    if (RELU == 2) bn_mask_coef(stat, gamma, beta, c0, mc);
    // dropout behind the ReLU: g = dz * keep / (1 - p).  With the mask read from z (RELU == 1) dropped elements have z = 0 and
    // fall out by themselves; otherwise the keep bits are recomputed.  The last word of q carries the group index for that.
    const unsigned long long key = drop.seed ? drop_key(drop) : 0ull;
    const float gscale = drop.seed ? drop.scale : 1.0f;
    const int groups = C >> 3;
    constexpr int NT = 3;
    channel_reduce2<NT>(n_pix, C, sums,
        [&](long long p, int c8, uint4 (&q)[NT]) {
            q[0] = __ldg(reinterpret_cast<const uint4*>(dz + p * C) + c8);
            q[1] = __ldg(reinterpret_cast<const uint4*>(y + p * C) + c8);
            if (RELU == 1) q[2] = __ldg(reinterpret_cast<const uint4*>(z + p * C) + c8);
            else {
                const unsigned long long gi = (unsigned long long)(p * groups + c8);
                q[2] = make_uint4((unsigned int)gi, (unsigned int)(gi >> 32), 0u, 0u);
            }
        },
        [&](const uint4 (&q)[NT], float (&a)[8], float (&b)[8]) {
            float g[8], zz[8], yy[8];
            unpack8(q[0], g);
            unpack8(q[1], yy);
            if (RELU == 1) unpack8(q[2], zz);
            unsigned int keep = 0xffu;
            if (RELU != 1 && drop.seed) keep = drop_keep8(key, (long long)(((unsigned long long)q[2].y << 32) | q[2].x), drop.threshold);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bool on = RELU == 0 ? true : (RELU == 1 ? zz[i] > 0.0f : bn_mask_from_y(yy[i], mc.scale[i], mc.shift[i]));
                on = on && ((keep >> i) & 1u);
                const float gi = on ? g[i] * gscale : 0.0f;
                a[i] += gi;
                b[i] = fmaf(gi, (yy[i] - mean[i]) * inv[i], b[i]);
            }
        });
}

// dy = gamma / std * (g - dbeta / N - xhat * dgamma / N) = ca g + cy y + c0 per channel; d_residual (optional) = g
template <int RELU>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ z,
                                                           const __nv_bfloat16* __restrict__ y, const float* __restrict__ stat,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const double* __restrict__ sums,
                                                           __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ d_residual,
                                                           long long n_pix, int C, DropArgs drop) {
    const int groups = C >> 3;
    const long long total = n_pix * groups;
    const float inv_n = (float)(1.0 / (double)n_pix);
    const int c0 = (int)(threadIdx.x % groups) * 8;
    float ca[8], cy[8], cc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        const float mean = stat[2 * c], inv = stat[2 * c + 1];
        const float dbeta = (float)sums[2 * c], dgamma = (float)sums[2 * c + 1];
        ca[k] = gamma[c] * inv;
        cy[k] = -ca[k] * dgamma * inv_n * inv;                  // coefficient of y
        cc[k] = -ca[k] * dbeta * inv_n - cy[k] * mean;          // constant
    }
    BnMaskCoef mc;
    if (RELU == 2) bn_mask_coef(stat, gamma, beta, c0, mc);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const unsigned long long key = drop.seed ? drop_key(drop) : 0ull;
    const float gscale = drop.seed ? drop.scale : 1.0f;
    auto one = [&](long long i, const uint4& qg, const uint4& qy, const uint4& qz) {
        float g[8], zz[8], yy[8], o[8];
        unpack8(qg, g);
        unpack8(qy, yy);
        if (RELU == 1) unpack8(qz, zz);
        const unsigned int keep = (RELU != 1 && drop.seed) ? drop_keep8(key, i, drop.threshold) : 0xffu;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bool on = RELU == 0 ? true : (RELU == 1 ? zz[k] > 0.0f : bn_mask_from_y(yy[k], mc.scale[k], mc.shift[k]));
            on = on && ((keep >> k) & 1u);
            g[k] = on ? g[k] * gscale : 0.0f;
            o[k] = fmaf(ca[k], g[k], fmaf(cy[k], yy[k], cc[k]));
        }
        reinterpret_cast<uint4*>(dy)[i] = pack8(o);
        if (d_residual) reinterpret_cast<uint4*>(d_residual)[i] = pack8(g);
    };
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < total; i += 2 * stride) {               // two positions' loads in flight
        const uint4 g0 = __ldg(reinterpret_cast<const uint4*>(dz) + i), g1 = __ldg(reinterpret_cast<const uint4*>(dz) + i + stride);
        const uint4 y0 = __ldg(reinterpret_cast<const uint4*>(y) + i), y1 = __ldg(reinterpret_cast<const uint4*>(y) + i + stride);
        uint4 z0 = make_uint4(0, 0, 0, 0), z1 = z0;
        if (RELU == 1) {
            z0 = __ldg(reinterpret_cast<const uint4*>(z) + i);
            z1 = __ldg(reinterpret_cast<const uint4*>(z) + i + stride);
        }
        one(i, g0, y0, z0);
        one(i + stride, g1, y1, z1);
    }
    if (i < total) {
        const uint4 g0 = __ldg(reinterpret_cast<const uint4*>(dz) + i), y0 = __ldg(reinterpret_cast<const uint4*>(y) + i);
        const uint4 z0 = RELU == 1 ? __ldg(reinterpret_cast<const uint4*>(z) + i) : make_uint4(0, 0, 0, 0);
        one(i, g0, y0, z0);
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm (+ residual) + ReLU + the 2x2 average pooling behind it in ONE pass each way.  At the four places where the
// network pools (after conv_block1 and at the entry of layers 2-4, models/model_utils.py:220, :349, :476) the BatchNorm
// output is consumed by the pooling alone, so the full-resolution activation z is never written: the forward writes the
// pooled tensor only, the backward reads the pooled gradient (dz = dpool / 4 under each window, 0 in an odd last row /
// column) and recomputes the ReLU mask from y (+ residual).  Same arithmetic, in the same order, as bn_apply_kernel followed
// by avgpool2_kernel / avgpool2_bwd_kernel followed by the bn_bwd kernels: bit-identical results.
// ------------------------------------------------------------------------------------------------
struct PoolGeom {
    int H, W, Ho, Wo;
};

// All three kernels walk the image ROW by row (a block takes rows blockIdx.x, + gridDim.x, ...; the threads of a block share a
// row's pixels and channel groups), so that the (batch, row, column) of a pixel costs one division per row instead of
// three per 16 bytes: the first, per-pixel form ran the reductions at 1.9 TB/s on integer arithmetic.

// a block takes POOLED rows; thread = (column, channel group) of that row, four input pixels each
__global__ void __launch_bounds__(256) bn_apply_pool_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ stat,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ pooled,
                                                            int B, PoolGeom g, int C) {
    const int groups = C >> 3, lg = 31 - __clz(groups);
    const int c8 = (int)threadIdx.x & (groups - 1);
    float scale[8], shift[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        scale[k] = gamma[c8 * 8 + k] * stat[2 * (c8 * 8 + k) + 1];
        shift[k] = fmaf(-stat[2 * (c8 * 8 + k)], scale[k], beta[c8 * 8 + k]);
    }
    for (int row = blockIdx.x; row < B * g.Ho; row += gridDim.x) {
        const int b = row / g.Ho, ho = row - b * g.Ho;
        const __nv_bfloat16* y0 = y + ((long long)b * g.H + 2 * ho) * g.W * C;
        const __nv_bfloat16* r0 = residual ? residual + ((long long)b * g.H + 2 * ho) * g.W * C : nullptr;
        uint4* out = reinterpret_cast<uint4*>(pooled + (long long)row * g.Wo * C);
        for (int e = threadIdx.x; e < g.Wo * groups; e += blockDim.x) {
            const int wo = e >> lg;
            uint4 qy[4], qr[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const long long off = ((long long)(d >> 1) * g.W + 2 * wo + (d & 1)) * C;
                qy[d] = __ldg(reinterpret_cast<const uint4*>(y0 + off) + c8);
                if (residual) qr[d] = __ldg(reinterpret_cast<const uint4*>(r0 + off) + c8);
            }
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                float v[8], r[8];
                unpack8(qy[d], v);
                if (residual) unpack8(qr[d], r);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float o = fmaf(v[k], scale[k], shift[k]);
                    if (residual) o += r[k];
                    acc[k] += __bfloat162float(__float2bfloat16_rn(fmaxf(o, 0.0f)));      // z as bn_apply_kernel would store it
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] *= 0.25f;
            out[e] = pack8(acc);
        }
    }
}

// dz word of a pixel from its window's pooled gradient word: a quarter of it, rounded as avgpool2_bwd_kernel stores it
__device__ __forceinline__ uint4 quarter_word(const uint4& u) {
    float v[8];
    unpack8(u, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= 0.25f;
    return pack8(v);
}

// RES = 0: mask from y (no residual in the forward); RES = 1: mask from y + residual
template <int RES>
__global__ void __launch_bounds__(256) bn_bwd_reduce_pool_kernel(const __nv_bfloat16* __restrict__ dpool, const __nv_bfloat16* __restrict__ y,
                                                                 const __nv_bfloat16* __restrict__ residual, const float* __restrict__ stat,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 int B, int C, PoolGeom geom, double* __restrict__ sums) {
    __shared__ float s_part[256 * 16];
    const int groups = C >> 3, lg = 31 - __clz(groups);
    const int c8 = (int)threadIdx.x & (groups - 1), c0 = c8 * 8;
    float mean[8], inv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mean[k] = stat[2 * (c0 + k)];
        inv[k] = stat[2 * (c0 + k) + 1];
    }
    BnMaskCoef mc;
    bn_mask_coef(stat, gamma, beta, c0, mc);
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = b[i] = 0.0f;
    const int n_e = geom.W * groups;
    for (int row = blockIdx.x; row < B * geom.H; row += gridDim.x) {
        const int bb = row / geom.H, h = row - bb * geom.H;
        const uint4* yrow = reinterpret_cast<const uint4*>(y + (long long)row * geom.W * C);
        const uint4* rrow = RES ? reinterpret_cast<const uint4*>(residual + (long long)row * geom.W * C) : nullptr;
        const bool in_h = (h >> 1) < geom.Ho;
        const __nv_bfloat16* drow = dpool + ((long long)bb * geom.Ho + (h >> 1)) * geom.Wo * C;
        for (int e0 = threadIdx.x; e0 < n_e; e0 += 2 * blockDim.x) {         // two positions' loads in flight
            uint4 qy[2], qr[2], qd[2];
            bool on[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = e0 + u * blockDim.x;
                on[u] = e < n_e;
                const int w = e >> lg;
                const bool win = on[u] && in_h && (w >> 1) < geom.Wo;
                qy[u] = on[u] ? __ldg(yrow + e) : make_uint4(0u, 0u, 0u, 0u);
                if (RES) qr[u] = on[u] ? __ldg(rrow + e) : make_uint4(0u, 0u, 0u, 0u);
                qd[u] = win ? __ldg(reinterpret_cast<const uint4*>(drow + (long long)(w >> 1) * C) + c8) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!on[u]) continue;
                float g[8], yy[8], rr[8];
                unpack8(quarter_word(qd[u]), g);
                unpack8(qy[u], yy);
                if (RES) unpack8(qr[u], rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float o = RES ? fmaf(yy[i], mc.scale[i], mc.shift[i]) + rr[i] : fmaf(yy[i], mc.scale[i], mc.shift[i]);
                    const float gi = __float_as_int(o) > 0x00008000 ? g[i] : 0.0f;
                    a[i] += gi;
                    b[i] = fmaf(gi, (yy[i] - mean[i]) * inv[i], b[i]);
                }
            }
        }
    }
    channel_reduce_tail(a, b, true, C, sums, s_part);
}

template <int RES>
__global__ void __launch_bounds__(256) bn_bwd_apply_pool_kernel(const __nv_bfloat16* __restrict__ dpool, const __nv_bfloat16* __restrict__ y,
                                                                const __nv_bfloat16* __restrict__ residual, const float* __restrict__ stat,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const double* __restrict__ sums, __nv_bfloat16* __restrict__ dy,
                                                                __nv_bfloat16* __restrict__ d_residual, int B, int C, PoolGeom geom) {
    const int groups = C >> 3, lg = 31 - __clz(groups);
    const float inv_n = (float)(1.0 / ((double)B * geom.H * geom.W));
    const int c8 = (int)threadIdx.x & (groups - 1), c0 = c8 * 8;
    float ca[8], cy[8], cc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        const float mean = stat[2 * c], inv = stat[2 * c + 1];
        const float dbeta = (float)sums[2 * c], dgamma = (float)sums[2 * c + 1];
        ca[k] = gamma[c] * inv;
        cy[k] = -ca[k] * dgamma * inv_n * inv;
        cc[k] = -ca[k] * dbeta * inv_n - cy[k] * mean;
    }
    BnMaskCoef mc;
    bn_mask_coef(stat, gamma, beta, c0, mc);
    const int n_e = geom.W * groups;
    for (int row = blockIdx.x; row < B * geom.H; row += gridDim.x) {
        const int bb = row / geom.H, h = row - bb * geom.H;
        const long long base = (long long)row * geom.W * C;
        const uint4* yrow = reinterpret_cast<const uint4*>(y + base);
        const uint4* rrow = RES ? reinterpret_cast<const uint4*>(residual + base) : nullptr;
        uint4* dyrow = reinterpret_cast<uint4*>(dy + base);
        uint4* drrow = d_residual ? reinterpret_cast<uint4*>(d_residual + base) : nullptr;
        const bool in_h = (h >> 1) < geom.Ho;
        const __nv_bfloat16* drow = dpool + ((long long)bb * geom.Ho + (h >> 1)) * geom.Wo * C;
        for (int e = threadIdx.x; e < n_e; e += blockDim.x) {
            const int w = e >> lg;
            const bool win = in_h && (w >> 1) < geom.Wo;
            const uint4 qd = win ? __ldg(reinterpret_cast<const uint4*>(drow + (long long)(w >> 1) * C) + c8) : make_uint4(0u, 0u, 0u, 0u);
            const uint4 qy = __ldg(yrow + e);
            float g[8], yy[8], rr[8], o[8];
            unpack8(quarter_word(qd), g);
            unpack8(qy, yy);
            if (RES) unpack8(__ldg(rrow + e), rr);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float act = RES ? fmaf(yy[k], mc.scale[k], mc.shift[k]) + rr[k] : fmaf(yy[k], mc.scale[k], mc.shift[k]);
                if (!(__float_as_int(act) > 0x00008000)) g[k] = 0.0f;
                o[k] = fmaf(ca[k], g[k], fmaf(cy[k], yy[k], cc[k]));
            }
            dyrow[e] = pack8(o);
            if (drrow) drrow[e] = pack8(g);
        }
    }
}

// backward of F.avg_pool2d(2x2), floor mode, on NHWC bf16: dx[b][h][w][c] = dy[b][h/2][w/2][c] / 4 (0 in the odd last row / column)
__global__ void avgpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int B, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
    const long long total = (long long)B * H * W * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long p = i / C8;
        const int w = (int)(p % W);
        p /= W;
        const int h = (int)(p % H), b = (int)(p / H);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if ((h >> 1) < Ho && (w >> 1) < Wo) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(dy + ((((long long)b * Ho + (h >> 1)) * Wo + (w >> 1)) * C)) + c8), v);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] *= 0.25f;
        }
        reinterpret_cast<uint4*>(dx)[i] = pack8(v);
    }
}

// sums [C][2] = (dbeta, dgamma) in float64 -> the float32 parameter gradients
__global__ void bn_grads_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)sums[2 * c];
    dgamma[c] = (float)sums[2 * c + 1];
}

// ------------------------------------------------------------------------------------------------
// CompositeCutout (utilities/transforms.py:257-283): RandomCutoutNp (:58-125), SpecAugmentNp (:128-196) or
// RandomCutoutHoleNp (:199-254) all reduce to "up to 8 rectangles (time x frequency) per sample, filled in order with a
// value drawn between the sample's min and max; the last n_zero_channels channels get 0 instead".
// ------------------------------------------------------------------------------------------------
constexpr int kMaxCutRects = 8;

// min / max of every sample (np.min(x), np.max(x) before any cut): one block per sample
__global__ void __launch_bounds__(1024) sample_minmax_kernel(const float* __restrict__ x, long long n_per_sample, float* __restrict__ minmax) {
    const float* p = x + (long long)blockIdx.x * n_per_sample;
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = threadIdx.x; i < n_per_sample; i += blockDim.x) {
        const float v = p[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    __shared__ float s_lo[32], s_hi[32];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, m));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, m));
    }
    if ((threadIdx.x & 31) == 0) {
        s_lo[threadIdx.x >> 5] = lo;
        s_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = threadIdx.x < (blockDim.x >> 5) ? s_lo[threadIdx.x] : INFINITY;
        hi = threadIdx.x < (blockDim.x >> 5) ? s_hi[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, m));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, m));
        }
        if (threadIdx.x == 0) {
            minmax[2 * blockIdx.x] = lo;
            minmax[2 * blockIdx.x + 1] = hi;
        }
    }
}

// one thread per (sample, frame, frequency): the LAST rectangle that covers the position decides its value
__global__ void cutout_kernel(float* __restrict__ x, const int4* __restrict__ rects, const int* __restrict__ n_rects,
                              const double* __restrict__ u, const float* __restrict__ minmax, int B, int C, int T, int F,
                              int n_zero_channels) {
    const long long total = (long long)B * T * F;
    const long long plane = (long long)T * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const long long r = i / F;
        const int t = (int)(r % T), b = (int)(r / T);
        int hit = -1;
        for (int k = n_rects[b] - 1; k >= 0; --k) {
            const int4 q = rects[b * kMaxCutRects + k];           // {top, bottom, left, right}, exclusive ends
            if (t >= q.x && t < q.y && f >= q.z && f < q.w) {
                hit = k;
                break;
            }
        }
        if (hit < 0) continue;
        // np.random.uniform(min, max) = min + (max - min) * u in float64, stored into the float32 array
        const double lo = (double)minmax[2 * b], hi = (double)minmax[2 * b + 1];
        const float fill = (float)(lo + (hi - lo) * u[b * kMaxCutRects + hit]);
        float* dst = x + (long long)b * C * plane + (long long)t * F + f;
        for (int c = 0; c < C; ++c) dst[c * plane] = c < C - n_zero_channels ? fill : 0.0f;
    }
}

// y_doa [B][Ty][3 n] = x | y | z per class: the label side of the two channel swaps
__global__ void augment_doa_kernel(const float* __restrict__ y, float* __restrict__ out, const int4* __restrict__ ops, int B, int Ty, int n) {
    const long long total = (long long)B * Ty * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % n);
        const long long r = i / n;                    // b * Ty + ty
        const int b = (int)(r / Ty);
        const int4 op = ops[b];
        const float* row = y + r * 3 * n;
        float vx = row[k], vy = row[n + k], vz = row[2 * n + k];
        const int m = op.y;
        if (m & 1) { const float t = vx; vx = vy; vy = t; }
        if (op.x == 0) {
            if (m & 2) vx = -vx;
            if (m & 4) vy = -vy;
            if (m & 8) vz = -vz;
        } else {
            if (m & 2) { const float t = -vx; vx = -vy; vy = t; }
            if (m & 4) { vy = -vy; vz = -vz; }
        }
        float* orow = out + r * 3 * n;
        orow[k] = vx;
        orow[n + k] = vy;
        orow[2 * n + k] = vz;
    }
}

// ------------------------------------------------------------------------------------------------
// BaseModel.compute_loss for output_format = 'reg_xyz' (models/interfaces.py:273-355):
//   sed_loss = mean over (row, class) of BCE-with-logits;  doa_loss = sum over x, y, z of sum(|pred - gt| mask) / sum(mask)
//   loss     = w_sed sed_loss + w_doa doa_loss
// One cell = one (row, class): logit[cell], event_gt[cell] (the mask) and the three regressions at [row][c + k n].
// seld_loss_sum_kernel accumulates {sum bce, sum |dx| m, sum |dy| m, sum |dz| m, sum m} in float64 (block tree, then one
// atomicAdd per block and value); seld_loss_finish_kernel turns the sums into (loss, sed_loss, doa_loss) and, when asked,
// d loss / d logit = w_sed (sigmoid(logit) - gt) / cells and d loss / d doa = w_doa sign(pred - gt) mask / sum(mask).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seld_loss_sum_kernel(const float* __restrict__ logit, const float* __restrict__ doa,
                                                            const float* __restrict__ event_gt, const float* __restrict__ doa_gt,
                                                            long long rows, int n, double* __restrict__ sums) {
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const long long cells = rows * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n;
        const int c = (int)(i - r * n);
        const float l = logit[i], y = event_gt[i];
        // torch's stable form: max(l, 0) - l y + log1p(exp(-|l|))
        acc[0] += (double)(fmaxf(l, 0.0f) - l * y + log1pf(expf(-fabsf(l))));
        const float* p = doa + r * 3 * n + c;
        const float* g = doa_gt + r * 3 * n + c;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[1 + k] += (double)(fabsf(p[k * n] - g[k * n]) * y);
        acc[4] += (double)y;
    }
    __shared__ double red[5][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double v = acc[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        atomicAdd(sums + threadIdx.x, v);
    }
}

__global__ void seld_loss_finish_kernel(const float* __restrict__ logit, const float* __restrict__ doa, const float* __restrict__ event_gt,
                                        const float* __restrict__ doa_gt, long long rows, int n, float w_sed, float w_doa,
                                        const double* __restrict__ sums, float* __restrict__ loss, float* __restrict__ g_logit,
                                        float* __restrict__ g_doa) {
    const long long cells = rows * n;
    const double norm = sums[4];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double sed = sums[0] / (double)cells;
        const double dl = (sums[1] + sums[2] + sums[3]) / norm;      // 0 / 0 = NaN without any active cell, as the reference
        loss[0] = (float)((double)w_sed * sed + (double)w_doa * dl);
        loss[1] = (float)sed;
        loss[2] = (float)dl;
    }
    if (!g_logit && !g_doa) return;
    const float inv_cells = (float)(1.0 / (double)cells), inv_norm = (float)(1.0 / norm);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n;
        const int c = (int)(i - r * n);
        const float l = logit[i], y = event_gt[i];
        if (g_logit) g_logit[i] = w_sed * (1.0f / (1.0f + expf(-l)) - y) * inv_cells;
        if (g_doa) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const long long j = r * 3 * n + c + k * n;
                const float d = doa[j] - doa_gt[j];
                g_doa[j] = w_doa * (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * y * inv_norm;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One Adam step on a flat float32 parameter buffer: torch.optim.Adam as the reference configures it
// (models/interfaces.py:85-95: default betas / eps, no weight decay, no amsgrad) with the learning rate and beta1 that
// LearningRateScheduler sets before every batch (utilities/learning_utils.py:39-52).  The float32 operations are
// torch's single-tensor path in its order (lerp, addcmul, sqrt / sqrt(bias_correction2) + eps, addcdiv).
// ------------------------------------------------------------------------------------------------
__global__ void adam_step_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
                                 float* __restrict__ exp_avg_sq, long long n, float w, float beta2, float one_minus_beta2, float eps,
                                 float step_size, float bias2_sqrt) {
    // w = 1 - beta1 and one_minus_beta2 are formed in double by the host, as torch forms them from its Python floats
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float g = grad[i];
        float m = exp_avg[i];
        const float diff = g - m;
        m = w < 0.5f ? fmaf(w, diff, m) : g - diff * (1.0f - w);          // Tensor.lerp_
        float v = exp_avg_sq[i] * beta2;
        v = fmaf(one_minus_beta2 * g, g, v);                               // addcmul_(grad, grad, value = 1 - beta2)
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        const float denom = sqrtf(v) / bias2_sqrt + eps;
        param[i] = param[i] - step_size * (m / denom);                     // addcdiv_(exp_avg, denom, value = -step_size)
    }
}

// The same update with the six per-step scalars read from device memory (hyper = {1 - beta1, beta2, 1 - beta2, eps,
// step_size, bias2_sqrt}, filled by crnn_adam_hyper): a CUDA graph of the training step can be replayed with a new
// learning rate / beta1 / step count by rewriting 24 bytes.
__global__ void adam_step_hyper_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
                                       float* __restrict__ exp_avg_sq, long long n, const float* __restrict__ hyper) {
    const float w = hyper[0], beta2 = hyper[1], one_minus_beta2 = hyper[2], eps = hyper[3], step_size = hyper[4], bias2_sqrt = hyper[5];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float g = grad[i];
        float m = exp_avg[i];
        const float diff = g - m;
        m = w < 0.5f ? fmaf(w, diff, m) : g - diff * (1.0f - w);
        float v = exp_avg_sq[i] * beta2;
        v = fmaf(one_minus_beta2 * g, g, v);
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        const float denom = sqrtf(v) / bias2_sqrt + eps;
        param[i] = param[i] - step_size * (m / denom);
    }
}

}  // namespace crnn
}  // namespace salsa
