// Recurrent part of one bidirectional GRU layer in the TRAINING step (backward for the GRU): the forward
// recurrence that also keeps what the backward pass needs, and back-propagation through time.  Reference: nn.GRU inside
// SeldDecoder (models/decoders.py:44-46, :125-147), trained through autograd (models/seld_models.py:68-76).
//
// Same decomposition as gru_layer_kernel: a thread-block cluster of 8 CTAs owns one (direction, group of 8 clips), CTA c the
// hidden units 32 c .. 32 c + 31, everything float32.  Per time step
//   forward   gh = W_hh h + b_hh (96 rows of W_hh per CTA in shared memory), gates, h broadcast to the cluster through DSMEM;
//             kept per (clip, step, unit): r, z, n and hn = (W_hn h + b_hn)
//   backward  from dh = dL/dh_t (upstream + carried): the pre-activation gradients dgi = (dr, dz, dn) of the input projections
//             and dgh = (dr, dz, r-gated dn) of the hidden projections, both written out (the weight / bias / input gradients
//             are plain GEMMs and column sums over them, done by the caller), dgh broadcast to the cluster, and the carried
//             gradient dh_{t-1} = z dh + W_hh^T dgh with this CTA's 768 x 32 slice of W_hh^T in shared memory.
// Gate order r, z, n as in nn.GRU; n = tanh(gi_n + r * hn).
#pragma once
#include "crnn_kernels.cuh"

namespace salsa {
namespace crnn {

struct GruTrainArgs {
    const float* xproj;     // [B*T][2*768]  W_ih x + b_ih (direction-major: fwd r,z,n | bwd r,z,n)
    const float* w_hh;      // [2][768][256]
    const float* b_hh;      // [2][768]
    float* y;               // [B*T][512]    hidden states (fwd | bwd)
    float* save;            // [B*T][2][4][256]   r, z, n, hn per direction
    int B, T;
};

__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(kGruThreads, 1) gru_train_fwd_kernel(GruTrainArgs a) {
    extern __shared__ __align__(16) float gsm[];
    float* wT = gsm;                                             // [256 k][96 rows]
    float* hbuf = wT + 3 * kGruUnits * kGruHidden;               // [2][256 k][8 clips]
    float* gates = hbuf + 2 * kGruHidden * kGruClips;            // [96 rows][8 clips]
    float* hstage = gates + 3 * kGruUnits * kGruClips;           // [32 own units][8 clips]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / kGruCluster;
    const int dir = cid & 1, grp = cid >> 1;
    const int b0 = grp * kGruClips;
    const int tid = threadIdx.x;
    constexpr int R = 3 * kGruUnits;

    const float* w = a.w_hh + (size_t)dir * 3 * kGruHidden * kGruHidden;
    for (int i = tid; i < R * kGruHidden; i += kGruThreads) {
        const int lr = i / kGruHidden, k = i - lr * kGruHidden;
        const int g = lr / kGruUnits, jl = lr - g * kGruUnits;
        wT[k * R + lr] = w[(size_t)(g * kGruHidden + rank * kGruUnits + jl) * kGruHidden + k];
    }
    for (int i = tid; i < 2 * kGruHidden * kGruClips; i += kGruThreads) hbuf[i] = 0.0f;
    cluster.sync();

    const int lr = tid % R, half = tid / R;
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
        float* hnext = hbuf + ((s + 1) & 1) * kGruHidden * kGruClips;
        float xr = 0.0f, xz = 0.0f, xn = 0.0f;
        if (live) {
            const float* xp = a.xproj + ((size_t)b * a.T + t) * (2 * 3 * kGruHidden) + dir * 3 * kGruHidden + j;
            xr = __ldg(xp);
            xz = __ldg(xp + kGruHidden);
            xn = __ldg(xp + 2 * kGruHidden);
        }
        if (tid < 2 * R) {
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const float4* h4 = reinterpret_cast<const float4*>(hc) + half;
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
            const float hn = gates[(2 * kGruUnits + jl) * kGruClips + bl] + b_hn;
            const float r = 1.0f / (1.0f + expf(-(xr + ar + b_hr)));
            const float z = 1.0f / (1.0f + expf(-(xz + az + b_hz)));
            const float n = tanhf(xn + r * hn);
            const float h_new = (1.0f - z) * n + z * h_prev;
            h_prev = h_new;
            hstage[jl * kGruClips + bl] = h_new;
            if (live) {
                const size_t row = (size_t)b * a.T + t;
                a.y[row * (2 * kGruHidden) + dir * kGruHidden + j] = h_new;
                float* sv = a.save + (row * 2 + dir) * 4 * kGruHidden + j;
                sv[0] = r;
                sv[kGruHidden] = z;
                sv[2 * kGruHidden] = n;
                sv[3 * kGruHidden] = hn;
            }
        }
        __syncthreads();
        // this CTA's 32 x 8 new hidden values are 1 KB contiguous in every CTA's buffer: 64 threads send them as 16-byte
        // stores, 512 contiguous bytes per warp (scalar stores from the (unit, clip) threads would be 32 sectors per warp)
        if (tid < kGruUnits * kGruClips / 4) {
            const float4 v = reinterpret_cast<const float4*>(hstage)[tid];
#pragma unroll
            for (int c = 0; c < kGruCluster; ++c)
                reinterpret_cast<float4*>(cluster.map_shared_rank(hnext, c))[rank * (kGruUnits * kGruClips / 4) + tid] = v;
        }
        cluster.sync();
    }
}

struct GruBwdArgs {
    const float* dy;        // [B*T][512]   dL/dy (fwd | bwd)
    const float* y;         // [B*T][512]   forward hidden states
    const float* save;      // [B*T][2][4][256]
    const float* w_hh;      // [2][768][256]
    float* dgi;             // [B*T][2*768] dL/d(W_ih x + b_ih)
    float* dgh;             // [B*T][2*768] dL/d(W_hh h + b_hh)
    int B, T;
};

constexpr int kGruRows = 3 * kGruHidden;                                           // 768
constexpr size_t kGruBwdSmemBytes = (size_t)(kGruRows * kGruUnits + 2 * kGruRows * kGruClips + 4 * kGruUnits * kGruClips) * sizeof(float);

__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(kGruThreads, 1) gru_train_bwd_kernel(GruBwdArgs a) {
    extern __shared__ __align__(16) float gsm[];
    float* wS = gsm;                                             // [768 rows][32 own units]: W_hh[row][32 rank + k]
    float* gbuf = wS + kGruRows * kGruUnits;                     // [2][768 rows][8 clips]   dgh of the whole cluster
    float* part = gbuf + 2 * kGruRows * kGruClips;               // [4 row quarters][32 units][8 clips]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / kGruCluster;
    const int dir = cid & 1, grp = cid >> 1;
    const int b0 = grp * kGruClips;
    const int tid = threadIdx.x;

    const float* w = a.w_hh + (size_t)dir * kGruRows * kGruHidden;
    for (int i = tid; i < kGruRows * kGruUnits; i += kGruThreads) {
        const int row = i / kGruUnits, k = i - row * kGruUnits;
        wS[i] = w[(size_t)row * kGruHidden + rank * kGruUnits + k];
    }
    cluster.sync();

    const int jl = tid & 31, bl = tid >> 5;               // gate mapping: (unit, clip)
    const int j = rank * kGruUnits + jl;
    const int b = b0 + bl;
    const bool live = b < a.B;
    // matvec mapping: unit kl = tid & 31, clip half (4 clips) = (tid >> 5) & 1, row quarter = tid >> 6
    const int kl = tid & 31, half = (tid >> 5) & 1, rq = tid >> 6;
    float dh_carry = 0.0f;

    for (int s = 0; s < a.T; ++s) {
        const int t = dir ? s : a.T - 1 - s;                      // reverse of the forward order
        const int tp = dir ? t + 1 : t - 1;                        // time index of h_{prev} in the forward recurrence
        float* gcur = gbuf + (s & 1) * kGruRows * kGruClips;
        float dr_pre = 0.0f, dz_pre = 0.0f, dhn = 0.0f, dh_direct = 0.0f;
        if (live) {
            const size_t row = (size_t)b * a.T + t;
            const float* sv = a.save + (row * 2 + dir) * 4 * kGruHidden + j;
            const float r = __ldg(sv), z = __ldg(sv + kGruHidden), n = __ldg(sv + 2 * kGruHidden), hn = __ldg(sv + 3 * kGruHidden);
            const float h_prev = (tp >= 0 && tp < a.T) ? __ldg(a.y + ((size_t)b * a.T + tp) * (2 * kGruHidden) + dir * kGruHidden + j) : 0.0f;
            const float dh = __ldg(a.dy + row * (2 * kGruHidden) + dir * kGruHidden + j) + dh_carry;
            const float dn_pre = dh * (1.0f - z) * (1.0f - n * n);
            dz_pre = dh * (h_prev - n) * z * (1.0f - z);
            dr_pre = dn_pre * hn * r * (1.0f - r);
            dhn = dn_pre * r;
            dh_direct = dh * z;
            float* gi = a.dgi + row * (2 * kGruRows) + dir * kGruRows + j;
            gi[0] = dr_pre;
            gi[kGruHidden] = dz_pre;
            gi[2 * kGruHidden] = dn_pre;
            float* gh = a.dgh + row * (2 * kGruRows) + dir * kGruRows + j;
            gh[0] = dr_pre;
            gh[kGruHidden] = dz_pre;
            gh[2 * kGruHidden] = dhn;
        }
        // the three hidden-projection gradients of this (unit, clip) go to every CTA of the cluster: staged locally ([3 gates]
        // [32 units][8 clips] = three 1 KB pieces that are contiguous in the destination), then sent as 16-byte stores, 512
        // contiguous bytes per warp
        part[jl * kGruClips + bl] = dr_pre;
        part[(kGruUnits + jl) * kGruClips + bl] = dz_pre;
        part[(2 * kGruUnits + jl) * kGruClips + bl] = dhn;
        __syncthreads();
        if (tid < 3 * kGruUnits * kGruClips / 4) {
            constexpr int kPiece = kGruUnits * kGruClips / 4;     // float4 per gate piece (64)
            const int gate = tid / kPiece, off = tid - gate * kPiece;
            const float4 v = reinterpret_cast<const float4*>(part)[tid];
            const int dst = (gate * kGruHidden + rank * kGruUnits) * (kGruClips / 4) + off;
#pragma unroll
            for (int c = 0; c < kGruCluster; ++c) reinterpret_cast<float4*>(cluster.map_shared_rank(gcur, c))[dst] = v;
        }
        cluster.sync();
        // dh_prev[k][clip] = sum over the 768 rows of W_hh[row][k] * dgh[row][clip]: a quarter of the rows per thread
        {
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const float4* g4 = reinterpret_cast<const float4*>(gcur) + half;
            const int r0 = rq * (kGruRows / 4);
#pragma unroll 8
            for (int rr = 0; rr < kGruRows / 4; ++rr) {
                const float wv = wS[(r0 + rr) * kGruUnits + kl];
                const float4 gv = g4[(r0 + rr) * 2];
                acc[0] = fmaf(wv, gv.x, acc[0]);
                acc[1] = fmaf(wv, gv.y, acc[1]);
                acc[2] = fmaf(wv, gv.z, acc[2]);
                acc[3] = fmaf(wv, gv.w, acc[3]);
            }
            *reinterpret_cast<float4*>(part + (rq * kGruUnits + kl) * kGruClips + half * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncthreads();
        dh_carry = dh_direct;
#pragma unroll
        for (int q = 0; q < 4; ++q) dh_carry += part[(q * kGruUnits + jl) * kGruClips + bl];
        __syncthreads();                                          // `part` is rewritten in the next step
    }
}

}  // namespace crnn
}  // namespace salsa
