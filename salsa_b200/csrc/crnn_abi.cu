// C ABI of the SELD CRNN operators (see include/salsa_crnn.h).
#include <cuda.h>
#include <cuda_runtime.h>

#include <math.h>

#include <algorithm>
#include <mutex>
#include <string>

#include "common.cuh"
#include "crnn_conv.cuh"
#include "crnn_kernels.cuh"
#include "crnn_wgrad.cuh"
#include "crnn_gru_train.cuh"
#include "salsa_crnn.h"

namespace salsa {
namespace crnn {

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static int make_tmap(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(SALSA_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SALSA_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return SALSA_OK;
}

// tuning knobs (crnn_set_option): -1 = automatic
static int g_opt_resident = -1;
static int g_opt_tma_store = -1;
static int g_opt_gru_mma = -1;

// output tensor map of the TMA-store epilogue: bf16 NHWC [B][H][W][Cout], box = one 16 x 8 tile x 64 channels
static int make_out_tmap(CUtensorMap* m, const void* out, int B, int H, int W, int Cout, int pool) {
    if (pool) {       // the pooled tensor [B][H/2][W/2][Cout], box = one pooled tile of 8 x 4 pixels
        H /= 2;
        W /= 2;
    }
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(pool ? kTileW / 2 : kTileW), (cuuint32_t)(pool ? kTileH / 2 : kTileH), 1};
    return make_tmap(m, out, 4, dims, str, box);
}

template <int N_TILE>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const ConvArgs& a, cudaStream_t st) {
    const size_t smem = ConvSmem<N_TILE>::total(a.taps, a.resident_b);
    SALSA_CUDA(cudaFuncSetAttribute(conv_tc_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::min(a.n_tiles, sms);
    conv_tc_kernel<N_TILE><<<grid, kConvThreads, smem, st>>>(ta, tw, to, a);
    count_launch();
    return check_cuda(cudaGetLastError(), "conv_tc_kernel");
}

// x: NHWC bf16 [B][H][W][Cin]; for a GEMM: B = 1, W = 8, H = ceil(M / 8), pix_limit = M
static int conv_generic(const void* x, const void* w, const float* bias, const void* residual, void* out, float* out_f32, int B,
                        int H, int W, int Cin, int Cout, int taps, int relu, int planes, int pool, long long pix_limit,
                        cudaStream_t st) {
    if (planes < 1 || planes > 3) return fail(SALSA_EINVAL, "conv: planes must be 1 (bf16), 2 (bf16x2) or 3 (bf16x3)");
    if (!x || !w || (!out && !out_f32)) return fail(SALSA_EINVAL, "conv: null pointer");
    if (B <= 0 || H <= 0 || W <= 0) return fail(SALSA_EINVAL, "conv: bad dimensions");
    if (Cin % kKC != 0 || Cin <= 0) return fail(SALSA_EINVAL, "conv: Cin must be a multiple of 64");
    if (Cout % 64 != 0 || Cout <= 0) return fail(SALSA_EINVAL, "conv: Cout must be a multiple of 64");
    if (taps != 1 && taps != 9) return fail(SALSA_EINVAL, "conv: kernel size must be 1 or 3");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return fail(SALSA_EINVAL, "conv: unaligned pointer");
    // bf16x3 needs two accumulators per stage: 4 * n_tile TMEM columns <= 512
    const int n_tile = (Cout % 256 == 0 && planes == 1) ? 256 : (Cout % 128 == 0 ? 128 : 64);
    const int resident = (planes == 1 && Cin == 64 && Cout == 64 && taps == 9) && (g_opt_resident < 0 ? 1 : g_opt_resident);
    CUtensorMap ta, tw;
    {
        const cuuint64_t cp = (cuuint64_t)Cin * planes;      // channels per pixel in memory
        cuuint64_t dims[4] = {cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {cp * 2, (cuuint64_t)W * cp * 2, (cuuint64_t)H * W * cp * 2};
        cuuint32_t box[4] = {(cuuint32_t)kKC, (cuuint32_t)((taps == 9 && CONV_SINGLE_HALO) ? kHalo1Cols : kTileW),
                             (cuuint32_t)(taps == 9 ? kTileH + 2 : kTileH), 1};
        int rc = make_tmap(&ta, x, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cin * planes, (cuuint64_t)taps * Cout};
        cuuint64_t str[1] = {(cuuint64_t)Cin * planes * 2};
        cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)n_tile};
        int rc = make_tmap(&tw, w, 2, dims, str, box);
        if (rc) return rc;
    }
    ConvArgs a;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    a.taps = taps;
    a.planes = planes;
    a.tiles_w = (W + kTileW - 1) / kTileW;
    a.tiles_h = (H + kTileH - 1) / kTileH;
    a.n_tiles = B * a.tiles_h * a.tiles_w * (Cout / n_tile);
    a.relu = relu;
    a.resident_b = resident;
    a.tma_store = (planes == 1 && out && !out_f32) && (g_opt_tma_store < 0 ? 1 : g_opt_tma_store);
    a.pool = pool;
    if (pool && (!a.tma_store || H < 2 || W < 2)) return fail(SALSA_EINVAL, "conv: fused pooling needs the bf16 single-plane output path");
    a.pix_limit = pix_limit;
    a.bias = bias;
    a.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    a.out = reinterpret_cast<__nv_bfloat16*>(out);
    a.out_f32 = out_f32;
    CUtensorMap to = ta;
    if (a.tma_store) {
        int rc = make_out_tmap(&to, out, B, H, W, Cout, pool);
        if (rc) return rc;
    }
    if (n_tile == 256) return launch_conv<256>(ta, tw, to, a, st);
    if (n_tile == 128) return launch_conv<128>(ta, tw, to, a, st);
    return launch_conv<64>(ta, tw, to, a, st);
}

// first convolution: x NHWC bf16 [B][H][W][planes*16], w [9][64][planes*16]
static int conv_first(const void* x, const void* w, const float* bias, void* out, int B, int H, int W, int relu, int planes,
                      cudaStream_t st) {
    if (!x || !w || !out) return fail(SALSA_EINVAL, "conv_first: null pointer");
    if (planes < 1 || planes > 3) return fail(SALSA_EINVAL, "conv_first: planes must be 1, 2 or 3");
    CUtensorMap ta, tw;
    {
        const cuuint64_t cp = (cuuint64_t)kC1 * planes;
        cuuint64_t dims[4] = {cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {cp * 2, (cuuint64_t)W * cp * 2, (cuuint64_t)H * W * cp * 2};
        cuuint32_t box[4] = {(cuuint32_t)kC1, (cuuint32_t)kTileW, (cuuint32_t)(kTileH + 2), 1};
        int rc = make_tmap(&ta, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)kC1 * planes, (cuuint64_t)9 * 64};
        cuuint64_t str[1] = {(cuuint64_t)kC1 * planes * 2};
        cuuint32_t box[2] = {(cuuint32_t)kC1, 64};
        int rc = make_tmap(&tw, w, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B);
        if (rc) return rc;
    }
    ConvArgs a;
    a.B = B; a.H = H; a.W = W; a.Cin = kC1; a.Cout = 64;
    a.taps = 9;
    a.planes = planes;
    a.tiles_w = (W + kTileW - 1) / kTileW;
    a.tiles_h = (H + kTileH - 1) / kTileH;
    a.n_tiles = B * a.tiles_h * a.tiles_w;
    a.relu = relu;
    a.resident_b = 1;
    a.tma_store = planes == 1 && (g_opt_tma_store < 0 ? 1 : g_opt_tma_store);
    a.pool = 0;
    a.pix_limit = (long long)B * H * W;
    a.bias = bias;
    a.residual = nullptr;
    a.out = reinterpret_cast<__nv_bfloat16*>(out);
    a.out_f32 = nullptr;
    const size_t smem = conv_first_smem(planes);
    SALSA_CUDA(cudaFuncSetAttribute(conv_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    CUtensorMap to = ta;
    if (a.tma_store) {
        int rc = make_out_tmap(&to, out, B, H, W, 64, 0);
        if (rc) return rc;
    }
    conv_first_kernel<<<std::min(a.n_tiles, sms), kConvThreads, smem, st>>>(ta, tw, to, a);
    count_launch();
    return check_cuda(cudaGetLastError(), "conv_first_kernel");
}

static int grid_for(long long total, int block) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (total + block - 1) / block;
    return (int)std::max(1LL, std::min<long long>(want, (long long)sms * 16));
}

}  // namespace crnn
}  // namespace salsa

using namespace salsa;
using namespace salsa::crnn;

extern "C" {

int crnn_conv2d(const void* x, const void* w, const float* bias, const void* residual, void* out, float* out_f32, int32_t B,
                int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize, int32_t relu, int32_t planes, int32_t pool,
                void* stream) {
    if (ksize != 1 && ksize != 3) return fail(SALSA_EINVAL, "conv: kernel size must be 1 or 3");
    return conv_generic(x, w, bias, residual, out, out_f32, B, H, W, Cin, Cout, ksize * ksize, relu, planes, pool,
                        (long long)B * H * W, (cudaStream_t)stream);
}

int crnn_conv_first(const void* x, const void* w, const float* bias, void* out, int32_t B, int32_t H, int32_t W, int32_t relu,
                    int32_t planes, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0) return fail(SALSA_EINVAL, "conv_first: bad dimensions");
    return conv_first(x, w, bias, out, B, H, W, relu, planes, (cudaStream_t)stream);
}

int crnn_gemm(const void* a, const void* w, const float* bias, void* out, float* out_f32, int32_t M, int32_t N, int32_t K,
              int32_t relu, int32_t planes, void* stream) {
    if (M <= 0) return fail(SALSA_EINVAL, "gemm: M must be positive");
    return conv_generic(a, w, bias, nullptr, out, out_f32, 1, (M + 7) / 8, 8, K, N, 1, relu, planes, 0, M, (cudaStream_t)stream);
}

int crnn_set_option(const char* name, int32_t value) {
    const std::string n = name ? name : "";
    if (n == "resident_b") g_opt_resident = value;
    else if (n == "tma_store") g_opt_tma_store = value;
    else if (n == "gru_mma") g_opt_gru_mma = value;
    else return fail(SALSA_EINVAL, "unknown option " + n);
    return SALSA_OK;
}

int crnn_pack_input(const float* x, void* y, int32_t B, int32_t C, int32_t T, int32_t F, int32_t T_use, int32_t Cpad,
                    int32_t planes, const float* mean, const float* std, int32_t n_scaled, void* stream) {
    if (n_scaled < 0 || n_scaled > C || (n_scaled > 0 && (!mean || !std))) return fail(SALSA_EINVAL, "pack_input: bad scaler");
    if (!x || !y) return fail(SALSA_EINVAL, "pack_input: null pointer");
    if (Cpad % 8 != 0 || C > Cpad || T_use > T || B <= 0) return fail(SALSA_EINVAL, "pack_input: bad dimensions");
    const long long n = (long long)B * T_use * F;
    pack_input_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), B, C, T, F, T_use, Cpad, planes, mean, std, n_scaled);
    count_launch();
    return check_cuda(cudaGetLastError(), "pack_input_kernel");
}

int crnn_avgpool2(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, int32_t planes, void* stream) {
    if (!x || !y) return fail(SALSA_EINVAL, "avgpool2: null pointer");
    if (C % 8 != 0 || H < 2 || W < 2) return fail(SALSA_EINVAL, "avgpool2: bad dimensions");
    const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
    avgpool2_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                         reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C, planes);
    count_launch();
    return check_cuda(cudaGetLastError(), "avgpool2_kernel");
}

int crnn_avgpool2_backward(const void* dy, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
    if (!dy || !dx) return fail(SALSA_EINVAL, "avgpool2_backward: null pointer");
    if (C % 8 != 0 || H < 2 || W < 2 || B <= 0) return fail(SALSA_EINVAL, "avgpool2_backward: bad dimensions");
    const long long n = (long long)B * H * W * (C / 8);
    avgpool2_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy),
                                                                             reinterpret_cast<__nv_bfloat16*>(dx), B, H, W, C);
    count_launch();
    return check_cuda(cudaGetLastError(), "avgpool2_bwd_kernel");
}

int crnn_freq_mean(const void* x, void* y, int32_t BH, int32_t W, int32_t C, int32_t planes, void* stream) {
    if (!x || !y) return fail(SALSA_EINVAL, "freq_mean: null pointer");
    if (C % 8 != 0 || W <= 0) return fail(SALSA_EINVAL, "freq_mean: bad dimensions");
    const long long n = (long long)BH * (C / 8);
    freq_mean_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                          reinterpret_cast<__nv_bfloat16*>(y), BH, W, C, planes);
    count_launch();
    return check_cuda(cudaGetLastError(), "freq_mean_kernel");
}

int crnn_gru_layer(const float* xproj, const float* w_hh, const float* b_hh, void* y, int32_t B, int32_t T, int32_t planes,
                   void* stream) {
    if (!xproj || !w_hh || !b_hh || !y) return fail(SALSA_EINVAL, "gru_layer: null pointer");
    if (B <= 0 || T <= 0) return fail(SALSA_EINVAL, "gru_layer: bad dimensions");
    SALSA_CUDA(cudaFuncSetAttribute(gru_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kGruSmemBytes + kGruStageBytes)));
    GruArgs a;
    a.xproj = xproj;
    a.w_hh = w_hh;
    a.b_hh = b_hh;
    a.y = reinterpret_cast<__nv_bfloat16*>(y);
    a.B = B;
    a.T = T;
    a.planes = planes;
    const int groups = (B + kGruClips - 1) / kGruClips;
    if (planes == 1 && (g_opt_gru_mma < 0 ? 1 : g_opt_gru_mma)) {
        SALSA_CUDA(cudaFuncSetAttribute(gru_layer_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGruMmaSmemBytes));
        gru_layer_mma_kernel<<<groups * 2 * kGruCluster, kGruThreads, kGruMmaSmemBytes, (cudaStream_t)stream>>>(a);
        count_launch();
        return check_cuda(cudaGetLastError(), "gru_layer_mma_kernel");
    }
    gru_layer_kernel<<<groups * 2 * kGruCluster, kGruThreads, kGruSmemBytes + kGruStageBytes, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "gru_layer_kernel");
}

int crnn_gru_layer_train(const float* xproj, const float* w_hh, const float* b_hh, float* y, float* save, int32_t B, int32_t T, void* stream) {
    if (!xproj || !w_hh || !b_hh || !y || !save) return fail(SALSA_EINVAL, "gru_layer_train: null pointer");
    if (B <= 0 || T <= 0) return fail(SALSA_EINVAL, "gru_layer_train: bad dimensions");
    constexpr size_t kFwdSmem = kGruSmemBytes + kGruStageBytes;        // + the broadcast staging tile
    SALSA_CUDA(cudaFuncSetAttribute(gru_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
    GruTrainArgs a;
    a.xproj = xproj; a.w_hh = w_hh; a.b_hh = b_hh; a.y = y; a.save = save; a.B = B; a.T = T;
    const int groups = (B + kGruClips - 1) / kGruClips;
    gru_train_fwd_kernel<<<groups * 2 * kGruCluster, kGruThreads, kFwdSmem, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "gru_train_fwd_kernel");
}

int crnn_gru_layer_backward(const float* dy, const float* y, const float* save, const float* w_hh, float* dgi, float* dgh, int32_t B,
                            int32_t T, void* stream) {
    if (!dy || !y || !save || !w_hh || !dgi || !dgh) return fail(SALSA_EINVAL, "gru_layer_backward: null pointer");
    if (B <= 0 || T <= 0) return fail(SALSA_EINVAL, "gru_layer_backward: bad dimensions");
    SALSA_CUDA(cudaFuncSetAttribute(gru_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGruBwdSmemBytes));
    GruBwdArgs a;
    a.dy = dy; a.y = y; a.save = save; a.w_hh = w_hh; a.dgi = dgi; a.dgh = dgh; a.B = B; a.T = T;
    const int groups = (B + kGruClips - 1) / kGruClips;
    gru_train_bwd_kernel<<<groups * 2 * kGruCluster, kGruThreads, kGruBwdSmemBytes, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "gru_train_bwd_kernel");
}

int crnn_head_finish(const float* z, float* logits, float* doa, int32_t rows, int32_t n_classes, void* stream) {
    if (!z || !logits || !doa) return fail(SALSA_EINVAL, "head_finish: null pointer");
    if (4 * n_classes > 64 || rows <= 0) return fail(SALSA_EINVAL, "head_finish: bad dimensions");
    head_finish_kernel<<<grid_for((long long)rows * 4 * n_classes, 256), 256, 0, (cudaStream_t)stream>>>(z, logits, doa, rows, n_classes);
    count_launch();
    return check_cuda(cudaGetLastError(), "head_finish_kernel");
}

int crnn_decode_events(const float* logits, const float* doa, int32_t rows, int32_t n_classes, float threshold, uint8_t* active,
                       int16_t* azi, int16_t* ele, void* stream) {
    if (!logits || !doa || !active || !azi || !ele) return fail(SALSA_EINVAL, "decode_events: null pointer");
    if (rows <= 0 || n_classes <= 0) return SALSA_OK;
    decode_events_kernel<<<grid_for((long long)rows * n_classes, 256), 256, 0, (cudaStream_t)stream>>>(logits, doa, rows, n_classes,
                                                                                                      threshold, active, azi, ele);
    count_launch();
    return check_cuda(cudaGetLastError(), "decode_events_kernel");
}

int crnn_gather_time(const float* in, const int32_t* idx, float* out, int32_t B, int32_t n_in, int32_t n_out, int32_t width,
                     void* stream) {
    if (!in || !idx || !out) return fail(SALSA_EINVAL, "gather_time: null pointer");
    if (B <= 0 || n_out <= 0) return SALSA_OK;
    gather_time_kernel<<<grid_for((long long)B * n_out * width, 256), 256, 0, (cudaStream_t)stream>>>(in, idx, out, B, n_in, n_out, width);
    count_launch();
    return check_cuda(cudaGetLastError(), "gather_time_kernel");
}

int crnn_augment(const float* x, float* out, const float* y_doa, float* y_out, const int32_t* ops, int32_t B, int32_t T, int32_t F,
                 int32_t Ty, int32_t n_classes, void* stream) {
    if (!x || !out || !ops) return fail(SALSA_EINVAL, "augment: null pointer");
    if (x == out || (y_doa && y_doa == y_out)) return fail(SALSA_EINVAL, "augment: in-place operation is not supported");
    if ((y_doa == nullptr) != (y_out == nullptr)) return fail(SALSA_EINVAL, "augment: y_doa and y_out go together");
    if (B <= 0 || T <= 0 || F <= 0) return SALSA_OK;
    augment_kernel<<<grid_for((long long)B * T * F, 256), 256, 0, (cudaStream_t)stream>>>(x, out, reinterpret_cast<const int4*>(ops), B, T, F);
    count_launch();
    int rc = check_cuda(cudaGetLastError(), "augment_kernel");
    if (rc || !y_doa || Ty <= 0 || n_classes <= 0) return rc;
    augment_doa_kernel<<<grid_for((long long)B * Ty * n_classes, 256), 256, 0, (cudaStream_t)stream>>>(
        y_doa, y_out, reinterpret_cast<const int4*>(ops), B, Ty, n_classes);
    count_launch();
    return check_cuda(cudaGetLastError(), "augment_doa_kernel");
}

int crnn_conv_wgrad(const void* x, const void* gy, float* dw, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize,
                    void* stream) {
    if (!x || !gy || !dw) return fail(SALSA_EINVAL, "conv_wgrad: null pointer");
    if (B <= 0 || H <= 0 || W <= 0) return fail(SALSA_EINVAL, "conv_wgrad: bad dimensions");
    if (ksize != 3 && ksize != 1) return fail(SALSA_EINVAL, "conv_wgrad: ksize must be 3 or 1");
    // Cin = 16: the first convolution's input (7 channels padded to 16, crnn_pack_input); the TMA box still spans 64 channels
    // and arrives with channels 16..63 zero-filled, dw is [ksize^2][Cout][16]
    if ((Cin % 64 != 0 && Cin != 16) || Cout % 64 != 0 || Cin <= 0 || Cout <= 0)
        return fail(SALSA_EINVAL, "conv_wgrad: Cin must be a multiple of 64 (or 16), Cout a multiple of 64");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(gy) & 15)) return fail(SALSA_EINVAL, "conv_wgrad: unaligned pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tx, tg;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)kTileW, (cuuint32_t)kTileH, 1};
        int rc = make_tmap(&tx, x, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)kHalo1Cols, (cuuint32_t)(kTileH + 2), 1};      // the halo is taken on dY: one 18 x 16 box
        int rc = make_tmap(&tg, gy, 4, dims, str, box);
        if (rc) return rc;
    }
    WgradArgs a;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    a.tiles_w = (W + kTileW - 1) / kTileW;
    a.tiles_h = (H + kTileH - 1) / kTileH;
    a.n_ktiles = B * a.tiles_h * a.tiles_w;
    a.dw = dw;
    a.taps = ksize * ksize;
    a.cin_valid = std::min(Cin, 64);
    SALSA_CUDA(cudaMemsetAsync(dw, 0, (size_t)a.taps * Cout * Cin * sizeof(float), st));
    SALSA_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemBytes));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int pairs = (Cout / 64) * ((Cin + 63) / 64);
    // split the pixel tiles over enough CTAs to fill the GPU about twice (one CTA per SM: the three-stage ring takes the
    // whole shared memory); every split adds one pass of 9 x 64 x 64 atomic adds
    const int splits = std::max(1, std::min(a.n_ktiles, (2 * sms + pairs - 1) / pairs));
    conv_wgrad_kernel<<<dim3(splits, pairs), kWgThreads, kWgSmemBytes, st>>>(tx, tg, a);
    count_launch();
    return check_cuda(cudaGetLastError(), "conv_wgrad_kernel");
}

static int make_drop(const uint64_t* seed, uint32_t salt, float p, DropArgs* d) {
    if (seed && !(p >= 0.0f && p < 1.0f)) return fail(SALSA_EINVAL, "dropout probability must be in [0, 1)");
    const bool on = seed && p > 0.0f;
    d->seed = on ? reinterpret_cast<const unsigned long long*>(seed) : nullptr;
    d->salt = salt;
    d->threshold = on ? (unsigned int)lrintf(p * 65536.0f) : 0u;
    d->scale = on ? 1.0f / (1.0f - p) : 1.0f;
    return SALSA_OK;
}

int crnn_bn_train_forward(const void* y, const float* gamma, const float* beta, const void* residual, void* z, float* stat,
                          double* sums, float* running_mean, float* running_var, int64_t n_pix, int32_t C, float eps, float momentum,
                          int32_t relu, const uint64_t* drop_seed, uint32_t drop_salt, float drop_p, void* stream) {
    DropArgs drop;
    if (int rcd = make_drop(drop_seed, drop_salt, drop_p, &drop)) return rcd;
    if (!y || !gamma || !beta || !z || !stat || !sums) return fail(SALSA_EINVAL, "bn_train_forward: null pointer");
    if (C <= 0 || C % 8 != 0 || C > 512 || 256 % (C / 8) != 0) return fail(SALSA_EINVAL, "bn_train_forward: C must be 64, 128, 256 or 512");
    if (n_pix <= 0) return fail(SALSA_EINVAL, "bn_train_forward: empty input");
    if ((running_mean == nullptr) != (running_var == nullptr)) return fail(SALSA_EINVAL, "bn_train_forward: running statistics go together");
    cudaStream_t st = (cudaStream_t)stream;
    SALSA_CUDA(cudaMemsetAsync(sums, 0, (size_t)C * 2 * sizeof(double), st));
    const int lanes = 256 / (C / 8);
    const int blocks = (int)std::min<long long>((n_pix + lanes * 32 - 1) / (lanes * 32), 148LL * 8);
    bn_stats_kernel<<<std::max(blocks, 1), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(y), n_pix, C, sums);
    count_launch();
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, n_pix, C, eps, momentum, stat, running_mean, running_var);
    count_launch();
    bn_apply_kernel<<<grid_for(n_pix * (C / 8), 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(y), stat, gamma, beta,
                                                                     reinterpret_cast<const __nv_bfloat16*>(residual),
                                                                     reinterpret_cast<__nv_bfloat16*>(z), n_pix, C, relu, drop);
    count_launch();
    return check_cuda(cudaGetLastError(), "bn_train_forward");
}

int crnn_bn_train_backward(const void* dz, const void* z, const void* y, const float* stat, const float* gamma, const float* beta,
                           void* dy, void* d_residual, double* sums, float* dgamma, float* dbeta, int64_t n_pix, int32_t C,
                           int32_t relu, const uint64_t* drop_seed, uint32_t drop_salt, float drop_p, void* stream) {
    DropArgs drop;
    if (int rcd = make_drop(drop_seed, drop_salt, drop_p, &drop)) return rcd;
    if (!dz || !y || !stat || !gamma || !dy || !sums || !dgamma || !dbeta) return fail(SALSA_EINVAL, "bn_train_backward: null pointer");
    if (relu < 0 || relu > 2) return fail(SALSA_EINVAL, "bn_train_backward: relu is 0 (none), 1 (mask from z) or 2 (mask recomputed from y)");
    if (relu == 1 && !z) return fail(SALSA_EINVAL, "bn_train_backward: the ReLU mask needs the forward output");
    if (relu == 2 && !beta) return fail(SALSA_EINVAL, "bn_train_backward: recomputing the ReLU mask needs beta");
    if (C <= 0 || C % 8 != 0 || C > 512 || 256 % (C / 8) != 0) return fail(SALSA_EINVAL, "bn_train_backward: C must be 64, 128, 256 or 512");
    if (n_pix <= 0) return fail(SALSA_EINVAL, "bn_train_backward: empty input");
    cudaStream_t st = (cudaStream_t)stream;
    SALSA_CUDA(cudaMemsetAsync(sums, 0, (size_t)C * 2 * sizeof(double), st));
    const int lanes = 256 / (C / 8);
    const int blocks = std::max(1, (int)std::min<long long>((n_pix + lanes * 32 - 1) / (lanes * 32), 148LL * 8));
    const int grid2 = grid_for((n_pix * (C / 8) + 1) / 2, 256);
    const __nv_bfloat16 *pdz = reinterpret_cast<const __nv_bfloat16*>(dz), *pz = reinterpret_cast<const __nv_bfloat16*>(z),
                        *py = reinterpret_cast<const __nv_bfloat16*>(y);
    __nv_bfloat16 *pdy = reinterpret_cast<__nv_bfloat16*>(dy), *pdr = reinterpret_cast<__nv_bfloat16*>(d_residual);
    if (relu == 0) {
        bn_bwd_reduce_kernel<0><<<blocks, 256, 0, st>>>(pdz, pz, py, stat, gamma, beta, n_pix, C, sums, drop);
        bn_bwd_apply_kernel<0><<<grid2, 256, 0, st>>>(pdz, pz, py, stat, gamma, beta, sums, pdy, pdr, n_pix, C, drop);
    } else if (relu == 1) {
        bn_bwd_reduce_kernel<1><<<blocks, 256, 0, st>>>(pdz, pz, py, stat, gamma, beta, n_pix, C, sums, drop);
        bn_bwd_apply_kernel<1><<<grid2, 256, 0, st>>>(pdz, pz, py, stat, gamma, beta, sums, pdy, pdr, n_pix, C, drop);
    } else {
        bn_bwd_reduce_kernel<2><<<blocks, 256, 0, st>>>(pdz, pz, py, stat, gamma, beta, n_pix, C, sums, drop);
        bn_bwd_apply_kernel<2><<<grid2, 256, 0, st>>>(pdz, pz, py, stat, gamma, beta, sums, pdy, pdr, n_pix, C, drop);
    }
    count_launch();
    count_launch();
    bn_grads_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, dgamma, dbeta);
    count_launch();
    return check_cuda(cudaGetLastError(), "bn_train_backward");
}

int crnn_bn_train_forward_pool(const void* y, const float* gamma, const float* beta, const void* residual, void* pooled, float* stat,
                               double* sums, float* running_mean, float* running_var, int32_t B, int32_t H, int32_t W, int32_t C,
                               float eps, float momentum, void* stream) {
    if (!y || !gamma || !beta || !pooled || !stat || !sums) return fail(SALSA_EINVAL, "bn_train_forward_pool: null pointer");
    if (C <= 0 || C % 8 != 0 || C > 512 || 256 % (C / 8) != 0) return fail(SALSA_EINVAL, "bn_train_forward_pool: C must be 64, 128, 256 or 512");
    if (B <= 0 || H < 2 || W < 2 || (long long)B * H * W >= (1LL << 31)) return fail(SALSA_EINVAL, "bn_train_forward_pool: bad dimensions");
    if ((running_mean == nullptr) != (running_var == nullptr)) return fail(SALSA_EINVAL, "bn_train_forward_pool: running statistics go together");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_pix = (long long)B * H * W;
    SALSA_CUDA(cudaMemsetAsync(sums, 0, (size_t)C * 2 * sizeof(double), st));
    const int lanes = 256 / (C / 8);
    const int blocks = (int)std::min<long long>((n_pix + lanes * 32 - 1) / (lanes * 32), 148LL * 8);
    bn_stats_kernel<<<std::max(blocks, 1), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(y), n_pix, C, sums);
    count_launch();
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, n_pix, C, eps, momentum, stat, running_mean, running_var);
    count_launch();
    const PoolGeom g = {H, W, H / 2, W / 2};
    bn_apply_pool_kernel<<<std::min(B * g.Ho, 148 * 8), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(y), stat, gamma, beta, reinterpret_cast<const __nv_bfloat16*>(residual),
        reinterpret_cast<__nv_bfloat16*>(pooled), B, g, C);
    count_launch();
    return check_cuda(cudaGetLastError(), "bn_train_forward_pool");
}

int crnn_bn_train_backward_pool(const void* dpool, const void* y, const void* residual, const float* stat, const float* gamma,
                                const float* beta, void* dy, void* d_residual, double* sums, float* dgamma, float* dbeta, int32_t B,
                                int32_t H, int32_t W, int32_t C, void* stream) {
    if (!dpool || !y || !stat || !gamma || !beta || !dy || !sums || !dgamma || !dbeta) return fail(SALSA_EINVAL, "bn_train_backward_pool: null pointer");
    if (d_residual && !residual) return fail(SALSA_EINVAL, "bn_train_backward_pool: a residual gradient needs the residual");
    if (C <= 0 || C % 8 != 0 || C > 512 || 256 % (C / 8) != 0) return fail(SALSA_EINVAL, "bn_train_backward_pool: C must be 64, 128, 256 or 512");
    if (B <= 0 || H < 2 || W < 2 || (long long)B * H * W >= (1LL << 31)) return fail(SALSA_EINVAL, "bn_train_backward_pool: bad dimensions");
    cudaStream_t st = (cudaStream_t)stream;
    SALSA_CUDA(cudaMemsetAsync(sums, 0, (size_t)C * 2 * sizeof(double), st));
    const int blocks = std::min(B * H, 148 * 8);                     // the kernels walk image rows
    const PoolGeom g = {H, W, H / 2, W / 2};
    const __nv_bfloat16 *pd = reinterpret_cast<const __nv_bfloat16*>(dpool), *py = reinterpret_cast<const __nv_bfloat16*>(y),
                        *pr = reinterpret_cast<const __nv_bfloat16*>(residual);
    __nv_bfloat16 *pdy = reinterpret_cast<__nv_bfloat16*>(dy), *pdr = reinterpret_cast<__nv_bfloat16*>(d_residual);
    if (residual) {
        bn_bwd_reduce_pool_kernel<1><<<blocks, 256, 0, st>>>(pd, py, pr, stat, gamma, beta, B, C, g, sums);
        bn_bwd_apply_pool_kernel<1><<<blocks, 256, 0, st>>>(pd, py, pr, stat, gamma, beta, sums, pdy, pdr, B, C, g);
    } else {
        bn_bwd_reduce_pool_kernel<0><<<blocks, 256, 0, st>>>(pd, py, pr, stat, gamma, beta, B, C, g, sums);
        bn_bwd_apply_pool_kernel<0><<<blocks, 256, 0, st>>>(pd, py, pr, stat, gamma, beta, sums, pdy, pdr, B, C, g);
    }
    count_launch();
    count_launch();
    bn_grads_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, dgamma, dbeta);
    count_launch();
    return check_cuda(cudaGetLastError(), "bn_train_backward_pool");
}

int crnn_cutout(float* x, const int32_t* rects, const int32_t* n_rects, const double* u, float* minmax, int32_t B, int32_t C,
                int32_t T, int32_t F, int32_t n_zero_channels, void* stream) {
    if (!x || !rects || !n_rects || !u || !minmax) return fail(SALSA_EINVAL, "cutout: null pointer");
    if (n_zero_channels < 0 || n_zero_channels > C) return fail(SALSA_EINVAL, "cutout: n_zero_channels out of range");
    if (B <= 0 || C <= 0 || T <= 0 || F <= 0) return SALSA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    sample_minmax_kernel<<<B, 1024, 0, st>>>(x, (long long)C * T * F, minmax);
    count_launch();
    int rc = check_cuda(cudaGetLastError(), "sample_minmax_kernel");
    if (rc) return rc;
    cutout_kernel<<<grid_for((long long)B * T * F, 256), 256, 0, st>>>(x, reinterpret_cast<const int4*>(rects), n_rects, u, minmax, B, C, T, F,
                                                                      n_zero_channels);
    count_launch();
    return check_cuda(cudaGetLastError(), "cutout_kernel");
}

int crnn_seld_loss(const float* logit, const float* doa, const float* event_gt, const float* doa_gt, int64_t rows, int32_t n_classes,
                   float w_sed, float w_doa, double* sums, float* loss, float* g_logit, float* g_doa, void* stream) {
    if (!logit || !doa || !event_gt || !doa_gt || !sums || !loss) return fail(SALSA_EINVAL, "seld_loss: null pointer");
    if (rows <= 0 || n_classes <= 0) return fail(SALSA_EINVAL, "seld_loss: empty input");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(sums, 0, 5 * sizeof(double), st), "cudaMemsetAsync");
    if (rc) return rc;
    const long long cells = (long long)rows * n_classes;
    const int blocks = std::min(grid_for(cells, 256), 1024);
    seld_loss_sum_kernel<<<blocks, 256, 0, st>>>(logit, doa, event_gt, doa_gt, rows, n_classes, sums);
    count_launch();
    if ((rc = check_cuda(cudaGetLastError(), "seld_loss_sum_kernel"))) return rc;
    seld_loss_finish_kernel<<<(g_logit || g_doa) ? blocks : 1, 256, 0, st>>>(logit, doa, event_gt, doa_gt, rows, n_classes, w_sed, w_doa, sums,
                                                                           loss, g_logit, g_doa);
    count_launch();
    return check_cuda(cudaGetLastError(), "seld_loss_finish_kernel");
}

int crnn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1, double beta2,
                   double eps, int32_t step, void* stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq) return fail(SALSA_EINVAL, "adam_step: null pointer");
    if (step < 1) return fail(SALSA_EINVAL, "adam_step: step counts from 1");
    if (n <= 0) return SALSA_OK;
    // bias corrections in double like torch's Python scalars, then rounded once
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float step_size = (float)(lr / bc1), bias2_sqrt = (float)sqrt(bc2);
    adam_step_kernel<<<std::min(grid_for(n, 256), 4096), 256, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, step_size, bias2_sqrt);
    count_launch();
    return check_cuda(cudaGetLastError(), "adam_step_kernel");
}

int crnn_adam_hyper(double lr, double beta1, double beta2, double eps, int32_t step, float* hyper_host) {
    if (!hyper_host) return fail(SALSA_EINVAL, "adam_hyper: null pointer");
    if (step < 1) return fail(SALSA_EINVAL, "adam_hyper: step counts from 1");
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    hyper_host[0] = (float)(1.0 - beta1);
    hyper_host[1] = (float)beta2;
    hyper_host[2] = (float)(1.0 - beta2);
    hyper_host[3] = (float)eps;
    hyper_host[4] = (float)(lr / bc1);
    hyper_host[5] = (float)sqrt(bc2);
    return SALSA_OK;
}

int crnn_adam_step_hyper(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* hyper_dev,
                         void* stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper_dev) return fail(SALSA_EINVAL, "adam_step_hyper: null pointer");
    if (n <= 0) return SALSA_OK;
    adam_step_hyper_kernel<<<std::min(grid_for(n, 256), 4096), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n,
                                                                                             hyper_dev);
    count_launch();
    return check_cuda(cudaGetLastError(), "adam_step_hyper_kernel");
}

}  // extern "C"
