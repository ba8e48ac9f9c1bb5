// Model-level C entry points of the SELD CRNN (include/salsa_crnn.h): crnn_load_weights / crnn_workspace_bytes /
// crnn_forward.  What models/seld_models.py:39-49 + experiments/inference.py:102-116 need from a host that is not
// Python: a reference state dict goes in once, then one call per batch runs PannResNet22 (models/encoders.py:48-56),
// SeldDecoder (models/decoders.py:106-154, bigru / avg) and the output stage on caller-owned buffers.
//
// The layer schedule is a fixed sequence of the operator entry points of crnn_abi.cu on slices of ONE caller-provided
// workspace (planned below), captured into a CUDA graph per (shape, buffers) so that a forward is a single graph launch:
// no per-layer tensor-map encoding, no allocation, no host work between the ~30 kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "salsa_crnn.h"

namespace salsa {
namespace crnn {
namespace {

constexpr float kBnEps = 1e-5f;

// float -> bf16 bits, round to nearest even (what torch's .to(torch.bfloat16) does)
uint16_t bf16_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);      // NaN stays NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float bf16_value(uint16_t b) {
    const uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// rows x cols float matrix -> rows x (planes * cols) bf16: hi | mid | lo planes side by side along the last axis
// (salsa_b200/crnn_ops.py: split_planes; each plane is the bf16 rounding of what the previous planes left)
std::vector<uint16_t> split_planes(const std::vector<float>& m, size_t rows, size_t cols, int planes) {
    std::vector<uint16_t> out(rows * cols * planes);
    for (size_t r = 0; r < rows; ++r)
        for (size_t c = 0; c < cols; ++c) {
            float rest = m[r * cols + c];
            for (int p = 0; p < planes; ++p) {
                const uint16_t b = bf16_bits(rest);
                out[(r * planes + p) * cols + c] = b;
                rest = rest - bf16_value(b);
            }
        }
    return out;
}

struct DeviceWeight {
    void* w = nullptr;       // bf16
    float* bias = nullptr;   // fp32
};

struct Model {
    int planes = 1, n_classes = 12, device = 0;
    DeviceWeight cb1, cb2, conv[4][2][2], ds[4];
    void* w_ih[2] = {};
    float *b_ih[2] = {}, *w_hh[2] = {}, *b_hh[2] = {};
    DeviceWeight fc1, fc2;
    std::vector<void*> allocations;
    // graphs of the forward, keyed by everything a launch bakes in
    struct Key {
        const void *feat, *mean, *std, *logits, *doa, *work;
        int B, T_in, T_use, F, n_scaled;
        bool operator<(const Key& o) const { return memcmp(this, &o, sizeof(Key)) < 0; }
    };
    struct Captured {
        cudaGraphExec_t exec;
        int n_kernels;
    };
    std::map<Key, Captured> graphs;
    std::map<Key, int> seen;            // buffers seen once are served directly; the second sighting captures the graph
    cudaStream_t capture_stream = nullptr;      // the schedule is recorded on a stream of our own (the caller's may be the legacy
                                                // default stream, which cannot capture); the graph runs on the caller's
    std::mutex mutex;
};

// SALSA_B200_CRNN_GRAPH=0 (read once) or crnn_model_option("graph", 0): launch the kernels one by one (profilers list them)
int g_use_graph = [] {
    const char* e = getenv("SALSA_B200_CRNN_GRAPH");
    return (e && e[0] == '0') ? 0 : 1;
}();

using Table = std::map<std::string, std::pair<const float*, int64_t>>;

int find(const Table& t, const std::string& name, int64_t numel, const float** out) {
    auto it = t.find(name);
    if (it == t.end()) return fail(SALSA_EINVAL, "crnn_load_weights: missing tensor " + name);
    if (it->second.second != numel)
        return fail(SALSA_EINVAL, "crnn_load_weights: tensor " + name + " has " + std::to_string(it->second.second) + " values, expected " +
                                      std::to_string(numel));
    if (!it->second.first) return fail(SALSA_EINVAL, "crnn_load_weights: null data for " + name);
    *out = it->second.first;
    return SALSA_OK;
}

template <typename T>
int upload(Model* m, const std::vector<T>& h, void** d) {
    SALSA_CUDA(cudaMalloc(d, h.size() * sizeof(T)));
    m->allocations.push_back(*d);
    SALSA_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SALSA_OK;
}

// nn.Conv2d weight (Cout, Cin, k, k) + eval-mode BatchNorm2d -> bf16 [k*k][Cout][planes * Cin_pad] with the BatchNorm
// scale folded in, fp32 bias [Cout] (models/model_utils.py:213-216; salsa_b200/crnn.py: SeldModel._fold)
int fold_conv(Model* m, const Table& t, const std::string& conv_key, const std::string& bn, int cout, int cin, int k, int pad_to,
              DeviceWeight* out) {
    const float *w, *gamma, *beta, *mean, *var;
    int rc;
    if ((rc = find(t, conv_key, (int64_t)cout * cin * k * k, &w)) || (rc = find(t, bn + ".weight", cout, &gamma)) ||
        (rc = find(t, bn + ".bias", cout, &beta)) || (rc = find(t, bn + ".running_mean", cout, &mean)) ||
        (rc = find(t, bn + ".running_var", cout, &var)))
        return rc;
    const int cin_pad = (cin + pad_to - 1) / pad_to * pad_to;
    std::vector<float> wp((size_t)k * k * cout * cin_pad, 0.0f), bias(cout);
    for (int co = 0; co < cout; ++co) {
        const float scale = gamma[co] / sqrtf(var[co] + kBnEps);
        bias[co] = beta[co] - mean[co] * scale;
        for (int ci = 0; ci < cin; ++ci)
            for (int tap = 0; tap < k * k; ++tap)
                wp[((size_t)tap * cout + co) * cin_pad + ci] = w[((size_t)co * cin + ci) * k * k + tap] * scale;
    }
    const std::vector<uint16_t> packed = split_planes(wp, (size_t)k * k * cout, cin_pad, m->planes);
    if ((rc = upload(m, packed, &out->w))) return rc;
    return upload(m, bias, (void**)&out->bias);
}

void release(Model* m) {
    for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second.exec);
    for (void* p : m->allocations) cudaFree(p);
    if (m->capture_stream) cudaStreamDestroy(m->capture_stream);
    delete m;
}

// ---------------------------------------------------------------------------------------------
// workspace plan: every intermediate tensor of the schedule with its lifetime, packed greedily
// ---------------------------------------------------------------------------------------------
struct Slot {
    size_t bytes;
    int first, last;     // steps of the schedule (inclusive) during which the tensor is alive
    size_t offset;
};

struct Plan {
    std::vector<Slot> slots;
    size_t total = 0;
    int add(size_t bytes, int step) {             // alive from the step that writes it ...
        slots.push_back({(bytes + 255) / 256 * 256, step, step, 0});
        return (int)slots.size() - 1;
    }
    void touch(int slot, int step) { slots[slot].last = std::max(slots[slot].last, step); }     // ... to the last step that uses it
    void pack() {
        std::vector<int> order(slots.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return slots[a].bytes > slots[b].bytes; });
        std::vector<int> placed;
        for (int i : order) {
            // lowest offset at which slot i overlaps no already placed slot that is alive at the same time
            std::vector<std::pair<size_t, size_t>> busy;
            for (int j : placed)
                if (!(slots[j].last < slots[i].first || slots[i].last < slots[j].first)) busy.push_back({slots[j].offset, slots[j].offset + slots[j].bytes});
            std::sort(busy.begin(), busy.end());
            size_t at = 0;
            for (auto& b : busy)
                if (at + slots[i].bytes <= b.first) break;
                else at = std::max(at, b.second);
            slots[i].offset = at;
            total = std::max(total, at + slots[i].bytes);
            placed.push_back(i);
        }
    }
};

size_t pad8(size_t n) { return (n + 7) / 8 * 8; }

}  // namespace
}  // namespace crnn
}  // namespace salsa

using namespace salsa;
using namespace salsa::crnn;

namespace {

// The layer schedule, run in two modes with the same control flow: declaring (`packed` null: every intermediate tensor and
// the steps that touch it are recorded in `declare`) and launching (`packed` = the packed plan, `work` = workspace base).
int run_schedule(const Model* m, const Plan* packed, Plan* declare, const float* feat, const float* mean, const float* std_, int n_scaled,
                 int B, int T_in, int T, int F, float* logits, float* doa, char* work, cudaStream_t st) {
    const int P = m->planes;
    const bool run = packed != nullptr;
    int step = 0, n_slots = 0, rc = SALSA_OK;
    auto new_tensor = [&](size_t bytes) -> int { return run ? n_slots++ : declare->add(bytes, step); };
    auto ptr = [&](int slot) -> void* { return run ? (void*)(work + packed->slots[slot].offset) : nullptr; };
    auto act_bytes = [&](int H, int W, int C) -> size_t { return (size_t)B * H * W * C * P * 2; };
    auto zero = [&](int slot, size_t bytes) -> int {
        return run ? check_cuda(cudaMemsetAsync(ptr(slot), 0, bytes, st), "cudaMemsetAsync") : SALSA_OK;
    };
    void* stream = (void*)st;
    // one operator call = one step; `uses` are the workspace tensors it reads or writes
#define STEP(uses, call)                              \
    do {                                              \
        if (!run)                                     \
            for (int s_ : uses)                       \
                if (s_ >= 0) declare->touch(s_, step); \
        if (run && (rc = (call))) return rc;          \
        ++step;                                       \
    } while (0)
    using L = std::initializer_list<int>;

    // ---- encoder (models/encoders.py:48-56; salsa_b200/crnn.py: SeldModel.encode)
    const int x0 = new_tensor(act_bytes(T, F, 16));
    STEP((L{x0}), crnn_pack_input(feat, ptr(x0), B, 7, T_in, F, T, 16, P, mean, std_, n_scaled, stream));
    const int a1 = new_tensor(act_bytes(T, F, 64));
    STEP((L{x0, a1}), crnn_conv_first(ptr(x0), m->cb1.w, m->cb1.bias, ptr(a1), B, T, F, 1, P, stream));
    int H = T, W = F;
    // 3x3 convolution (+ residual) + ReLU (+ the 2x2 average pooling that follows: fused into the epilogue for planes == 1,
    // a separate kernel otherwise)
    auto conv3 = [&](int in, const DeviceWeight& w, int residual, int Cin, int Cout, bool pool, int* out) -> int {
        const bool fuse = pool && P == 1 && H >= 2 && W >= 2;
        const void* res = residual >= 0 ? ptr(residual) : nullptr;
        if (!pool || fuse) {
            *out = new_tensor(act_bytes(fuse ? H / 2 : H, fuse ? W / 2 : W, Cout));
            STEP((L{in, residual, *out}), crnn_conv2d(ptr(in), w.w, w.bias, res, ptr(*out), nullptr, B, H, W, Cin, Cout, 3, 1, P, fuse ? 1 : 0, stream));
            return SALSA_OK;
        }
        const int full = new_tensor(act_bytes(H, W, Cout));
        STEP((L{in, residual, full}), crnn_conv2d(ptr(in), w.w, w.bias, res, ptr(full), nullptr, B, H, W, Cin, Cout, 3, 1, P, 0, stream));
        *out = new_tensor(act_bytes(H / 2, W / 2, Cout));
        STEP((L{full, *out}), crnn_avgpool2(ptr(full), ptr(*out), B, H, W, Cout, P, stream));
        return SALSA_OK;
    };
    int h = -1;
    if ((rc = conv3(a1, m->cb2, -1, 64, 64, true, &h))) return rc;                      // + F.avg_pool2d (models/model_utils.py:220)
    H /= 2;
    W /= 2;
    int inpl = 64;
    const int planes_of[4] = {64, 128, 256, 512};
    for (int li = 0; li < 4; ++li) {
        const int ch = planes_of[li];
        for (int bi = 0; bi < 2; ++bi) {
            // `h` arrives already pooled: the producer of a stride-2 block's input applies the block's avg_pool2d
            // (models/model_utils.py:349, :476) in its epilogue
            int identity = h;
            const int cin = bi == 0 ? inpl : ch;
            if (li > 0 && bi == 0) {
                identity = new_tensor(act_bytes(H, W, ch));
                STEP((L{h, identity}), crnn_conv2d(ptr(h), m->ds[li].w, m->ds[li].bias, nullptr, ptr(identity), nullptr, B, H, W, cin, ch, 1, 0, P, 0, stream));
            }
            const int out1 = new_tensor(act_bytes(H, W, ch));
            STEP((L{h, out1}), crnn_conv2d(ptr(h), m->conv[li][bi][0].w, m->conv[li][bi][0].bias, nullptr, ptr(out1), nullptr, B, H, W, cin, ch, 3, 1, P, 0, stream));
            const bool feeds_stride2 = bi == 1 && li < 3;
            if ((rc = conv3(out1, m->conv[li][bi][1], identity, ch, ch, feeds_stride2, &h))) return rc;
            if (feeds_stride2) {
                H /= 2;
                W /= 2;
            }
        }
        inpl = ch;
    }
    // ---- decoder (models/decoders.py:106-154; salsa_b200/crnn.py: SeldModel.decode)
    const int Tp = H, rows = B * Tp;
    const size_t rows_pad = pad8(rows);
    const int fm = new_tensor(rows_pad * 512 * P * 2);
    if ((rc = zero(fm, rows_pad * 512 * P * 2))) return rc;
    STEP((L{h, fm}), crnn_freq_mean(ptr(h), ptr(fm), rows, W, 512, P, stream));
    int seq = fm;
    for (int layer = 0; layer < 2; ++layer) {
        const int xproj = new_tensor(rows_pad * 1536 * 4);
        STEP((L{seq, xproj}), crnn_gemm(ptr(seq), m->w_ih[layer], m->b_ih[layer], nullptr, (float*)ptr(xproj), rows, 1536, 512, 0, P, stream));
        const int y = new_tensor(rows_pad * 512 * P * 2);
        if ((rc = zero(y, rows_pad * 512 * P * 2))) return rc;
        STEP((L{xproj, y}), crnn_gru_layer((const float*)ptr(xproj), m->w_hh[layer], m->b_hh[layer], ptr(y), B, Tp, P, stream));
        seq = y;
    }
    const int f1 = new_tensor(rows_pad * 1024 * P * 2);
    if ((rc = zero(f1, rows_pad * 1024 * P * 2))) return rc;
    STEP((L{seq, f1}), crnn_gemm(ptr(seq), m->fc1.w, m->fc1.bias, ptr(f1), nullptr, rows, 1024, 512, 1, P, stream));
    const int z = new_tensor(rows_pad * 64 * 4);
    STEP((L{f1, z}), crnn_gemm(ptr(f1), m->fc2.w, m->fc2.bias, nullptr, (float*)ptr(z), rows, 64, 1024, 0, P, stream));
    STEP((L{z}), crnn_head_finish((const float*)ptr(z), logits, doa, rows, m->n_classes, stream));
#undef STEP
    return rc;
}

int build_plan(const Model* m, int B, int T, int F, Plan* plan) {
    int rc = run_schedule(m, nullptr, plan, nullptr, nullptr, nullptr, 0, B, T, T, F, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    plan->pack();
    return SALSA_OK;
}

int check_shape(int B, int T, int F) {
    if (B <= 0 || T < 16 || F < 16) return fail(SALSA_EINVAL, "crnn: need B >= 1 and at least 16 frames and 16 frequencies (four 2x2 poolings)");
    return SALSA_OK;
}

}  // namespace

extern "C" {

int crnn_load_weights(const crnn_tensor_t* tensors, int32_t n_tensors, int32_t planes, int32_t n_classes, void** model_out) {
    if (!tensors || !model_out || n_tensors <= 0) return fail(SALSA_EINVAL, "crnn_load_weights: null argument");
    if (planes < 1 || planes > 3) return fail(SALSA_EINVAL, "crnn_load_weights: planes must be 1 (bf16), 2 (bf16x2) or 3 (bf16x3)");
    if (n_classes <= 0 || 4 * n_classes > 64) return fail(SALSA_EINVAL, "crnn_load_weights: at most 16 classes");
    Table t;
    for (int i = 0; i < n_tensors; ++i)
        if (tensors[i].name) t[tensors[i].name] = {tensors[i].data, tensors[i].numel};
    Model* m = new Model;
    m->planes = planes;
    m->n_classes = n_classes;
    cudaGetDevice(&m->device);
    int rc = SALSA_OK;
    auto conv = [&](const std::string& key, const std::string& bn, int cout, int cin, int k, int pad, DeviceWeight* out) {
        if (!rc) rc = fold_conv(m, t, key, bn, cout, cin, k, pad, out);
    };
    conv("encoder.conv_block1.conv1.weight", "encoder.conv_block1.bn1", 64, 7, 3, 16, &m->cb1);
    conv("encoder.conv_block1.conv2.weight", "encoder.conv_block1.bn2", 64, 64, 3, 64, &m->cb2);
    int inpl = 64;
    const int planes_of[4] = {64, 128, 256, 512};
    for (int li = 0; li < 4; ++li) {
        for (int bi = 0; bi < 2; ++bi) {
            const std::string p = "encoder.resnet.layer" + std::to_string(li + 1) + "." + std::to_string(bi);
            conv(p + ".conv1.weight", p + ".bn1", planes_of[li], bi == 0 ? inpl : planes_of[li], 3, 64, &m->conv[li][bi][0]);
            conv(p + ".conv2.weight", p + ".bn2", planes_of[li], planes_of[li], 3, 64, &m->conv[li][bi][1]);
            if (li > 0 && bi == 0) conv(p + ".downsample.1.weight", p + ".downsample.2", planes_of[li], inpl, 1, 64, &m->ds[li]);
        }
        inpl = planes_of[li];
    }
    // bidirectional GRU, 2 layers: forward | reverse stacked (models/decoders.py:44-46)
    for (int layer = 0; layer < 2 && !rc; ++layer) {
        const std::string l = std::to_string(layer);
        const float *wf, *wr, *bf, *br, *hf, *hr, *bhf, *bhr;
        if ((rc = find(t, "decoder.gru.weight_ih_l" + l, 768 * 512, &wf)) || (rc = find(t, "decoder.gru.weight_ih_l" + l + "_reverse", 768 * 512, &wr)) ||
            (rc = find(t, "decoder.gru.bias_ih_l" + l, 768, &bf)) || (rc = find(t, "decoder.gru.bias_ih_l" + l + "_reverse", 768, &br)) ||
            (rc = find(t, "decoder.gru.weight_hh_l" + l, 768 * 256, &hf)) || (rc = find(t, "decoder.gru.weight_hh_l" + l + "_reverse", 768 * 256, &hr)) ||
            (rc = find(t, "decoder.gru.bias_hh_l" + l, 768, &bhf)) || (rc = find(t, "decoder.gru.bias_hh_l" + l + "_reverse", 768, &bhr)))
            break;
        std::vector<float> w_ih(2 * 768 * 512), b_ih(2 * 768), w_hh(2 * 768 * 256), b_hh(2 * 768);
        memcpy(w_ih.data(), wf, 768 * 512 * 4);
        memcpy(w_ih.data() + 768 * 512, wr, 768 * 512 * 4);
        memcpy(b_ih.data(), bf, 768 * 4);
        memcpy(b_ih.data() + 768, br, 768 * 4);
        memcpy(w_hh.data(), hf, 768 * 256 * 4);
        memcpy(w_hh.data() + 768 * 256, hr, 768 * 256 * 4);
        memcpy(b_hh.data(), bhf, 768 * 4);
        memcpy(b_hh.data() + 768, bhr, 768 * 4);
        if ((rc = upload(m, split_planes(w_ih, 1536, 512, planes), &m->w_ih[layer])) || (rc = upload(m, b_ih, (void**)&m->b_ih[layer])) ||
            (rc = upload(m, w_hh, (void**)&m->w_hh[layer])) || (rc = upload(m, b_hh, (void**)&m->b_hh[layer])))
            break;
    }
    // heads: the four first layers side by side (512 -> 4 x 256), the four second layers block diagonal (1024 -> 64)
    if (!rc) {
        const char* heads[4] = {"event", "x", "y", "z"};
        std::vector<float> w1(1024 * 512), b1(1024), w2(64 * 1024, 0.0f), b2(64, 0.0f);
        for (int i = 0; i < 4 && !rc; ++i) {
            const std::string h = std::string("decoder.") + heads[i];
            const float *a, *ab, *c, *cb;
            if ((rc = find(t, h + "_fc_1.weight", 256 * 512, &a)) || (rc = find(t, h + "_fc_1.bias", 256, &ab)) ||
                (rc = find(t, h + "_fc_2.weight", (int64_t)n_classes * 256, &c)) || (rc = find(t, h + "_fc_2.bias", n_classes, &cb)))
                break;
            memcpy(w1.data() + (size_t)i * 256 * 512, a, 256 * 512 * 4);
            memcpy(b1.data() + i * 256, ab, 256 * 4);
            for (int r = 0; r < n_classes; ++r) {
                memcpy(w2.data() + (size_t)(i * n_classes + r) * 1024 + i * 256, c + (size_t)r * 256, 256 * 4);
                b2[i * n_classes + r] = cb[r];
            }
        }
        if (!rc) rc = upload(m, split_planes(w1, 1024, 512, planes), &m->fc1.w);
        if (!rc) rc = upload(m, b1, (void**)&m->fc1.bias);
        if (!rc) rc = upload(m, split_planes(w2, 64, 1024, planes), &m->fc2.w);
        if (!rc) rc = upload(m, b2, (void**)&m->fc2.bias);
    }
    if (rc) {
        release(m);
        return rc;
    }
    *model_out = m;
    return SALSA_OK;
}

int crnn_free_model(void* model) {
    if (model) release(reinterpret_cast<Model*>(model));
    return SALSA_OK;
}

size_t crnn_workspace_bytes(const void* model, int32_t B, int32_t T, int32_t F) {
    if (!model || check_shape(B, T, F)) return 0;
    Plan plan;
    if (build_plan(reinterpret_cast<const Model*>(model), B, T, F, &plan)) return 0;
    return plan.total;
}

int crnn_forward(void* model, const float* feat, int32_t B, int32_t T_in, int32_t T_use, int32_t F, const float* mean, const float* std_,
                 int32_t n_scaled, float* logits, float* doa, void* workspace, size_t workspace_bytes, void* stream) {
    if (!model || !feat || !logits || !doa) return fail(SALSA_EINVAL, "crnn_forward: null pointer");
    Model* m = reinterpret_cast<Model*>(model);
    int rc = check_shape(B, T_use, F);
    if (rc) return rc;
    if (T_use > T_in) return fail(SALSA_EINVAL, "crnn_forward: T_use exceeds the frames of the input");
    if (n_scaled < 0 || n_scaled > 7 || (n_scaled > 0 && (!mean || !std_))) return fail(SALSA_EINVAL, "crnn_forward: bad scaler");
    Plan plan;
    if ((rc = build_plan(m, B, T_use, F, &plan))) return rc;
    if (!workspace || workspace_bytes < plan.total) return fail(SALSA_ENOMEM, "crnn_forward: workspace smaller than crnn_workspace_bytes()");
    cudaStream_t st = (cudaStream_t)stream;
    char* work = reinterpret_cast<char*>(workspace);
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &capturing);
    if (!g_use_graph || capturing != cudaStreamCaptureStatusNone)
        return run_schedule(m, &plan, nullptr, feat, mean, std_, n_scaled, B, T_in, T_use, F, logits, doa, work, st);
    Model::Key key;
    memset(&key, 0, sizeof(key));
    key.feat = feat; key.mean = mean; key.std = std_; key.logits = logits; key.doa = doa; key.work = workspace;
    key.B = B; key.T_in = T_in; key.T_use = T_use; key.F = F; key.n_scaled = n_scaled;
    std::lock_guard<std::mutex> lock(m->mutex);
    auto it = m->graphs.find(key);
    if (it == m->graphs.end() && m->seen.find(key) == m->seen.end()) {
        // a set of buffers that may never come back (a fresh input tensor per call): not worth a graph instantiation yet
        if (m->seen.size() >= 64) m->seen.clear();
        m->seen[key] = 1;
        return run_schedule(m, &plan, nullptr, feat, mean, std_, n_scaled, B, T_in, T_use, F, logits, doa, work, st);
    }
    if (it == m->graphs.end()) {
        // first call with these buffers: run the schedule under stream capture (tensor maps are encoded once, here)
        cudaGraph_t graph = nullptr;
        const uint64_t before = salsa_launch_count(0);
        if (!m->capture_stream) SALSA_CUDA(cudaStreamCreateWithFlags(&m->capture_stream, cudaStreamNonBlocking));
        SALSA_CUDA(cudaStreamBeginCapture(m->capture_stream, cudaStreamCaptureModeThreadLocal));
        rc = run_schedule(m, &plan, nullptr, feat, mean, std_, n_scaled, B, T_in, T_use, F, logits, doa, work, m->capture_stream);
        const cudaError_t e = cudaStreamEndCapture(m->capture_stream, &graph);
        const int n_kernels = (int)(salsa_launch_count(0) - before);
        count_launch(-n_kernels);              // nothing ran yet: the replay below counts them
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (e != cudaSuccess) return check_cuda(e, "cudaStreamEndCapture");
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return check_cuda(ei, "cudaGraphInstantiate");
        if (m->graphs.size() >= 16) {          // callers that rotate many buffers: keep the cache bounded
            cudaGraphExecDestroy(m->graphs.begin()->second.exec);
            m->graphs.erase(m->graphs.begin());
        }
        it = m->graphs.emplace(key, Model::Captured{exec, n_kernels}).first;
    }
    SALSA_CUDA(cudaGraphLaunch(it->second.exec, st));
    count_launch(it->second.n_kernels);          // kernels of this library inside the replayed graph
    return SALSA_OK;
}

int crnn_model_option(const char* name, int32_t value) {
    const std::string n = name ? name : "";
    if (n == "graph") g_use_graph = value != 0;
    else return fail(SALSA_EINVAL, "unknown option " + n);
    return SALSA_OK;
}

}  // extern "C"
