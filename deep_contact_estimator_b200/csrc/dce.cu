// libdce_b200.so — C ABI (include/dce.h) over the sm_100a kernels.
//
// Replaces, for the contact-classification hot path only:
//   contact_cnn.forward            /root/reference/src/contact_cnn.py:60-66
//   window extract + z-score       /root/reference/utils/data_handler.py:55-57
//   argmax + decimal2binary        /root/reference/src/inference_one_seq.py:26-27,59-62
// No torch types here: plain device pointers, sizes and a stream.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stddef.h>
#include <new>

#include "../../include/dce.h"
#include "dce_common.cuh"
#include "dce_fp32.cuh"
#include "dce_tc.cuh"
#include "dce_tc_run.cuh"
#include "dce_latency.cuh"

namespace {

thread_local int g_last_cuda_error = 0;
thread_local int g_launches = 0;

inline int cuda_fail(cudaError_t e) { g_last_cuda_error = (int)e; return DCE_ECUDA; }
#define DCE_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cuda_fail(e_); } while (0)
using dce::align_up;
using dce::Ctx;

// ---- packed-weight buffer layout (one allocation so it can be broadcast) ----
struct Fp32Layout {
    size_t w1, w2, w3, w4;          // conv [3][cin][cout]
    size_t f1, f2, f3, f3t;         // fc [K][N], [K][N], [16][512], [512][16]
    size_t b[7];                    // biases in layer order
    size_t w4q, f1s, f2s;           // latency kernel: per-CTA contiguous slices (dce_latency.cuh)
    size_t end;
};

Fp32Layout make_fp32_layout(size_t base) {
    Fp32Layout L; size_t o = base;
    auto take = [&](size_t floats) { size_t r = o; o = align_up(o + floats * 4, 256); return r; };
    L.w1 = take(3 * 54 * 64);  L.w2 = take(3 * 64 * 64);  L.w3 = take(3 * 64 * 128);  L.w4 = take(3 * 128 * 128);
    L.f1 = take((size_t)4736 * 2048);  L.f2 = take((size_t)2048 * 512);  L.f3 = take(16 * 512); L.f3t = take(16 * 512);
    const int bn[7] = {64, 64, 128, 128, 2048, 512, 16};
    for (int i = 0; i < 7; ++i) L.b[i] = take(bn[i]);
    L.w4q = take(3 * 128 * 128);  L.f1s = take((size_t)4736 * 2048);  L.f2s = take((size_t)2048 * 512);
    L.end = o;
    return L;
}

}  // namespace

struct dce_weights {
    int device = -1;
    int sm_count = 0;
    bool packed = false;
    char* buf = nullptr;            // device
    size_t bytes = 0;
    Fp32Layout f32;
    dce::tc::PackedLayout tc;
    dce::tc::Options opt;           // ablation / debugging switches (dce_weights_set_option); read-only for dce_forward / dce_stream
};

namespace {

constexpr int64_t kChunkFp32 = 4096;   // windows per internal pass (bounds the workspace)

struct Fp32Workspace { size_t act4, h1, h2, end; };
Fp32Workspace fp32_workspace(int64_t n) {
    Fp32Workspace W; size_t o = 512;      // [0,256): the latency kernel's barrier counters (dce_latency.cuh); [256,512): fc.3 tickets (dce_tc.cuh)
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    W.act4 = take((size_t)n * 4736 * 4); W.h1 = take((size_t)n * 2048 * 4); W.h2 = take((size_t)n * 512 * 4);
    W.end = o; return W;
}

template <class T> const T* at(const dce_weights* w, size_t off) { return reinterpret_cast<const T*>(w->buf + off); }
template <class T> T* at_mut(dce_weights* w, size_t off) { return reinterpret_cast<T*>(w->buf + off); }

int pack_fp32(dce_weights* w, const float* const p[DCE_NUM_PARAMS], Ctx& ctx) {
    cudaStream_t s = ctx.stream;
    using namespace dce::fp32;
    const Fp32Layout& L = w->f32;
    struct { int idx, cout, cin; size_t off; } convs[4] = {
        {0, 64, 54, L.w1}, {2, 64, 64, L.w2}, {4, 128, 64, L.w3}, {6, 128, 128, L.w4}};
    for (auto& c : convs) {
        int n = 3 * c.cin * c.cout;
        DCE_KL(ctx, "pack_conv", pack_conv_kernel<<<(n + 255) / 256, 256, 0, s>>>(p[c.idx], at_mut<float>(w, c.off), c.cout, c.cin));
    }
    DCE_KL(ctx, "pack_fc1", pack_fc_kernel<<<dim3(4736 / 32, 2048 / 32), dim3(32, 8), 0, s>>>(p[8], at_mut<float>(w, L.f1), 2048, 4736, 1));
    DCE_KL(ctx, "pack_fc2", pack_fc_kernel<<<dim3(2048 / 32, 512 / 32), dim3(32, 8), 0, s>>>(p[10], at_mut<float>(w, L.f2), 512, 2048, 0));
    DCE_CUDA(cudaMemcpyAsync(at_mut<float>(w, L.f3), p[12], 16 * 512 * 4, cudaMemcpyDeviceToDevice, s));
    DCE_KL(ctx, "pack_fc3", pack_fc_kernel<<<dim3(512 / 32, 1), dim3(32, 8), 0, s>>>(p[12], at_mut<float>(w, L.f3t), 16, 512, 0));
    const int bidx[7] = {1, 3, 5, 7, 9, 11, 13};
    const int bn[7] = {64, 64, 128, 128, 2048, 512, 16};
    for (int i = 0; i < 7; ++i)
        DCE_CUDA(cudaMemcpyAsync(at_mut<float>(w, L.b[i]), p[bidx[i]], bn[i] * 4, cudaMemcpyDeviceToDevice, s));
    // latency kernel slices, cut from the images above
    DCE_KL(ctx, "pack_w4q", dce::lat::pack_w4q_kernel<<<(4 * 384 * 32 + 255) / 256, 256, 0, s>>>(at<float>(w, L.w4), at_mut<float>(w, L.w4q)));
    DCE_KL(ctx, "pack_f1s", dce::lat::pack_f1s_kernel<<<(dce::lat::kSlices * 4736 * 4 + 255) / 256, 256, 0, s>>>(at<float>(w, L.f1), at_mut<float>(w, L.f1s)));
    DCE_KL(ctx, "pack_f2s", dce::lat::pack_f2s_kernel<<<(dce::lat::kSlices * 2048 + 255) / 256, 256, 0, s>>>(at<float>(w, L.f2), at_mut<float>(w, L.f2s)));
    return DCE_OK;
}

// fp32 mode: conv stack (one CTA per window) -> fc.0 SGEMM -> fc.3 SGEMM -> fc.6 + argmax + bits
int run_fp32(const dce_weights* w, const float* x, bool normalize, int64_t first, int64_t n,
             float* logits, int32_t* cls, uint8_t* bits, char* ws, Ctx& ctx) {
    cudaStream_t s = ctx.stream;
    using namespace dce::fp32;
    const Fp32Layout& L = w->f32;
    static dce::DeviceOnce attr_once;
    if (auto first_ = attr_once.need()) {
        DCE_CUDA(cudaFuncSetAttribute(conv_stack_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes));
        DCE_CUDA(cudaFuncSetAttribute(conv_stack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes));
        DCE_CUDA(cudaFuncSetAttribute(fc3_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFc3SmemBytes));
    }
    ConvParams cp{at<float>(w, L.w1), at<float>(w, L.b[0]), at<float>(w, L.w2), at<float>(w, L.b[1]),
                  at<float>(w, L.w3), at<float>(w, L.b[2]), at<float>(w, L.w4), at<float>(w, L.b[3])};
    for (int64_t c0 = 0; c0 < n; c0 += kChunkFp32) {
        const int64_t m = (n - c0 < kChunkFp32) ? n - c0 : kChunkFp32;
        Fp32Workspace W = fp32_workspace(m);
        float* act4 = reinterpret_cast<float*>(ws + W.act4);
        float* h1 = reinterpret_cast<float*>(ws + W.h1);
        float* h2 = reinterpret_cast<float*>(ws + W.h2);
        const int grid = (int)((m < (int64_t)w->sm_count * 8) ? m : (int64_t)w->sm_count * 8);
        if (normalize)
            DCE_KL(ctx, "fp32_conv_stack_norm", conv_stack_kernel<true><<<grid, 256, kConvSmemBytes, s>>>(x, first + c0, m, cp, act4));
        else
            DCE_KL(ctx, "fp32_conv_stack", conv_stack_kernel<false><<<grid, 256, kConvSmemBytes, s>>>(x + c0 * (150 * 54), 0, m, cp, act4));
        DCE_KL(ctx, "fp32_fc1_sgemm", sgemm_bias_kernel<true><<<dim3(2048 / 128, (unsigned)((m + 127) / 128)), 256, 0, s>>>(
            act4, at<float>(w, L.f1), at<float>(w, L.b[4]), h1, (int)m, 2048, 4736));
        DCE_KL(ctx, "fp32_fc2_sgemm", sgemm_bias_kernel<true><<<dim3(512 / 128, (unsigned)((m + 127) / 128)), 256, 0, s>>>(
            h1, at<float>(w, L.f2), at<float>(w, L.b[5]), h2, (int)m, 512, 2048));
        const int g3 = (int)((m + 15) / 16 < (int64_t)w->sm_count * 2 ? (m + 15) / 16 : (int64_t)w->sm_count * 2);
        DCE_KL(ctx, "fp32_fc3_argmax", fc3_argmax_kernel<<<g3, 256, kFc3SmemBytes, s>>>(h2, at<float>(w, L.f3t), at<float>(w, L.b[6]), m,
                                            logits ? logits + c0 * 16 : nullptr, cls ? cls + c0 : nullptr,
                                            bits ? bits + c0 * 4 : nullptr));
    }
    return DCE_OK;
}

int check_common(const dce_weights* w, const void* ws, size_t ws_bytes, int64_t n, int precision) {
    if (!w) return DCE_EINVAL;
    if (!w->packed) return DCE_ENOTPACKED;
    if (n < 0) return DCE_EINVAL;
    if (precision != DCE_PREC_FP32 && precision != DCE_PREC_BF16X3) return DCE_EINVAL;
    if (n > 0) {
        if (!ws) return DCE_EINVAL;
        if ((uintptr_t)ws % 256) return DCE_EALIGN;
        if (ws_bytes < dce_workspace_bytes(n, precision)) return DCE_EWORKSPACE;
    }
    return DCE_OK;
}

}  // namespace

extern "C" {

int dce_version(void) { return DCE_VERSION; }

const char* dce_strerror(int code) {
    switch (code) {
        case DCE_OK: return "ok";
        case DCE_EINVAL: return "invalid argument";
        case DCE_EALIGN: return "misaligned pointer";
        case DCE_EARCH: return "device is not sm_100 (B200)";
        case DCE_ECUDA: return "CUDA runtime error (see dce_last_cuda_error)";
        case DCE_ENOTPACKED: return "weights not packed";
        case DCE_EWORKSPACE: return "workspace too small";
        case DCE_EUNSUPPORTED: return "unsupported precision/mode";
        default: return "unknown error";
    }
}

int dce_last_cuda_error(void) { return g_last_cuda_error; }
int dce_last_launch_count(void) { return g_launches; }

int dce_weights_create(dce_weights** out, int device) {
    if (!out) return DCE_EINVAL;
    *out = nullptr;
    cudaDeviceProp prop;
    DCE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10 || prop.minor != 0) return DCE_EARCH;      // the library holds sm_100a code only
    int prev = 0;
    DCE_CUDA(cudaGetDevice(&prev));
    DCE_CUDA(cudaSetDevice(device));
    dce_weights* w = new (std::nothrow) dce_weights();
    if (!w) { cudaSetDevice(prev); return DCE_EINVAL; }
    w->device = device;
    w->sm_count = prop.multiProcessorCount;
    w->f32 = make_fp32_layout(0);
    w->tc = dce::tc::make_packed_layout(w->f32.end);
    w->bytes = w->tc.end;
    cudaError_t e = cudaMalloc(&w->buf, w->bytes);
    if (e == cudaSuccess) e = cudaMemset(w->buf, 0, w->bytes);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { delete w; return cuda_fail(e); }
    *out = w;
    return DCE_OK;
}

int dce_weights_destroy(dce_weights* w) {
    if (!w) return DCE_OK;
    if (w->opt.trace) cudaFree(w->opt.trace);
    if (w->buf) cudaFree(w->buf);
    delete w;
    return DCE_OK;
}

int dce_weights_pack(dce_weights* w, const float* const params_dev[DCE_NUM_PARAMS], void* stream) {
    if (!w || !params_dev) return DCE_EINVAL;
    for (int i = 0; i < DCE_NUM_PARAMS; ++i) {
        if (!params_dev[i]) return DCE_EINVAL;
        if ((uintptr_t)params_dev[i] % 4) return DCE_EALIGN;
    }
    Ctx ctx; ctx.stream = (cudaStream_t)stream;
    int rc = pack_fp32(w, params_dev, ctx);
    if (rc == DCE_OK) rc = dce::tc::pack(w->buf, w->tc, params_dev, ctx);
    g_launches = ctx.launches;
    if (rc == DCE_ECUDA && ctx.err != cudaSuccess) g_last_cuda_error = (int)ctx.err;
    if (rc != DCE_OK) return rc;
    w->packed = true;
    return DCE_OK;
}

size_t dce_weights_packed_bytes(const dce_weights* w) { return w ? w->bytes : 0; }
void* dce_weights_packed_ptr(dce_weights* w) { return w ? (void*)w->buf : nullptr; }
int dce_weights_adopt(dce_weights* w) { if (!w) return DCE_EINVAL; w->packed = true; return DCE_OK; }

size_t dce_workspace_bytes(int64_t max_windows, int precision) {
    if (max_windows <= 0) return 512;
    size_t need = 0;
    if (precision == DCE_PREC_FP32) {
        const int64_t n = max_windows < kChunkFp32 ? max_windows : kChunkFp32;
        need = fp32_workspace(n).end;
    } else if (precision == DCE_PREC_BF16X3) {
        need = dce::tc::workspace_bytes(max_windows);
    } else {
        return 0;
    }
    const size_t lat = dce::lat::workspace_bytes();      // calls of <= 4 windows run the fused latency kernel
    return need > lat ? need : lat;
}

}  // extern "C"

namespace {

// shared body of dce_forward / dce_stream / their profiling variants
int run_any(const dce_weights* w, const float* src, bool is_stream, int64_t T, int64_t first, int64_t n,
            float* logits, int32_t* cls, uint8_t* bits, void* ws, size_t ws_bytes, int precision, Ctx& ctx) {
    int rc = check_common(w, ws, ws_bytes, n, precision);
    if (rc != DCE_OK) return rc;
    if (is_stream) {
        if (first < 0 || T < 0) return DCE_EINVAL;
        if (n > 0 && first + n + (DCE_WINDOW - 1) > T) return DCE_EINVAL;      // utils/data_handler.py:24
    }
    if (n == 0) return DCE_OK;
    if (!src) return DCE_EINVAL;
    if ((uintptr_t)src % 16 || (logits && (uintptr_t)logits % 16) || (bits && (uintptr_t)bits % 4) || (cls && (uintptr_t)cls % 4))
        return DCE_EALIGN;
    rc = DCE_EUNSUPPORTED;
    if (n <= dce::lat::kMaxB && w->opt.latency_kernel) {
        // latency mode (K3): one cooperative fp32 kernel for the whole path, both precision modes
        const Fp32Layout& L = w->f32;
        dce::lat::Weights wt;
        wt.w1 = at<float>(w, L.w1); wt.w2 = at<float>(w, L.w2); wt.w3 = at<float>(w, L.w3); wt.w4q = at<float>(w, L.w4q);
        wt.f1s = at<float>(w, L.f1s); wt.f2s = at<float>(w, L.f2s); wt.f3t = at<float>(w, L.f3t);
        for (int i = 0; i < 7; ++i) wt.b[i] = at<float>(w, L.b[i]);
        rc = dce::lat::run(wt, w->sm_count, src, is_stream, first, (int)n, logits, cls, bits, (char*)ws, ctx,
                           w->opt.latency_coop, w->opt.latency_tma_in);
    }
    if (rc != DCE_EUNSUPPORTED) { /* done (or failed) in the latency kernel */ }
    else if (precision == DCE_PREC_FP32)
        rc = run_fp32(w, src, is_stream, first, n, logits, cls, bits, (char*)ws, ctx);
    else
    {
        const Fp32Layout& L = w->f32;
        dce::tc::BiasPtrs bp;
        for (int i = 0; i < 7; ++i) bp.b[i] = at<float>(w, L.b[i]);
        bp.w3 = at<float>(w, L.f3t); bp.f1 = at<float>(w, L.f1); bp.f2 = at<float>(w, L.f2);
        rc = dce::tc::run(w->buf, w->tc, bp, w->opt, w->sm_count, src, is_stream, T, first, n, logits, cls, bits, (char*)ws, ctx);
    }
    if (rc == DCE_ECUDA && ctx.err != cudaSuccess) g_last_cuda_error = (int)ctx.err;
    return rc;
}

}  // namespace

extern "C" {

int dce_forward(const dce_weights* w, const float* x_dev, int64_t B,
                float* logits_dev, int32_t* cls_dev, uint8_t* bits_dev,
                void* workspace_dev, size_t workspace_bytes, int precision, void* stream) {
    Ctx ctx; ctx.stream = (cudaStream_t)stream;
    int rc = run_any(w, x_dev, false, 0, 0, B, logits_dev, cls_dev, bits_dev, workspace_dev, workspace_bytes, precision, ctx);
    g_launches = ctx.launches;
    return rc;
}

int dce_stream(const dce_weights* w, const float* data_dev, int64_t T,
               int64_t first_window, int64_t n_windows,
               float* logits_dev, int32_t* cls_dev, uint8_t* bits_dev,
               void* workspace_dev, size_t workspace_bytes, int precision, void* stream) {
    Ctx ctx; ctx.stream = (cudaStream_t)stream;
    int rc = run_any(w, data_dev, true, T, first_window, n_windows, logits_dev, cls_dev, bits_dev,
                     workspace_dev, workspace_bytes, precision, ctx);
    g_launches = ctx.launches;
    return rc;
}

int dce_latency_server_start(const dce_weights* w, const float* x_host, int n, float* logits_host, int32_t* cls_host,
                             uint8_t* bits_host, dce_latency_ctrl* ctrl, void* workspace_dev, size_t workspace_bytes,
                             double idle_timeout_s, void* stream) {
    int rc = check_common(w, workspace_dev, workspace_bytes, n, DCE_PREC_BF16X3);
    if (rc != DCE_OK) return rc;
    if (n < 1 || n > dce::lat::kMaxB || !x_host || !ctrl || !(idle_timeout_s > 0.0)) return DCE_EINVAL;
    if ((uintptr_t)x_host % 16 || (uintptr_t)ctrl % 128 || (logits_host && (uintptr_t)logits_host % 16) ||
        (bits_host && (uintptr_t)bits_host % 4) || (cls_host && (uintptr_t)cls_host % 4)) return DCE_EALIGN;
    ctrl->seq_in = 0; ctrl->seq_out = 0; ctrl->quit = 0; ctrl->alive = 0;
    const Fp32Layout& L = w->f32;
    dce::lat::Weights wt;
    wt.w1 = at<float>(w, L.w1); wt.w2 = at<float>(w, L.w2); wt.w3 = at<float>(w, L.w3); wt.w4q = at<float>(w, L.w4q);
    wt.f1s = at<float>(w, L.f1s); wt.f2s = at<float>(w, L.f2s); wt.f3t = at<float>(w, L.f3t);
    for (int i = 0; i < 7; ++i) wt.b[i] = at<float>(w, L.b[i]);
    Ctx ctx; ctx.stream = (cudaStream_t)stream;
    rc = dce::lat::run(wt, w->sm_count, x_host, false, 0, n, logits_host, cls_host, bits_host, (char*)workspace_dev, ctx, 1, 1,
                       reinterpret_cast<volatile unsigned*>(ctrl), (unsigned long long)(idle_timeout_s * 1e9));
    g_launches = ctx.launches;
    if (rc == DCE_ECUDA && ctx.err != cudaSuccess) g_last_cuda_error = (int)ctx.err;
    return rc;
}

int dce_latency_row_server_start(const dce_weights* w, dce_latency_row_ctrl* ctrl, float* ring_dev, void* workspace_dev,
                                 size_t workspace_bytes, double idle_timeout_s, void* stream) {
    int rc = check_common(w, workspace_dev, workspace_bytes, 1, DCE_PREC_BF16X3);
    if (rc != DCE_OK) return rc;
    if (!ring_dev || !ctrl || !(idle_timeout_s > 0.0)) return DCE_EINVAL;
    if ((uintptr_t)ring_dev % 16 || (uintptr_t)ctrl % 128) return DCE_EALIGN;
    static_assert(sizeof(dce_latency_row_ctrl) == 512 && offsetof(dce_latency_row_ctrl, seq_out) == dce::lat::kRowSeqOut * 4 &&
                  offsetof(dce_latency_row_ctrl, alive) == dce::lat::kRowAlive * 4, "include/dce.h and dce_latency.cuh agree on the layout");
    static_assert(sizeof(dce_latency_ctrl) == 128 && offsetof(dce_latency_ctrl, seq_out) == dce::lat::kCtrlSeqOut * 4 &&
                  offsetof(dce_latency_ctrl, alive) == dce::lat::kCtrlAlive * 4 && offsetof(dce_latency_ctrl, cls0) == dce::lat::kCtrlCls0 * 4, "layout");
    memset((void*)ctrl, 0, sizeof(*ctrl));
    const Fp32Layout& L = w->f32;
    dce::lat::Weights wt;
    wt.w1 = at<float>(w, L.w1); wt.w2 = at<float>(w, L.w2); wt.w3 = at<float>(w, L.w3); wt.w4q = at<float>(w, L.w4q);
    wt.f1s = at<float>(w, L.f1s); wt.f2s = at<float>(w, L.f2s); wt.f3t = at<float>(w, L.f3t);
    for (int i = 0; i < 7; ++i) wt.b[i] = at<float>(w, L.b[i]);
    Ctx ctx; ctx.stream = (cudaStream_t)stream;
    rc = dce::lat::run(wt, w->sm_count, ring_dev, true, 0, 1, nullptr, nullptr, nullptr, (char*)workspace_dev, ctx, 1, 1,
                       reinterpret_cast<volatile unsigned*>(ctrl), (unsigned long long)(idle_timeout_s * 1e9), true);
    g_launches = ctx.launches;
    if (rc == DCE_ECUDA && ctx.err != cudaSuccess) g_last_cuda_error = (int)ctx.err;
    return rc;
}

static int profile_any(const dce_weights* w, const float* src, bool is_stream, int64_t T, int64_t first, int64_t n,
                       float* logits_dev, int32_t* cls_dev, uint8_t* bits_dev, void* workspace_dev, size_t workspace_bytes,
                       int precision, void* stream, int max_kernels, float* ms_out, const char** names_out, int* n_out) {
    if (!ms_out || !names_out || !n_out || max_kernels <= 0) return DCE_EINVAL;
    dce::Profiler prof;
    Ctx ctx; ctx.stream = (cudaStream_t)stream; ctx.prof = &prof;
    int rc = run_any(w, src, is_stream, T, first, n, logits_dev, cls_dev, bits_dev, workspace_dev, workspace_bytes, precision, ctx);
    g_launches = ctx.launches;
    cudaError_t se = cudaStreamSynchronize(ctx.stream);
    int cnt = 0;
    for (int i = 0; i < prof.n; ++i) {
        float ms = 0.f;
        if (se == cudaSuccess) cudaEventElapsedTime(&ms, prof.start[i], prof.stop[i]);
        if (cnt < max_kernels) { ms_out[cnt] = ms; names_out[cnt] = prof.names[i]; ++cnt; }
        cudaEventDestroy(prof.start[i]); cudaEventDestroy(prof.stop[i]);
    }
    *n_out = cnt;
    if (rc == DCE_OK && se != cudaSuccess) return cuda_fail(se);
    return rc;
}

int dce_forward_profile(const dce_weights* w, const float* x_dev, int64_t B,
                        float* logits_dev, int32_t* cls_dev, uint8_t* bits_dev,
                        void* workspace_dev, size_t workspace_bytes, int precision, void* stream,
                        int max_kernels, float* ms_out, const char** names_out, int* n_out) {
    return profile_any(w, x_dev, false, 0, 0, B, logits_dev, cls_dev, bits_dev, workspace_dev, workspace_bytes, precision,
                       stream, max_kernels, ms_out, names_out, n_out);
}

int dce_stream_profile(const dce_weights* w, const float* data_dev, int64_t T, int64_t first_window, int64_t n_windows,
                       float* logits_dev, int32_t* cls_dev, uint8_t* bits_dev,
                       void* workspace_dev, size_t workspace_bytes, int precision, void* stream,
                       int max_kernels, float* ms_out, const char** names_out, int* n_out) {
    return profile_any(w, data_dev, true, T, first_window, n_windows, logits_dev, cls_dev, bits_dev, workspace_dev,
                       workspace_bytes, precision, stream, max_kernels, ms_out, names_out, n_out);
}

int dce_weights_set_option(dce_weights* w, const char* key, int value) {
    if (!w || !key) return DCE_EINVAL;
    dce::tc::Options& o = w->opt;
    if (!strcmp(key, "fuse_block1")) { o.fuse_block1 = value; return DCE_OK; }
    if (!strcmp(key, "fuse_block2")) { o.fuse_block2 = value; return DCE_OK; }
    if (!strcmp(key, "fuse_fc3")) { o.fuse_fc3 = value; return DCE_OK; }
    if (!strcmp(key, "block2_dbg")) { o.block2_dbg = value; return DCE_OK; }
    if (!strcmp(key, "fuse_argmax")) { o.fuse_argmax = value; return DCE_OK; }
    if (!strcmp(key, "latency_kernel")) { o.latency_kernel = value; return DCE_OK; }
    if (!strcmp(key, "latency_coop")) { o.latency_coop = value; return DCE_OK; }
    if (!strcmp(key, "latency_tma_in")) { o.latency_tma_in = value; return DCE_OK; }
    if (!strcmp(key, "block1_dbg")) { o.block1_dbg = value; return DCE_OK; }
    if (!strcmp(key, "tapgemm_dbg")) { o.tapgemm_dbg = value; return DCE_OK; }
    if (!strcmp(key, "sm_limit")) { if (value < 0) return DCE_EINVAL; o.sm_limit = value; return DCE_OK; }
    if (!strcmp(key, "trace_layer")) { o.trace_layer = value; return DCE_OK; }   // -1: block1; 2..5: conv3, conv4, fc.0, fc.3; 6: block2
    if (!strcmp(key, "trace")) {                 // value != 0: allocate (once) and arm a 60-tile x 16-event clock64 trace of CTA 0
        int prev = 0;
        DCE_CUDA(cudaGetDevice(&prev));
        DCE_CUDA(cudaSetDevice(w->device));
        cudaError_t e = cudaSuccess;
        if (value && !o.trace) { e = cudaMalloc(&o.trace, 60 * 16 * 8); if (e == cudaSuccess) e = cudaMemset(o.trace, 0, 60 * 16 * 8); }
        if (!value && o.trace) { cudaFree(o.trace); o.trace = nullptr; }
        cudaSetDevice(prev);
        return e == cudaSuccess ? DCE_OK : cuda_fail(e);
    }
    return DCE_EINVAL;
}

int dce_debug_read_trace(dce_weights* w, long long* host_out, int n) {
    if (!w || !w->opt.trace || !host_out || n <= 0 || n > 60 * 16) return DCE_EINVAL;
    return cudaMemcpy(host_out, w->opt.trace, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? DCE_OK : DCE_ECUDA;
}

int dce_decimal2binary(const int64_t* cls_dev, int64_t n, uint8_t* bits_dev, void* stream) {
    if (n < 0) return DCE_EINVAL;
    if (n == 0) return DCE_OK;
    if (!cls_dev || !bits_dev) return DCE_EINVAL;
    if ((uintptr_t)cls_dev % 8 || (uintptr_t)bits_dev % 4) return DCE_EALIGN;
    Ctx ctx; ctx.stream = (cudaStream_t)stream;
    ctx.begin("decimal2binary");
    dce::fp32::decimal2binary_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx.stream>>>(cls_dev, n, bits_dev);
    const bool ok = ctx.end();
    g_launches = ctx.launches;
    return ok ? DCE_OK : cuda_fail(ctx.err);
}

int dce_ingest_f64(const double* src_dev, float* dst_dev, int64_t n, void* stream) {
    if (n < 0) return DCE_EINVAL;
    if (n == 0) return DCE_OK;
    if (!src_dev || !dst_dev) return DCE_EINVAL;
    if ((uintptr_t)src_dev % 16 || (uintptr_t)dst_dev % 8) return DCE_EALIGN;
    int64_t blocks = (n / 2 + 255) / 256; if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
    Ctx ctx; ctx.stream = (cudaStream_t)stream;
    ctx.begin("ingest_f64");
    dce::fp32::ingest_f64_kernel<<<(unsigned)blocks, 256, 0, ctx.stream>>>(src_dev, dst_dev, n);
    const bool ok = ctx.end();
    g_launches = ctx.launches;
    return ok ? DCE_OK : cuda_fail(ctx.err);
}

int dce_accuracy_counts(const int32_t* cls_dev, const int64_t* labels_dev, int64_t n, int64_t* counts_dev, void* stream) {
    if (n < 0) return DCE_EINVAL;
    if (n == 0) return DCE_OK;
    if (!cls_dev || !labels_dev || !counts_dev) return DCE_EINVAL;
    if ((uintptr_t)cls_dev % 4 || (uintptr_t)labels_dev % 8 || (uintptr_t)counts_dev % 8) return DCE_EALIGN;
    static_assert(dce::fp32::kNumCounts == DCE_NUM_COUNTS, "include/dce.h and the kernel agree on the counter layout");
    int64_t blocks = (n + 255) / 256; if (blocks > 592) blocks = 592;
    Ctx ctx; ctx.stream = (cudaStream_t)stream;
    ctx.begin("accuracy_counts");
    dce::fp32::accuracy_counts_kernel<<<(unsigned)blocks, 256, 0, ctx.stream>>>(
        cls_dev, labels_dev, n, reinterpret_cast<unsigned long long*>(counts_dev));
    const bool ok = ctx.end();
    g_launches = ctx.launches;
    return ok ? DCE_OK : cuda_fail(ctx.err);
}

}  // extern "C"
