// torch.ops.dce.* — the torch C++ extension over the C ABI of libdce_b200.so (include/dce.h).
//
// The reference's boundary for this path is the Python call `model(input_data)` inside
// `with torch.no_grad()` (/root/reference/src/inference_one_seq.py:25, src/test.py:87) and the
// loops around it (src/inference_one_seq.py:19-30).  This file registers those two entry
// points with the PyTorch dispatcher:
//
//     torch.ops.dce.forward(handle, x, workspace, precision, want_logits, want_cls, want_bits)
//         -> (logits (B,16) f32, cls (B,) i32, bits (B,4) u8)                = dce_forward
//     torch.ops.dce.stream(handle, data, first_window, n_windows, workspace, precision,
//                          want_logits, want_cls, want_bits)   -> same triple = dce_stream
//     torch.ops.dce.accuracy_counts(cls, labels, counts) -> counts            = dce_accuracy_counts
//
// `handle` is the dce_weights* as an int (ContactEngine owns it); outputs not asked for come
// back as empty tensors.  Nothing here computes: tensors are checked, outputs are allocated
// from PyTorch's caching allocator, the current CUDA stream is passed down, and a non-zero
// return code becomes a RuntimeError (TORCH_CHECK).  A Meta implementation gives shapes to
// fake-tensor tracing / torch.compile.  Links against libdce_b200.so ($ORIGIN rpath): the C ABI
// stays the only way into the kernels.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>

#include "../../include/dce.h"

namespace {

using at::Tensor;
using Triple = std::tuple<Tensor, Tensor, Tensor>;

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == DCE_OK, what, ": ", dce_strerror(rc), " (code ", rc, ")",
                rc == DCE_ECUDA ? ", cudaError " : "", rc == DCE_ECUDA ? std::to_string(dce_last_cuda_error()) : std::string());
}

Triple make_outputs(const Tensor& like, int64_t n, bool want_logits, bool want_cls, bool want_bits) {
    auto o = like.options();
    return Triple(at::empty({want_logits ? n : 0, DCE_CLASSES}, o.dtype(at::kFloat)),
                  at::empty({want_cls ? n : 0}, o.dtype(at::kInt)),
                  at::empty({want_bits ? n : 0, DCE_LEGS}, o.dtype(at::kByte)));
}

void check_workspace(const Tensor& ws, const Tensor& x, int64_t n, int64_t precision) {
    TORCH_CHECK(ws.is_cuda() && ws.device() == x.device() && ws.scalar_type() == at::kByte && ws.is_contiguous(),
                "dce: workspace must be a contiguous uint8 CUDA tensor on the input's device");
    TORCH_CHECK((size_t)ws.numel() >= dce_workspace_bytes(n, (int)precision), "dce: workspace too small: ", ws.numel(),
                " < ", dce_workspace_bytes(n, (int)precision), " bytes (dce_workspace_bytes)");
}

template <class T> T* ptr_or_null(Tensor& t) { return t.numel() ? t.data_ptr<T>() : nullptr; }

// contact_cnn.forward (+ torch.max + decimal2binary): src/contact_cnn.py:60-66, src/inference_one_seq.py:26-27
Triple forward_cuda(int64_t handle, const Tensor& x, const Tensor& workspace, int64_t precision,
                    bool want_logits, bool want_cls, bool want_bits) {
    TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kFloat, "dce::forward: x must be a float32 CUDA tensor");
    TORCH_CHECK(x.dim() == 3 && x.size(1) == DCE_WINDOW && x.size(2) == DCE_CHANNELS, "dce::forward: expected (B,",
                DCE_WINDOW, ",", DCE_CHANNELS, "), got ", x.sizes());
    Tensor xc = x.contiguous();
    if (reinterpret_cast<uintptr_t>(xc.data_ptr()) % 16) xc = xc.clone();      // a view at an odd offset: the ABI wants 16-byte alignment
    const int64_t n = xc.size(0);
    Triple out = make_outputs(xc, n, want_logits, want_cls, want_bits);
    if (n == 0) return out;
    check_workspace(workspace, xc, n, precision);
    const c10::cuda::CUDAGuard guard(xc.device());
    const cudaStream_t stream = c10::cuda::getCurrentCUDAStream(xc.device().index()).stream();
    check_rc(dce_forward(reinterpret_cast<const dce_weights*>(handle), xc.data_ptr<float>(), n,
                         ptr_or_null<float>(std::get<0>(out)), ptr_or_null<int32_t>(std::get<1>(out)),
                         ptr_or_null<uint8_t>(std::get<2>(out)), workspace.data_ptr(), (size_t)workspace.numel(),
                         (int)precision, stream),
             "dce_forward");
    return out;
}

// the body of inference(): src/inference_one_seq.py:19-30 over the resident log of utils/data_handler.py:26
Triple stream_cuda(int64_t handle, const Tensor& data, int64_t first_window, int64_t n_windows, const Tensor& workspace,
                   int64_t precision, bool want_logits, bool want_cls, bool want_bits) {
    TORCH_CHECK(data.is_cuda() && data.scalar_type() == at::kFloat, "dce::stream: data must be a float32 CUDA tensor");
    TORCH_CHECK(data.dim() == 2 && data.size(1) == DCE_CHANNELS, "dce::stream: expected (T,", DCE_CHANNELS, "), got ", data.sizes());
    Tensor dc = data.contiguous();
    if (reinterpret_cast<uintptr_t>(dc.data_ptr()) % 16) dc = dc.clone();      // e.g. log[k:] with odd k (rows are 216 B)
    const int64_t T = dc.size(0), total = T >= DCE_WINDOW ? T - DCE_WINDOW + 1 : 0;
    TORCH_CHECK(first_window >= 0 && n_windows >= 0 && first_window + n_windows <= total, "dce::stream: window range [",
                first_window, ", ", first_window + n_windows, ") outside [0, ", total, ")");
    Triple out = make_outputs(dc, n_windows, want_logits, want_cls, want_bits);
    if (n_windows == 0) return out;
    check_workspace(workspace, dc, n_windows, precision);
    const c10::cuda::CUDAGuard guard(dc.device());
    const cudaStream_t stream = c10::cuda::getCurrentCUDAStream(dc.device().index()).stream();
    check_rc(dce_stream(reinterpret_cast<const dce_weights*>(handle), dc.data_ptr<float>(), T, first_window, n_windows,
                        ptr_or_null<float>(std::get<0>(out)), ptr_or_null<int32_t>(std::get<1>(out)),
                        ptr_or_null<uint8_t>(std::get<2>(out)), workspace.data_ptr(), (size_t)workspace.numel(),
                        (int)precision, stream),
             "dce_stream");
    return out;
}

// fused eval counters: src/inference_one_seq.py:46-54, src/test.py:19-47,89-104
Tensor accuracy_counts_cuda(const Tensor& cls, const Tensor& labels, const Tensor& counts) {
    TORCH_CHECK(cls.is_cuda() && cls.scalar_type() == at::kInt && cls.is_contiguous(), "dce::accuracy_counts: cls must be contiguous int32 CUDA");
    TORCH_CHECK(labels.is_cuda() && labels.scalar_type() == at::kLong && labels.is_contiguous() && labels.numel() == cls.numel(),
                "dce::accuracy_counts: labels must be contiguous int64 CUDA of the same length");
    TORCH_CHECK(counts.is_cuda() && counts.scalar_type() == at::kLong && counts.is_contiguous() && counts.numel() >= DCE_NUM_COUNTS,
                "dce::accuracy_counts: counts must be int64[", DCE_NUM_COUNTS, "] on the device");
    const c10::cuda::CUDAGuard guard(cls.device());
    const cudaStream_t stream = c10::cuda::getCurrentCUDAStream(cls.device().index()).stream();
    check_rc(dce_accuracy_counts(cls.data_ptr<int32_t>(), labels.data_ptr<int64_t>(), cls.numel(), counts.data_ptr<int64_t>(), stream),
             "dce_accuracy_counts");
    return counts;
}

// shapes only (fake tensors, torch.compile tracing)
Triple forward_meta(int64_t, const Tensor& x, const Tensor&, int64_t, bool want_logits, bool want_cls, bool want_bits) {
    return make_outputs(x, x.size(0), want_logits, want_cls, want_bits);
}
Triple stream_meta(int64_t, const Tensor& data, int64_t, int64_t n_windows, const Tensor&, int64_t, bool want_logits, bool want_cls,
                   bool want_bits) {
    return make_outputs(data, n_windows, want_logits, want_cls, want_bits);
}

}  // namespace

TORCH_LIBRARY(dce, m) {
    m.def("forward(int handle, Tensor x, Tensor workspace, int precision, bool want_logits, bool want_cls, bool want_bits)"
          " -> (Tensor, Tensor, Tensor)");
    m.def("stream(int handle, Tensor data, int first_window, int n_windows, Tensor workspace, int precision, bool want_logits,"
          " bool want_cls, bool want_bits) -> (Tensor, Tensor, Tensor)");
    m.def("accuracy_counts(Tensor cls, Tensor labels, Tensor(a!) counts) -> Tensor(a!)");
}
TORCH_LIBRARY_IMPL(dce, CUDA, m) {
    m.impl("forward", &forward_cuda);
    m.impl("stream", &stream_cuda);
    m.impl("accuracy_counts", &accuracy_counts_cuda);
}
TORCH_LIBRARY_IMPL(dce, Meta, m) {
    m.impl("forward", &forward_meta);
    m.impl("stream", &stream_meta);
}
