/*
 * dce.h — C ABI of libdce_b200.so: the B200-native contact-classification path.
 *
 * The reference (UMich-CURLY/deep-contact-estimator) has no FFI / plugin layer
 * of its own; its hot path is the Python `contact_cnn` nn.Module plus three
 * tiny torch helpers.  This header is the boundary a binding would target.
 * Each entry point names the reference interface it replaces (file:line under
 * /root/reference).  The Python host side (deep_contact_estimator_b200/) binds
 * these with ctypes; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross the boundary;
 *   - every pointer named *_dev is a DEVICE pointer on the handle's device;
 *     the caller owns all input/output/workspace buffers (the library only
 *     owns the packed-weight buffer inside a dce_weights handle);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it,
 *     nothing synchronises, nothing allocates: calls are CUDA-graph capturable;
 *   - the handle's device must be the calling thread's current device when a
 *     call enqueues work (as for any launch on a cudaStream_t); only
 *     dce_weights_create() switches devices itself, and restores the caller's;
 *   - return value: 0 = DCE_OK, negative = DCE_E*; never throws or aborts;
 *   - a handle is immutable after dce_weights_pack() (its debugging switches,
 *     dce_weights_set_option, aside) and may be shared by host threads;
 *     dce_forward / dce_stream keep no state outside the handle and the
 *     caller's buffers and are re-entrant given distinct workspaces.
 */
#ifndef DCE_H_
#define DCE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCE_VERSION 100          /* 0.1.0 */

#if defined(__GNUC__)
#define DCE_API __attribute__((visibility("default")))
#else
#define DCE_API
#endif

/* model geometry — src/contact_cnn.py:10-58, window_size in the config yaml files */
#define DCE_WINDOW   150
#define DCE_CHANNELS 54
#define DCE_CLASSES  16
#define DCE_LEGS     4
#define DCE_NUM_PARAMS 14

/* error codes */
#define DCE_OK            0
#define DCE_EINVAL       -1      /* bad argument / shape / null pointer      */
#define DCE_EALIGN       -2      /* pointer not aligned as documented        */
#define DCE_EARCH        -3      /* device is not sm_100 (B200)              */
#define DCE_ECUDA        -4      /* CUDA runtime error; see dce_last_cuda_error */
#define DCE_ENOTPACKED   -5      /* handle has no packed weights yet         */
#define DCE_EWORKSPACE   -6      /* workspace too small                      */
#define DCE_EUNSUPPORTED -7      /* precision / mode not built               */

/* arithmetic mode of the forward pass */
#define DCE_PREC_FP32    0       /* fp32 FFMA kernels (exact-order-free fp32) */
#define DCE_PREC_BF16X3  1       /* tcgen05 bf16 hi/lo split, 3 MMAs, fp32 accumulate in TMEM */

typedef struct dce_weights dce_weights;

/* library / error introspection */
DCE_API int         dce_version(void);
DCE_API const char *dce_strerror(int code);
DCE_API int         dce_last_cuda_error(void);           /* cudaError_t of the last DCE_ECUDA on this thread */

/*
 * Weights handle — replaces `model = contact_cnn(); model.load_state_dict(...);
 * model.eval().to(device)` (src/inference_one_seq.py:153-156, src/test.py:130-134).
 */
DCE_API int dce_weights_create(dce_weights **out, int device);
DCE_API int dce_weights_destroy(dce_weights *w);

/*
 * K0: repack the 14 fp32 state_dict tensors (device pointers, contiguous, in
 * state_dict order: block1.0.weight, block1.0.bias, block1.2.weight,
 * block1.2.bias, block2.0.weight, block2.0.bias, block2.2.weight,
 * block2.2.bias, fc.0.weight, fc.0.bias, fc.3.weight, fc.3.bias, fc.6.weight,
 * fc.6.bias — src/contact_cnn.py:10-58) into the kernels' layouts: tap-major
 * conv weights, fc.0 columns permuted from c*37+t to t*128+c
 * (src/contact_cnn.py:64), bf16 hi/lo split images for the tensor-core path.
 */
DCE_API int dce_weights_pack(dce_weights *w, const float *const params_dev[DCE_NUM_PARAMS], void *stream);

/* The packed buffer, for a one-time ncclBroadcast to the other ranks.  After
 * writing the buffer externally call dce_weights_adopt() to mark it packed. */
DCE_API size_t dce_weights_packed_bytes(const dce_weights *w);
DCE_API void  *dce_weights_packed_ptr(dce_weights *w);
DCE_API int    dce_weights_adopt(dce_weights *w);

/* Scratch the caller must provide to dce_forward / dce_stream for up to
 * `max_windows` windows per call (the library chunks internally, so this
 * saturates at a fixed size). 256-byte aligned device memory, ZERO-FILLED
 * ONCE after allocation (cudaMemset): its first 256 bytes hold the grid-barrier
 * counters of the latency kernel, the next 256 the per-tile tickets of the batch
 * path's last kernel (every call leaves both at zero), and tape padding rows
 * must start finite.  One workspace serves one call at a time. */
DCE_API size_t dce_workspace_bytes(int64_t max_windows, int precision);

/*
 * K1: logits = contact_cnn.forward(x)  (src/contact_cnn.py:60-66), optionally
 * followed by `torch.max(output, 1)` (src/inference_one_seq.py:26) and
 * `decimal2binary` (src/inference_one_seq.py:59-62).
 *   x_dev       [B][150][54] fp32, contiguous (the DataLoader batch,
 *               src/inference_one_seq.py:24), 16-byte aligned
 *   logits_dev  [B][16] fp32 or NULL
 *   cls_dev     [B] int32 class 0..15 or NULL (first maximal index; NaN wins)
 *   bits_dev    [B][4] uint8 contact bits, MSB first = leg 0 (RF), or NULL
 * Latency mode: a call with B <= 4 is ONE cooperative fp32 kernel (the reference's
 * default batch_size 1 loop, config/inference_one_seq_params.yaml:10); for such
 * calls x_dev and the outputs may also be pinned HOST memory (cudaHostAlloc),
 * which the kernel reads and writes in place over PCIe.
 */
DCE_API int dce_forward(const dce_weights *w, const float *x_dev, int64_t B,
                float *logits_dev, int32_t *cls_dev, uint8_t *bits_dev,
                void *workspace_dev, size_t workspace_bytes, int precision, void *stream);

/*
 * K3 as a resident SERVER — the 1 kHz control loop (BASELINE configs[4]; one iteration of the reference loop at its
 * default batch_size 1, src/inference_one_seq.py:23-28) without a launch, a stream synchronisation or a copy engine
 * per step.  dce_latency_server_start() enqueues ONE cooperative kernel that stays resident on every SM and serves a
 * step each time the host rings the doorbell:
 *     write the n new windows to x_host;  ctrl->seq_in = ++seq;  spin until ctrl->seq_out == seq;
 *     read cls_host / bits_host / logits_host                          (dce_latency_server_step() below does this)
 * x_host [n][150][54] fp32, the three outputs and ctrl must be PINNED, device-mapped host memory (cudaHostAlloc);
 * NULL outputs are skipped.  The kernel retires when ctrl->quit is set non-zero or when no doorbell has arrived for
 * idle_timeout_s (ctrl->alive drops to 0; start it again for the next step), so a forgotten server cannot hold the
 * GPU; while it runs, other kernels on the device find no free SM.  Results are bit-identical to dce_forward on the
 * same windows.  workspace: as for dce_forward, for the server's exclusive use until it has retired.
 */
typedef struct dce_latency_ctrl {
    volatile uint32_t seq_in;        /* host -> device: step number requested (start at 0, +1 per step)       */
    volatile uint32_t quit;          /* host -> device: non-zero = retire                                      */
    uint32_t reserved0[14];
    volatile uint32_t seq_out;       /* device -> host: last step whose results are in the host buffers        */
    volatile int32_t  cls0;          /* result words: pass cls_host = &ctrl->cls0, bits_host = ctrl->bits0 and      */
    volatile uint8_t  bits0[4];      /*   logits_host = NULL (n = 1) and the results arrive with seq_out in ONE        */
    volatile uint32_t device_ns;     /*   16-byte store (no system fence: ~1 us less); device_ns = doorbell seen -> results written */
    volatile uint32_t alive;         /* device -> host: 1 while the server runs                                */
    uint32_t reserved1[11];
} dce_latency_ctrl;                  /* 128 bytes: one cache line per direction                                */

DCE_API int dce_latency_server_start(const dce_weights *w, const float *x_host, int n,
                             float *logits_host, int32_t *cls_host, uint8_t *bits_host,
                             dce_latency_ctrl *ctrl, void *workspace_dev, size_t workspace_bytes,
                             double idle_timeout_s, void *stream);

/* One step of a running server (host side only; no CUDA call).  Returns 0, or -1 if the server is not alive. */
static inline int dce_latency_server_step(dce_latency_ctrl *ctrl) {
    const uint32_t seq = ctrl->seq_in + 1u;
    if (!ctrl->alive) return -1;
    __atomic_store_n(&ctrl->seq_in, seq, __ATOMIC_RELEASE);          /* the window is written before the doorbell */
    while (__atomic_load_n(&ctrl->seq_out, __ATOMIC_ACQUIRE) != seq)
        if (!ctrl->alive) return -1;
    return 0;
}

/*
 * The same server fed ONE NEW SENSOR ROW per step — what a 1 kHz estimator actually receives (README.md:67-83 of the
 * reference: one leg_control_data + microstrain sample per tick).  The device keeps the last 150 rows in a ring
 * (ring_dev: [300][54] fp32 DEVICE memory owned by the caller; row k lives at slots k % 150 and k % 150 + 150, so the
 * newest window is always one contiguous slice) and z-scores the window itself exactly as stream mode does
 * (utils/data_handler.py:55-56): 216 bytes cross PCIe per step instead of 32 400 and the host does no arithmetic.
 * Protocol: the host writes the row as 18 chunks {3 floats, tag} and a 19th chunk {quit, slot, 0, tag}, every tag =
 * the step number (from 1; floats before the tag of their chunk) — dce_latency_row_server_push() below; the device
 * polls all chunks in one PCIe round trip, so doorbell and data arrive together.  Results come back in the control
 * block (seq_out, cls0, bits0 in one 16-byte store).  Retirement as for dce_latency_server_start; the ring survives
 * a restart (keep passing consecutive slots).
 */
typedef struct dce_latency_row_ctrl {
    volatile uint32_t chunk[19][4];  /* [i < 18] = {row[3i], row[3i+1], row[3i+2] as float bits, tag}; [18] = {quit, slot, 0, tag} */
    uint32_t reserved0[4];
    volatile uint32_t seq_out;       /* offset 320: {seq_out, cls0, bits0, device_ns} arrive as one store          */
    volatile int32_t  cls0;
    volatile uint8_t  bits0[4];
    volatile uint32_t device_ns;
    volatile uint32_t alive;
    uint32_t reserved1[43];
} dce_latency_row_ctrl;              /* 512 bytes */

DCE_API int dce_latency_row_server_start(const dce_weights *w, dce_latency_row_ctrl *ctrl, float *ring_dev,
                                 void *workspace_dev, size_t workspace_bytes, double idle_timeout_s, void *stream);

/* Push row `seq` (1, 2, ...) into ring slot `slot` (0..149, consecutive) and wait for its class / bits.  Host side only. */
static inline int dce_latency_row_server_push(dce_latency_row_ctrl *ctrl, const float row[54], uint32_t slot, uint32_t seq) {
    int i;
    if (!ctrl->alive) return -1;
    for (i = 0; i < 18; ++i) {
        union { float f; uint32_t u; } a, b, c;
        a.f = row[3 * i]; b.f = row[3 * i + 1]; c.f = row[3 * i + 2];
        ctrl->chunk[i][0] = a.u; ctrl->chunk[i][1] = b.u; ctrl->chunk[i][2] = c.u;
    }
    ctrl->chunk[18][1] = slot;
    __atomic_thread_fence(__ATOMIC_RELEASE);                          /* data before tags */
    for (i = 0; i < 19; ++i) ctrl->chunk[i][3] = seq;
    while (__atomic_load_n(&ctrl->seq_out, __ATOMIC_ACQUIRE) != seq)
        if (!ctrl->alive) return -1;
    return 0;
}

/*
 * K2: the body of `inference(dataloader, model, device)`
 * (src/inference_one_seq.py:19-30) over a device-resident sensor log: for
 * window i in [first_window, first_window + n_windows): rows i..i+149 of
 * `data_dev` ([T][54] fp32, utils/data_handler.py:26), minus the column mean,
 * divided by the unbiased column std (utils/data_handler.py:55-56), forward,
 * argmax, bits.  Overlapping windows are read once per tile from the
 * contiguous stream.  Outputs are indexed from 0 (= first_window).
 */
DCE_API int dce_stream(const dce_weights *w, const float *data_dev, int64_t T,
               int64_t first_window, int64_t n_windows,
               float *logits_dev, int32_t *cls_dev, uint8_t *bits_dev,
               void *workspace_dev, size_t workspace_bytes, int precision, void *stream);

/*
 * Profiling twin of dce_forward: the same launches with a cudaEvent pair around
 * every kernel; synchronises `stream`, then reports each kernel's duration.
 * names_out[i] are static strings.  Used by bench.py for the roofline figure.
 */
DCE_API int dce_forward_profile(const dce_weights *w, const float *x_dev, int64_t B,
                        float *logits_dev, int32_t *cls_dev, uint8_t *bits_dev,
                        void *workspace_dev, size_t workspace_bytes, int precision, void *stream,
                        int max_kernels, float *ms_out, const char **names_out, int *n_out);

DCE_API int dce_stream_profile(const dce_weights *w, const float *data_dev, int64_t T,
                        int64_t first_window, int64_t n_windows,
                        float *logits_dev, int32_t *cls_dev, uint8_t *bits_dev,
                        void *workspace_dev, size_t workspace_bytes, int precision, void *stream,
                        int max_kernels, float *ms_out, const char **names_out, int *n_out);

/*
 * `decimal2binary(x)` alone (src/inference_one_seq.py:59-62, src/test.py:109-111):
 * cls_dev [n] int64 -> bits_dev [n][4] uint8 (used for ground-truth labels).
 */
DCE_API int dce_decimal2binary(const int64_t *cls_dev, int64_t n, uint8_t *bits_dev, void *stream);

/*
 * Device-side ingest of the on-disk log format: `np.load(data_path)` is float64 (utils/mat2numpy.py:73,80 save
 * float64 .npy) and `contact_dataset.__init__` converts it to float32 on the host before the upload
 * (utils/data_handler.py:21-27).  Here the float64 rows are uploaded in chunks and converted on the device,
 * round-to-nearest as the host cast:  dst_dev[i] = (float)src_dev[i], i < n.  src 16-byte, dst 8-byte aligned.
 */
DCE_API int dce_ingest_f64(const double *src_dev, float *dst_dev, int64_t n, void *stream);

/*
 * Fused evaluation counters of `inference_and_compute_acc` / `compute_accuracy` and of test.py's metrics
 * (src/inference_one_seq.py:33-57, src/test.py:19-107): given predicted classes and labels, ACCUMULATE into
 * counts_dev (int64[DCE_NUM_COUNTS], zeroed by the caller):
 *   [0]                      #(pred == label)
 *   [1 + leg]                per-leg contact-bit agreement, leg 0..3 (RF, LF, RH, LH)
 *   [5 + 4 leg + 2 gt + pr]  the four 2x2 per-leg confusion matrices, rows = ground-truth bit, columns = predicted
 *                            bit — sklearn's confusion_matrix(gt, pred, labels=[0,1]) of src/test.py:19-27
 *   [21 + 16 gt + pr]        the 16x16 class confusion matrix (precision_score / jaccard_score of src/test.py:50-70
 *                            are functions of it)
 */
#define DCE_NUM_COUNTS 277
DCE_API int dce_accuracy_counts(const int32_t *cls_dev, const int64_t *labels_dev, int64_t n,
                        int64_t *counts_dev, void *stream);

/*
 * Ablation / debugging switches OF ONE HANDLE (not process-wide: two handles never see each other's switches).
 * Not part of the hot path: set them between calls, from one thread, while no call on the handle is in flight.  Keys:
 *   "fuse_block1"  1 (default): ingest + conv1 + conv2 + pool run as ONE kernel;
 *                  0: one kernel per layer (activations round-trip through HBM);
 *   "fuse_block2"  1 (default): conv3 + conv4 + pool run as ONE kernel (X3 stays in shared memory); 0: two launches;
 *   "fuse_fc3"     1 (default): fc.6 is folded into fc.3's epilogue (logit shares + a small reduce/argmax kernel);
 *                  0: fc.3 writes H2, a separate kernel does fc.6 + argmax + bits;
 *   "latency_kernel" 1 (default): calls of <= 4 windows run the single cooperative latency kernel;
 *                  0: the per-layer kernels (tensor-core convolutions + fp32 GEMV Linear layers);
 *   "latency_coop" / "latency_tma_in"  launch attribute / input staging ablations of that kernel;
 *   "sm_limit"     0 (default): use every SM; n > 0: the batch kernels launch at most n CTAs (what a MIG slice or a
 *                  smaller sm_100 part gives them: several tiles per CTA in every layer);
 *   "block1_dbg" / "block2_dbg" / "tapgemm_dbg"  bit masks of timing ablations inside the kernels (results invalid; the
 *                  bits are listed next to the parameter structs in csrc/);
 *   "trace" 1: allocate and arm a per-role clock64 timeline of CTA 0 (libraries built with -DDCE_TRACE=1);
 *   "trace_layer"  which kernel records it: -1 block1 (default), 2..5 conv3 / conv4 / fc.0 / fc.3, 6 block2.
 * Returns DCE_EINVAL for an unknown key.
 */
DCE_API int dce_weights_set_option(dce_weights *w, const char *key, int value);
/* "trace" armed: copy the first n clock64 samples ([tile][16 events]) of CTA 0 to the host. */
DCE_API int dce_debug_read_trace(dce_weights *w, long long *host_out, int n);

/* How many kernel launches the last dce_forward / dce_stream on this thread enqueued. */
DCE_API int dce_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DCE_H_ */
