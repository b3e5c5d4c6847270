// Host-side launch sequence of the tensor-core path (all layers).
#pragma once
#include "dce_tc.cuh"
#include "dce_tc_block1.cuh"
#include "dce_tc_block2.cuh"
#include "dce_small.cuh"

namespace dce {
namespace tc {

// Ablation / debugging switches.  They live in the dce_weights handle (dce_weights_set_option), not in process-wide
// state: two handles in one process do not see each other's switches, and dce_forward / dce_stream only read them.
struct Options {
    int fuse_block1 = 1;         // 0: ingest, conv1, conv2 as separate launches (activations round-trip through HBM)
    int fuse_block2 = 1;         // 0: conv3, conv4 as separate launches
    int fuse_fc3 = 1;            // 0: fc.3 writes H2, a separate kernel does fc.6 + argmax + bits
    int fuse_argmax = 1;         // 0: a separate kernel adds fc.3's logit shares and takes the argmax
    int latency_kernel = 1;      // 0: calls of <= 4 windows take the per-layer kernels
    int latency_coop = 1, latency_tma_in = 1;
    int block1_dbg = 0;          // timing ablations inside block1_kernel (results invalid)
    int tapgemm_dbg = 0;
    int block2_dbg = 0;
    int sm_limit = 0;            // > 0: the batch kernels use at most this many CTAs (what a MIG slice / smaller part would give them)
    int trace_layer = -1;        // which kernel records into `trace`: -1 block1, 2..5 conv3 / conv4 / fc.0 / fc.3, 6 block2
    long long* trace = nullptr;  // device buffer [60 tiles][16 events] of clock64 samples of CTA 0 (DCE_TRACE builds)
};

// pointers into the fp32 section of the packed buffer, passed in by dce.cu
struct BiasPtrs { const float* b[7]; const float* w3; const float* f1; const float* f2; };

inline int run(const char* buf, const PackedLayout& L, const BiasPtrs& bp, const Options& opt, int sm_count, const float* src,
               bool stream_mode, int64_t total_rows, int64_t first, int64_t n, float* logits, int32_t* cls, uint8_t* bits,
               char* ws, Ctx& ctx) {
    cudaStream_t s = ctx.stream;
    if (opt.sm_limit > 0 && opt.sm_limit < sm_count) sm_count = opt.sm_limit;
    {
        static DeviceOnce fc3_once;
        if (auto first_ = fc3_once.need()) {
            cudaError_t e = cudaFuncSetAttribute(fp32::fc3_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fp32::kFc3SmemBytes);
            if (e != cudaSuccess) { ctx.err = e; return DCE_ECUDA; }
        }
    }
    int m = 0;
    for (int64_t c0 = 0; c0 < n; c0 += m) {
        // chunks of kChunk windows; a tail of <= 4 windows would take the small-batch Linear kernels (fp32 GEMV: the same
        // logits to rounding, not to the bit), so the chunk before it gives up 64 windows: a window's result never
        // depends on where it sits in a call of more than 4 windows
        const int64_t rem = n - c0;
        m = (int)(rem <= kChunk ? rem : (rem - kChunk <= small::kMaxB ? kChunk - 64 : kChunk));
        const Workspace W = make_workspace(m);
        uint8_t* x0 = reinterpret_cast<uint8_t*>(ws + W.o_x0);
        uint8_t* x1 = reinterpret_cast<uint8_t*>(ws + W.o_x1);
        uint8_t* x2 = reinterpret_cast<uint8_t*>(ws + W.o_x2);
        uint8_t* x3 = reinterpret_cast<uint8_t*>(ws + W.o_x3);
        uint8_t* x4 = reinterpret_cast<uint8_t*>(ws + W.o_x4);
        uint8_t* h1 = reinterpret_cast<uint8_t*>(ws + W.o_h1);
        float* h2 = reinterpret_cast<float*>(ws + W.o_h2);
        float* mean = reinterpret_cast<float*>(ws + W.o_mean);
        float* sdev = reinterpret_cast<float*>(ws + W.o_sdev);

        int rc;
        TapGemmParams p{};
        const bool tiny = m <= small::kMaxB;       // latency mode: one M-tile per CTA tile (the second would be padding)
        if (stream_mode) {
            static DeviceOnce st_once;
            if (auto first_ = st_once.need()) {
                cudaError_t e = cudaFuncSetAttribute(window_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatSmemBytes);
                if (e != cudaSuccess) { ctx.err = e; return DCE_ECUDA; }
            }
            const int64_t fa = first + c0;
            const int tiles = (int)((fa + m - 1) / kStatWT - fa / kStatWT + 1);       // tiles are aligned to absolute window indices
            DCE_KL(ctx, "tc_window_stats", { cudaError_t le_ = launch_pdl(window_stats_kernel, dim3(tiles), dim3(256), kStatSmemBytes, s, src, fa, m, total_rows, mean, sdev, opt.fuse_block1 ? 1 : 0); (void)le_; });
        }
        if (opt.fuse_block1) {
            // ---- fused ingest + conv1 + conv2 + pool (a2-a6): windows -> X2
            static DeviceOnce attr_once;
            static int max_pairs[64] = {};                   // per device: how many 2-CTA clusters of this kernel are resident at once
            int dev_ = 0;
            cudaGetDevice(&dev_);
            dev_ &= 63;
            if (auto first_ = attr_once.need()) {
                cudaError_t e = cudaFuncSetAttribute(block1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB1SmemBytes);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(block1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB1SmemBytes);
                int n_ = 0;
                if (e == cudaSuccess) {
                    cudaLaunchConfig_t cfg{};
                    cfg.gridDim = dim3(2); cfg.blockDim = dim3(kB1Threads); cfg.dynamicSmemBytes = kB1SmemBytes;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    e = cudaOccupancyMaxActiveClusters(&n_, block1_kernel<false>, &cfg);
                }
                if (e != cudaSuccess || n_ < 1) { first_.fail(); ctx.err = (e != cudaSuccess) ? e : cudaErrorLaunchOutOfResources; return DCE_ECUDA; }
                max_pairs[dev_] = n_;
            }
            Block1Params b{};
            b.x = stream_mode ? src : src + (size_t)c0 * 150 * 54; b.first = first + c0; b.n_windows = m; b.total_rows = total_rows;
            b.mean = mean; b.rstd = sdev;   // window_stats_kernel<true> writes 1/std
            b.w1 = reinterpret_cast<const uint8_t*>(buf + L.w[kLayerConv1Stack]); b.w2 = reinterpret_cast<const uint8_t*>(buf + L.w[kLayerConv2Stack]);
            b.b1 = bp.b[0]; b.b2 = bp.b[1];
            b.out = x2; b.out_part_stride = W.x2.part_stride; b.out_kch_stride = W.x2.kch_stride; b.out_rows_cap = W.x2.m_tiles * 128;
            b.n_tiles = (m * kRW1 + kB1Rows - 1) / kB1Rows;
            b.dbg = opt.block1_dbg; b.trace = (opt.trace_layer < 0) ? opt.trace : nullptr;
            int pairs = (b.n_tiles + 1) / 2 < sm_count / 2 ? (b.n_tiles + 1) / 2 : sm_count / 2;     // CTA pairs: two tiles in lockstep
            if (pairs > max_pairs[dev_]) pairs = max_pairs[dev_];       // a TPC with one SM fused off hosts no pair: never queue a persistent cluster behind another
            if (pairs < 1) pairs = 1;
            const int grid = 2 * pairs;
            if (stream_mode)
                DCE_KL(ctx, "tc_block1_stream", { cudaError_t le_ = launch_pdl_pairs(block1_kernel<true>, dim3(grid), dim3(kB1Threads), kB1SmemBytes, s, b); (void)le_; });
            else
                DCE_KL(ctx, "tc_block1", { cudaError_t le_ = launch_pdl_pairs(block1_kernel<false>, dim3(grid), dim3(kB1Threads), kB1SmemBytes, s, b); (void)le_; });
        } else {
        // ---- ingest (a2/a3/a4): windows -> X0 tape
        const int iblocks = W.x0.m_tiles * 4;
        if (stream_mode) {
            DCE_KL(ctx, "tc_ingest_stream", ingest_kernel<true><<<iblocks, 256, 0, s>>>(
                src, first + c0, m, mean, sdev, x0, W.x0.part_stride, W.x0.kch_stride));
        } else {
            DCE_KL(ctx, "tc_ingest", ingest_kernel<false><<<iblocks, 256, 0, s>>>(
                src + (size_t)c0 * 150 * 54, 0, m, nullptr, nullptr, x0, W.x0.part_stride, W.x0.kch_stride));
        }
        // ---- conv1 (a5): X0 -> X1
        p = TapGemmParams{};
        p.a_tape = x0; p.a_part_stride = W.x0.part_stride; p.a_kch_stride = W.x0.kch_stride;
        p.w_packed = reinterpret_cast<const uint8_t*>(buf + L.w[0]); p.bias = bp.b[0];
        p.m_tiles = W.x0.m_tiles; p.n_tiles = 1; p.stages = kLayers[0].stages;
        p.out = x1; p.out_part_stride = W.x1.part_stride; p.out_kch_stride = W.x1.kch_stride; p.out_rows_cap = W.x1.m_tiles * 128;
        p.N = 64; p.rw = kRW1; p.tv = 150;
        if ((rc = launch_layer<64, 3, 4, 4, EPI_TAPE>(ctx, "tc_conv1", sm_count, p)) != DCE_OK) return rc;
        // ---- conv2 + pool (a6): X1 -> X2
        p.a_tape = x1; p.a_part_stride = W.x1.part_stride; p.a_kch_stride = W.x1.kch_stride;
        p.w_packed = reinterpret_cast<const uint8_t*>(buf + L.w[1]); p.bias = bp.b[1];
        p.m_tiles = W.x1.m_tiles; p.stages = kLayers[1].stages;
        p.out = x2; p.out_part_stride = W.x2.part_stride; p.out_kch_stride = W.x2.kch_stride; p.out_rows_cap = W.x2.m_tiles * 128;
        if ((rc = launch_layer<64, 3, 4, 4, EPI_POOL_TAPE>(ctx, "tc_conv2_pool", sm_count, p)) != DCE_OK) return rc;
        }
        if (opt.fuse_block2 && !tiny) {
            // ---- fused conv3 + conv4 + pool + flatten (a7-a9): X2 -> X4, X3 stays in shared memory
            static DeviceOnce b2_once;
            if (auto first_ = b2_once.need()) {
                cudaError_t e = cudaFuncSetAttribute(block2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kB2SmemBytes);
                if (e != cudaSuccess) { first_.fail(); ctx.err = e; return DCE_ECUDA; }
            }
            Block2Params b{};
            b.x2 = x2; b.x2_part_stride = W.x2.part_stride; b.x2_kch_stride = W.x2.kch_stride; b.n_windows = m;
            b.w3 = reinterpret_cast<const uint8_t*>(buf + L.w[kLayerConv3Stack]); b.w4 = reinterpret_cast<const uint8_t*>(buf + L.w[kLayerConv4Stack]);
            b.b3 = bp.b[2]; b.b4 = bp.b[3];
            b.out = x4; b.out_part_stride = W.x4.part_stride; b.out_rows_cap = W.x4.cap;
            b.n_tiles = (m * kRW2 + kB2Rows - 1) / kB2Rows;
            b.trace = (opt.trace_layer == 6) ? opt.trace : nullptr;
            b.dbg = opt.block2_dbg;
            const int grid = b.n_tiles < sm_count ? b.n_tiles : sm_count;
            DCE_KL(ctx, "tc_block2", { cudaError_t le_ = launch_pdl(block2_kernel, dim3(grid), dim3(kB2Threads), kB2SmemBytes, s, b); (void)le_; });
        } else {
        p = TapGemmParams{};
        p.n_tiles = 1;
        // ---- conv3 (a7): X2 -> X3
        p.a_tape = x2; p.a_part_stride = W.x2.part_stride; p.a_kch_stride = W.x2.kch_stride;
        p.w_packed = reinterpret_cast<const uint8_t*>(buf + L.w[2]); p.bias = bp.b[2];
        p.m_tiles = W.x2.m_tiles; p.stages = kLayers[2].stages;
        p.out = x3; p.out_part_stride = W.x3.part_stride; p.out_kch_stride = W.x3.kch_stride; p.out_rows_cap = W.x3.m_tiles * 128;
        p.N = 128; p.rw = kRW2; p.tv = 75;
        p.dbg = opt.tapgemm_dbg;
        p.trace = (opt.trace_layer == 2) ? opt.trace : nullptr;
        if (tiny) p.m_tiles = (m * kRW2 + 127) / 128;
        rc = tiny ? launch_layer<128, 3, 4, 3, EPI_TAPE, 1, 2>(ctx, "tc_conv3", sm_count, p)
                  : launch_layer<128, 3, 4, 3, EPI_TAPE, 2, 2>(ctx, "tc_conv3", sm_count, p);
        if (rc != DCE_OK) return rc;
        // ---- conv4 + pool + flatten (a8, a9): X3 -> X4 (fc.0 operand layout, k' = t*128 + c)
        p.a_tape = x3; p.a_part_stride = W.x3.part_stride; p.a_kch_stride = W.x3.kch_stride;
        p.w_packed = reinterpret_cast<const uint8_t*>(buf + L.w[3]); p.bias = bp.b[3];
        p.m_tiles = W.x3.m_tiles; p.stages = kLayers[3].stages;
        p.out = x4; p.out_part_stride = W.x4.part_stride; p.out_kch_stride = W.x4.kch_stride; p.out_rows_cap = W.x4.cap;
        p.trace = (opt.trace_layer == 3) ? opt.trace : nullptr;
        if (tiny) p.m_tiles = (m * kRW2 + 127) / 128;
        rc = tiny ? launch_layer<128, 3, 2, 6, EPI_POOL_FC, 1>(ctx, "tc_conv4_pool", sm_count, p)
                  : launch_layer<128, 3, 2, 4, EPI_POOL_FC, 2>(ctx, "tc_conv4_pool", sm_count, p);
        if (rc != DCE_OK) return rc;
        }
        if (m <= small::kMaxB) {
            // ---- latency mode (K3): fc.0 / fc.3 as split-N fp32 GEMVs over the fp32 weight images
            using namespace small;
            float* h1f = reinterpret_cast<float*>(h1);            // [m][2048] fp32 (aliases the unused H1 tape)
            static DeviceOnce gemv_once;
            auto k1 = gemv_bias_relu_kernel<4736, 2048, 16, true>;
            auto k2 = gemv_bias_relu_kernel<2048, 512, 8, false>;
            if (auto first_ = gemv_once.need()) {
                cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxB * 4736 * 4);
                if (e != cudaSuccess) { ctx.err = e; return DCE_ECUDA; }
            }
            DCE_KL(ctx, "fc1_gemv", k1<<<2048 / 16, 256, m * 4736 * 4, s>>>(x4, W.x4.part_stride, W.x4.kch_stride, m, bp.f1, bp.b[4], h1f));
            DCE_KL(ctx, "fc2_gemv", k2<<<512 / 8, 256, m * 2048 * 4, s>>>(h1f, 0, 0, m, bp.f2, bp.b[5], h2));
            DCE_KL(ctx, "fc3_argmax_bits", fp32::fc3_argmax_kernel<<<1, 256, fp32::kFc3SmemBytes, s>>>(
                h2, bp.w3, bp.b[6], m, logits ? logits + c0 * 16 : nullptr, cls ? cls + c0 : nullptr, bits ? bits + c0 * 4 : nullptr));
            continue;
        }
        // ---- fc.0 + ReLU (a10): X4 -> H1
        p = TapGemmParams{};
        p.a_tape = x4; p.a_part_stride = W.x4.part_stride; p.a_kch_stride = W.x4.kch_stride;
        p.w_packed = reinterpret_cast<const uint8_t*>(buf + L.w[4]); p.bias = bp.b[4];
        p.m_tiles = W.x4.m_tiles; p.n_tiles = kLayers[4].n_tiles; p.stages = kLayers[4].stages;
        p.out = h1; p.out_part_stride = W.h1.part_stride; p.out_kch_stride = W.h1.kch_stride; p.out_rows_cap = W.h1.m_tiles * 128;
        p.N = 2048; p.rw = 1; p.tv = 1;
        p.dbg = opt.tapgemm_dbg;
        p.trace = (opt.trace_layer == 4) ? opt.trace : nullptr;
        {   // the row-major operand [part][window][4736] as a 3-D tensor; a box is 128 windows x 32 K-elements (64 B) of one part
            ptx::EncodeTiledFn enc = ptx::encode_tiled_fn();
            if (!enc) return DCE_EUNSUPPORTED;
            const cuuint64_t dims[3] = {4736, (cuuint64_t)W.x4.cap, 2};
            const cuuint64_t strides[2] = {4736 * 2, (cuuint64_t)W.x4.part_stride};
            const cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
            if (enc(&p.a_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x4, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return DCE_EINVAL;
        }
        rc = launch_layer<240, 1, 4, 3, EPI_FC_TAPE, 2, 0, 1>(ctx, "tc_fc1", sm_count, p);
        if (rc != DCE_OK) return rc;
        // ---- fc.3 + ReLU (a11): H1 -> H2 fp32
        p.a_tape = h1; p.a_part_stride = W.h1.part_stride; p.a_kch_stride = W.h1.kch_stride;
        p.w_packed = reinterpret_cast<const uint8_t*>(buf + L.w[5]); p.bias = bp.b[5];
        p.m_tiles = W.h1.m_tiles; p.n_tiles = kLayers[5].n_tiles; p.stages = kLayers[5].stages;
        p.out = nullptr; p.out_f32 = h2; p.N = 512; p.n_valid = m;
        p.trace = (opt.trace_layer == 5) ? opt.trace : nullptr;
        if (opt.fuse_fc3) {
            // ---- fc.3 + ReLU with fc.6 folded into the epilogue (a11, a12): H2 stays in registers; 8 logit shares per window
            p.w3t = bp.w3;
            if (opt.fuse_argmax) {
                // ---- ... and the share reduction, argmax and contact bits (a12-a14) folded into the same launch: the CTA
                // that finishes an M-tile's last n-tile does them
                p.tickets = reinterpret_cast<unsigned*>(ws + W.o_tickets); p.b3 = bp.b[6];
                p.logits = logits ? logits + c0 * 16 : nullptr; p.cls = cls ? cls + c0 : nullptr; p.bits = bits ? bits + c0 * 4 : nullptr;
                if ((rc = launch_layer<128, 1, 4, 6, EPI_FC_LOGITS, 1, 0, 0, 2>(ctx, "tc_fc2_fc3_argmax", sm_count, p)) != DCE_OK) return rc;
                continue;
            }
            rc = launch_layer<128, 1, 4, 6, EPI_FC_LOGITS, 1, 0, 0, 2>(ctx, "tc_fc2_fc3", sm_count, p);
            if (rc != DCE_OK) return rc;
            DCE_KL(ctx, "logits_argmax_bits", { cudaError_t le_ = launch_pdl(fp32::logit_shares_argmax_kernel, dim3((m + 127) / 128), dim3(128), 0, s,
                (const float*)h2, bp.b[6], (int64_t)m, 2 * kLayers[5].n_tiles, logits ? logits + c0 * 16 : nullptr, cls ? cls + c0 : nullptr,
                bits ? bits + c0 * 4 : nullptr); (void)le_; });
            continue;
        }
        if ((rc = launch_layer<128, 1, 4, 6, EPI_FC_F32, 1>(ctx, "tc_fc2", sm_count, p)) != DCE_OK) return rc;
        // ---- fc.6 + argmax + bits (a12-a14), fp32 CUDA cores (16 K FLOP per window)
        const int g3 = (int)((m + 15) / 16 < sm_count * 2 ? (m + 15) / 16 : sm_count * 2);
        DCE_KL(ctx, "fc3_argmax_bits", { cudaError_t le_ = launch_pdl(fp32::fc3_argmax_kernel, dim3(g3), dim3(256), fp32::kFc3SmemBytes, s,
            (const float*)
            h2, bp.w3, bp.b[6], (int64_t)m, logits ? logits + c0 * 16 : nullptr, cls ? cls + c0 : nullptr, bits ? bits + c0 * 4 : nullptr); (void)le_; });
    }
    return DCE_OK;
}

}  // namespace tc
}  // namespace dce
