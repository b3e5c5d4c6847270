// DCE_PREC_BF16X3: the tcgen05 tensor-core path.
//
// Every layer of contact_cnn (/root/reference/src/contact_cnn.py:10-58) is one
// launch of the same persistent, warp-specialised kernel `tapgemm_kernel`:
//
//     D[row][n] = sum_{tap < TAPS} sum_{c} A[row + tap - 1][c] * W[n][tap][c]
//
// with TAPS = 3 for the Conv1d(k=3, pad=1) layers and TAPS = 1 for the Linear
// layers.  Arithmetic is split-bf16 ("bf16x3"): every fp32 operand x is stored as
// hi = bf16(x), lo = bf16(x - hi) and each K-step issues three tcgen05.mma
// (hi*hi + hi*lo + lo*hi) accumulating in fp32 in TMEM.  Plain bf16 or TF32 miss
// the 1e-4 parity bar; the dropped lo*lo term is ~2^-16 relative.
//
// Activation "tapes".  Activations live channels-last, as bf16 hi/lo, in
//     tape[part][kchunk = c / 8][row][8]                       (16 bytes per (kchunk,row))
// where for conv layers all windows of a chunk are stacked along `row` with
// RW rows per window (150 samples + 2 zero guard rows -> 152; after pooling
// 75 + 1 -> 76), so a conv tap is a shift by one row and window-edge padding
// is a guard row.  A 128-row M-tile plus its one-row halo is, per kchunk, ONE
// contiguous 2080-byte span: it is fetched with 1-D bulk TMA (cp.async.bulk)
// straight into the UMMA SWIZZLE_NONE K-major layout, and the three taps read
// the same staged slab through descriptors whose start address differs by
// 16 bytes (the folded im2col).  Weights are pre-packed (K0) into the exact
// shared-memory image of each (n-tile, k-stage) so a stage's B operand is one
// bulk copy.
//
// Roles per CTA (416 threads, 1 CTA/SM, persistent over tiles):
//   warps 0-7 : epilogue      (256 threads)  tcgen05.ld -> bias/ReLU/pool -> bf16 hi/lo -> next tape
//   warps 8.. : MMA issuers   (one lane each, one warp per accumulator) tcgen05.mma + tcgen05.commit; warp 8 owns TMEM alloc
//   last 4    : TMA producers (one lane each) smem ring, full/empty mbarriers
// TMEM holds two accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>
#include "dce_common.cuh"
#include "dce_tc_ptx.cuh"
#include "dce_fp32.cuh"

namespace dce {
namespace tc {

constexpr int kRW1 = 152;            // tape rows per window, T = 150 layers
constexpr int kRW2 = 76;             // tape rows per window, T = 75 layers
constexpr int kGuard = 8;            // leading guard rows of every tape (row r lives at index r + 8)
constexpr int kSlabRows = 130;       // 1 + 128 + 1
constexpr int kSlabBytes = kSlabRows * 16;
constexpr int64_t kChunk = 4096;     // windows per internal pass

enum Epi { EPI_TAPE = 0, EPI_POOL_TAPE = 1, EPI_POOL_FC = 2, EPI_FC_TAPE = 3, EPI_FC_F32 = 4, EPI_FC_LOGITS = 5 };

struct TapGemmParams {
    const uint8_t* a_tape;       // part 0 (hi); lo at + a_part_stride
    size_t a_part_stride;
    size_t a_kch_stride;         // bytes between consecutive kchunks = row capacity * 16
    const uint8_t* w_packed;     // [n_tile][stage][hi|lo][tap][j][BN][8] bf16
    const float* bias;           // [N]
    int m_tiles, n_tiles, stages;
    // epilogue
    uint8_t* out;                // tape outputs: part 0
    size_t out_part_stride, out_kch_stride;
    int out_rows_cap;            // rows the output tape can hold (excluding guards)
    float* out_f32;              // EPI_FC_F32: [n_valid][N];  EPI_FC_LOGITS: logit shares [2 * n_tiles][n_valid][16]
    const float* w3t;            // EPI_FC_LOGITS: the NEXT Linear layer's weight, k-major [N][16] (fc.6)
    int N;                       // total output features
    int rw, tv;                  // rows per window / valid rows per window of the INPUT tape (conv modes)
    int n_valid;                 // EPI_FC_F32: valid rows
    long long* trace;            // optional clock64 timeline of CTA 0: [tile][8] (tools/trace_tapgemm.py)
    int dbg;                     // timing ablations (results invalid): 1 = every tile loads the A slabs of tile 0; 2 = skip epilogue stores
    const float* acc_scale;      // F8 kernels: one float, 1 / (the layer's power-of-two weight scale), applied to the accumulator
    unsigned int* f8_status;     // F8 kernels: range diagnostic word (f8_range_note), or nullptr
};

struct Tape {
    int rows, m_tiles, cap;      // logical rows, 128-row tiles, row capacity incl. guards
    int kch;
    size_t kch_stride, part_stride, bytes;
};
inline Tape make_tape(int64_t rows, int kch) {
    Tape t; t.rows = (int)rows; t.m_tiles = (int)((rows + 127) / 128); t.m_tiles += t.m_tiles & 1;   // even: MT = 2 tiles
    t.cap = kGuard + t.m_tiles * 128 + 136; t.kch = kch;   // trailing guard: the fused kernels' 124-row tiles read up to 130 rows past a tile start
    t.kch_stride = (size_t)t.cap * 16; t.part_stride = t.kch_stride * kch; t.bytes = align_up(t.part_stride * 2, 256);
    return t;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split8(const float* y, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
        const float2 hf = __bfloat1622float2(hb);
        const __nv_bfloat162 lb = __floats2bfloat162_rn(y[2 * i] - hf.x, y[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hb);
        l[i] = *reinterpret_cast<const uint32_t*>(&lb);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// ---- fp16 main + e4m3 corrections (experimental FC operand format, option "fc_f16f8") ----------------
//     x * w ~= f16(x) * f16(w sw) + e4m3((x - f16(x)) 2^12) * e4m3(w sw 2^3) + e4m3(x 2^1) * e4m3((w sw - f16(w sw)) 2^14)
// sw = the layer's power-of-two weight scale (max |w sw| in [1, 2)).  Both correction products carry 2^15 and
// are accumulated first; the first fp16 MMA rescales the accumulator by 2^-15 (scale-input-d).  fp8 MMAs take
// K = 32 per instruction, so a K-step costs 2 + 2 MMA slots per 32 elements instead of 6
// (tools/emulate_split_precision.py: 6e-6 norm-wise on the logits; tools/microbench/umma_f16f8.cu).
constexpr int kF8ExL = 12, kF8EwH = 3, kF8ExH = 1, kF8EwL = 14, kF8ScaleD = 15;
static_assert(kF8ExL + kF8EwH == kF8ScaleD && kF8ExH + kF8EwL == kF8ScaleD, "both correction products carry the same scale");

// One layout rule for every tensor in this format: a row with C channels is C/8 fp16 chunks (8 channels, 16 bytes
// each), then C/16 lo8 chunks, then C/16 hi8 chunks (16 channels, 16 bytes each).  In a tape the fp16 chunks are
// part 0 and the e4m3 chunks part 1 (chunk pitch kch_stride); in a shared-memory slab all C/4 chunks follow each
// other (pitch kSlabBytes: C/8 fp16, C/16 lo8, C/16 hi8).  Writers and readers below go through these helpers, and
// tools/host_check_f16f8.cu tabulates them on the host for the CPU tests.
struct F8Dst { size_t f16, lo8, hi8; };      // byte offsets of fp16 chunk 2*g16 (chunk 2*g16+1 follows one pitch later), lo8 / hi8 chunk g16
__host__ __device__ __forceinline__ F8Dst f8_tape_dst(int g16, int C, size_t part_stride, size_t kch_stride) {
    F8Dst d;
    d.f16 = (size_t)(2 * g16) * kch_stride;
    d.lo8 = part_stride + (size_t)g16 * kch_stride;
    d.hi8 = part_stride + (size_t)(C / 16 + g16) * kch_stride;
    return d;
}
__host__ __device__ __forceinline__ F8Dst f8_slab_dst(int g16, int C) {
    F8Dst d;
    d.f16 = (size_t)(2 * g16) * (130 * 16);
    d.lo8 = (size_t)(C / 8 + g16) * (130 * 16);
    d.hi8 = (size_t)(C / 8 + C / 16 + g16) * (130 * 16);
    return d;
}
// tapgemm F8 producer: tape offset of activation copy (part, j) of stage s (K = 32 * stages, KSA = 4 chunks per part).
// Sweep 1 (s < stages/2) streams 16-element chunks s*4 + j of e4m3 image `part` (0 = lo8, 1 = hi8); sweep 2 streams
// fp16 chunks (s - stages/2)*8 + part*4 + j.
__host__ __device__ __forceinline__ size_t f8_stage_src(int s, int stages, int part, int j, size_t part_stride, size_t kch_stride) {
    const int half = stages >> 1;
    return (s < half) ? part_stride + (size_t)(part * 2 * stages + s * 4 + j) * kch_stride
                      : (size_t)((s - half) * 8 + part * 4 + j) * kch_stride;
}

// B-operand offsets inside one conv weight block of 192 * cout bytes (pack_conv_f16f8_kernel): an e4m3 block is
// [img: w8 | wl8][tap][2 chunks of 16 channels][cout][16 B], an fp16 block [tap][4 chunks of 8 channels][cout][16 B];
// the descriptor's LBO (distance between the two K chunks of one MMA) is 16 * cout in both.
__host__ __device__ __forceinline__ uint32_t f8_wblk_e4m3(int cout, int img, int tap) { return (uint32_t)((img * 3 + tap) * 32 * cout); }
__host__ __device__ __forceinline__ uint32_t f8_wblk_f16(int cout, int tap, int kk) { return (uint32_t)((tap * 2 + kk) * 32 * cout); }

// Range diagnostic of the format (dce_f16f8_status): bit `layer` of *status is set when a layer wrote an activation
// above 224 (the e4m3 image of 2 x saturates, so that element's second correction term is lost: fp16-only accuracy for
// it), bit 8 + layer when it wrote one above 65504 (the fp16 image itself saturated).  Layers: 0 conv1, 1 conv2,
// 2 conv3, 3 conv4, 4 fc.0.  Thread-local test, an atomic only in the abnormal case.
__device__ __forceinline__ void f8_range_note(const float* y, int n, unsigned int* status, int layer) {
    float mx = 0.f;
    for (int i = 0; i < n; ++i) mx = fmaxf(mx, y[i]);
    if (status && mx > 224.f) atomicOr(status, (1u << layer) | (mx > 65504.f ? (256u << layer) : 0u));
}

// The issue plans of the format, as host-callable functions: which operands MMA i of a stage / weight block reads, of
// which kind, and how it treats the accumulator.  The kernels issue exactly what these return, and
// tools/host_check_f16f8.cu dumps them so that the numpy emulation executes the SAME plan on real packed bytes.
struct F8Mma {
    uint32_t a_off, b_off;   // byte offsets: A relative to the M-tile's stage image (tapgemm) or the slab (fused kernels), B relative to the block
    uint8_t e4m3;            // 1: kind::f8f6f4, K = 32 (two 16-element chunks); 0: kind::f16, K = 16 (two 8-element chunks)
    uint8_t mode;            // 0: D = A*B;  1: D += A*B;  2: D = A*B + D * 2^-15 (scale-input-d: the corrections come down to the main scale)
};
// tapgemm, Linear layers: MMA i = 0..3 of stage s (stages/2 correction stages, then stages/2 main stages; 64 K-elements each).
// Stage image of an M-tile: [part 0: 4 slabs][part 1: 4 slabs]; B block: [part 0: 4 chunks][part 1: 4 chunks] of b_tapch bytes.
__host__ __device__ __forceinline__ F8Mma f8_fc_mma(int s, int stages, int i, uint32_t a_part, uint32_t b_part, uint32_t b_tapch) {
    F8Mma m;
    const int half = stages >> 1;
    if (s < half) {                     // i = 2 kk + product: product 0 = (x - f16 x) * w, 1 = x * (w - f16 w); centre row of the slab: + 16
        const int kk = i >> 1, prod = i & 1;
        m.a_off = 16 + prod * a_part + 2 * kk * (130 * 16);
        m.b_off = prod * b_part + 2 * kk * b_tapch;
        m.e4m3 = 1;
        m.mode = (s == 0 && i == 0) ? 0 : 1;
    } else {                            // 8 consecutive fp16 chunks of A (both parts) and of B
        m.a_off = 16 + 2 * i * (130 * 16);
        m.b_off = 2 * i * b_tapch;
        m.e4m3 = 0;
        m.mode = (s == half && i == 0) ? 2 : 1;
    }
    return m;
}
// fused conv kernels: MMA i = 0..1 of tap `tap` of weight block s (half e4m3 blocks, then half fp16 blocks, 32 input channels
// each); C = channels of the slab, cout = output channels (block = 192 * cout bytes); A offset includes the tap's row shift.
__host__ __device__ __forceinline__ F8Mma f8_conv_mma(int s, int half, int tap, int i, int C, int cout) {
    F8Mma m;
    if (s < half) {                     // i = product: 0 = (x - f16 x) * w (lo8 image, w8 image), 1 = x * (w - f16 w) (hi8, wl8)
        const F8Dst o = f8_slab_dst(2 * s, C);              // the slab's 16-channel groups 2s, 2s + 1
        m.a_off = (uint32_t)(i ? o.hi8 : o.lo8) + tap * 16;
        m.b_off = f8_wblk_e4m3(cout, i, tap);
        m.e4m3 = 1;
        m.mode = (s == 0 && tap == 0 && i == 0) ? 0 : 1;
    } else {                            // i = kk: fp16 chunks 4g + 2kk, 4g + 2kk + 1 of the slab, g = s - half
        m.a_off = (uint32_t)f8_slab_dst(2 * (s - half) + i, C).f16 + tap * 16;
        m.b_off = f8_wblk_f16(cout, tap, i);
        m.e4m3 = 0;
        m.mode = (s == half && tap == 0 && i == 0) ? 2 : 1;
    }
    return m;
}

// NaN-propagating min / max that also compile for the host, so tools/host_check_f16f8.cu can run the operand
// conversion and the weight packers below on the CPU and compare them byte for byte with the numpy emulation.
__host__ __device__ __forceinline__ float min_nan(float a, float b) {
#ifdef __CUDA_ARCH__
    float d;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
#else
    return (a != a || b != b) ? (a + b) : (a < b ? a : b);
#endif
}
__host__ __device__ __forceinline__ float max_nan_hd(float a, float b) {
#ifdef __CUDA_ARCH__
    float d;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
#else
    return (a != a || b != b) ? (a + b) : (a > b ? a : b);
#endif
}
// 16 consecutive K elements of one row -> two 16-byte fp16 chunks + one 16-byte chunk of each e4m3 image.
// Activations saturate at the largest finite fp16 (65504); NaN propagates through all three images.
// SIGNED = false: the values are ReLU outputs (>= 0 or NaN), only the upper clamp is needed.
template <bool SIGNED = false>
__host__ __device__ __forceinline__ void split16_f16f8(const float* y, uint4& f16a, uint4& f16b, uint4& lo8, uint4& hi8) {
    uint32_t h[8], l[4], g[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float a = min_nan(y[2 * i], 65504.f), b = min_nan(y[2 * i + 1], 65504.f);
        if (SIGNED) {
            a = max_nan_hd(a, -65504.f);
            b = max_nan_hd(b, -65504.f);
        }
        const __half2 hb = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(hb);
        h[i] = (uint32_t)__half_as_ushort(__low2half(hb)) | ((uint32_t)__half_as_ushort(__high2half(hb)) << 16);
        const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2((a - hf.x) * (float)(1 << kF8ExL), (b - hf.y) * (float)(1 << kF8ExL)), __NV_SATFINITE, __NV_E4M3);
        const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(a * (float)(1 << kF8ExH), b * (float)(1 << kF8ExH)), __NV_SATFINITE, __NV_E4M3);
        if (i & 1) { l[i >> 1] |= lo << 16; g[i >> 1] |= hi << 16; } else { l[i >> 1] = lo; g[i >> 1] = hi; }
    }
    f16a = make_uint4(h[0], h[1], h[2], h[3]);
    f16b = make_uint4(h[4], h[5], h[6], h[7]);
    lo8 = make_uint4(l[0], l[1], l[2], l[3]);
    hi8 = make_uint4(g[0], g[1], g[2], g[3]);
}

__host__ __device__ __forceinline__ void split16_f16f8_signed(const float* y, uint4& f16a, uint4& f16b, uint4& lo8, uint4& hi8) {
    split16_f16f8<true>(y, f16a, f16b, lo8, hi8);
}

// NaN-propagating max (FMNMX.NAN): torch's ReLU and MaxPool1d both propagate NaN.
__device__ __forceinline__ float max_nan(float a, float b) {
    float d;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float relu_nan(float v) { return max_nan(v, 0.f); }

constexpr int kEpiWarps = 8;                       // 2 per TMEM lane quadrant: column halves
constexpr int kMmaWarp = kEpiWarps;                // warps 8 .. 8+MT-1: one MMA issuer per accumulator.  Between two stages an
                                                   // issuer spends ~500 cycles in its own serial code (mbarrier probe, fences,
                                                   // descriptors) while the tensor pipe's short queue drains; with one issuer per
                                                   // accumulator the other issuer's MMAs fill that bubble, and every accumulator
                                                   // still sees its MMAs in one fixed order (bit-reproducible).
constexpr int kProdWarps = 4;                      // the bulk copies of a stage are dealt round-robin to 4 producer warps
__host__ __device__ constexpr int tapgemm_threads(int MT) { return (kEpiWarps + MT + kProdWarps) * 32; }

// MT = 128-row M-tiles per CTA tile: MT = 2 runs two accumulators against every staged B block, which
// halves the L2 -> smem weight traffic per MMA (the fc.0 / conv3 / conv4 / fc.3 tiles are L2-bound at MT = 1).
// WST > 0: the layer has WST k-stages and a single n-tile, and its whole weight image (WST * B_BYTES)
// stays resident in shared memory for the life of the CTA (conv3: 96 KB) — otherwise all 148 CTAs
// re-stream the same few weight blocks from the same L2 lines for every tile, which hot-spots L2.
template <int BN, int TAPS, int KSA, int NSTAGE, int MT = 1, int WST = 0>
struct TapGemmCfg {
    static constexpr int A_PART = KSA * kSlabBytes;
    static constexpr int A_TILE = 2 * A_PART;                 // hi + lo slabs of one M-tile
    static constexpr int A_BYTES = MT * A_TILE;
    static constexpr int B_TAPCH = BN * 16;
    static constexpr int B_PART = TAPS * KSA * B_TAPCH;
    static constexpr int B_BYTES = 2 * B_PART;
    static constexpr int STAGE_BYTES = A_BYTES + (WST ? 0 : B_BYTES);
    static constexpr int WRES_BYTES = WST * B_BYTES;
    static constexpr int NBUF = (2 * MT * BN <= 512) ? 2 : 1;   // accumulator buffers (epilogue / MMA overlap)
    static constexpr int TMEM_COLS = NBUF * MT * BN;
    static constexpr int BAR_BYTES = (2 * NSTAGE + 5) * 8 + 8;
    static constexpr int RING_BYTES = NSTAGE * STAGE_BYTES;
    static constexpr int SMEM_BYTES = RING_BYTES + WRES_BYTES + BAR_BYTES + kEpiWarps * (BN / 2) * 4;
    static_assert(KSA % 2 == 0, "an MMA consumes two kchunks");
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns: power of two");
    static_assert(STAGE_BYTES % 16 == 0, "bulk copies are 16-byte granular");
    static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
};

#ifndef DCE_TRACE
#define DCE_TRACE 0
#endif
#define TG_TRACE(k, ev) do { if (DCE_TRACE && p.trace && blockIdx.x == 0 && (k) < 60 && (threadIdx.x & 31) == 0) p.trace[(k) * 16 + (ev)] = clock64(); } while (0)

// F8 = 1 (Linear layers only): operands in the fp16 + e4m3 format (split16_f16f8).  A tile's K loop is two sweeps
// of p.stages / 2 stages each, 64 K-elements per stage and the same bytes per stage as the bf16x3 layout: sweep 1
// streams the e4m3 images (A: [lo8 | hi8], B: [w8 | wl8]) and issues the correction MMAs, sweep 2 streams the
// fp16 images and issues the main MMAs, the first of them with scale-input-d.
// CL = 2 (Linear layers, option "fc_cluster"): launched as clusters of two CTAs with ONE tile per CTA; tiles 2j and
// 2j + 1 share their M-tile(s) (n_tiles is even), so every activation slab of a stage is fetched from L2 by one of the
// two CTAs and multicast to both (L2 -> SM bytes per stage: A + B -> A / 2 + B), and a ring slot is free when the
// issuers of BOTH CTAs have drained it (multicast tcgen05.commit, `empty` count 2 MT).
template <int BN, int TAPS, int KSA, int NSTAGE, int EPI, int MT = 1, int WST = 0, int F8 = 0, int CL = 0>
__global__ void __launch_bounds__(tapgemm_threads(MT), 1)
tapgemm_kernel(const TapGemmParams p) {
    static_assert(!F8 || (TAPS == 1 && KSA == 4 && WST == 0 && (EPI == EPI_FC_TAPE || EPI == EPI_FC_LOGITS)), "F8: Linear layers, 64 K-elements per stage");
    static_assert(CL == 0 || (CL == 2 && TAPS == 1 && WST == 0), "clusters: pairs, Linear layers");
    constexpr int kProducerWarp0 = kEpiWarps + MT;
    using Cfg = TapGemmCfg<BN, TAPS, KSA, NSTAGE, MT, WST>;
    constexpr int NBUF = Cfg::NBUF;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* wres = smem + Cfg::RING_BYTES;                       // resident weight image (WST > 0)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES + Cfg::WRES_BYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* tfull = empty + NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* s_bias = reinterpret_cast<float*>(smem + Cfg::RING_BYTES + Cfg::WRES_BYTES + Cfg::BAR_BYTES);   // [8 warps][BN/2]
    float4* s_w3 = reinterpret_cast<float4*>(s_bias + kEpiWarps * (BN / 2));    // EPI_FC_LOGITS: [BN columns][16] fc.6 weights of this n-tile
    static_assert(EPI != EPI_FC_LOGITS || MT == 1, "the logit-share epilogue keeps one row per thread");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = (p.m_tiles / MT) * p.n_tiles;      // p.m_tiles is a multiple of MT (make_tape)

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { ptx::mbar_init(&full[i], kProdWarps); ptx::mbar_init(&empty[i], CL ? CL * MT : MT); }
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull[b], MT); ptx::mbar_init(&tempty[b], kEpiWarps); }
        ptx::mbar_init(wbar, 1);
        ptx::fence_barrier_init();
    }
    pdl_launch_dependents();                                    // the next kernel's prologue may overlap our tail
    if (warp == kMmaWarp) { ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS); ptx::tmem_relinquish(); }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if constexpr (CL != 0) ptx::cluster_sync();                 // the peer's barriers exist before anything remote touches them
    pdl_wait();                                                 // everything the previous kernel wrote is visible from here on

    if (warp >= kProducerWarp0) {
        // ===== TMA producers (warp-uniform loops; one elected lane per warp issues) =====
        // copy c of a stage: c < NA -> A slab (mt, part, j) of 2080 B; c == NA -> the B block.  Warp pw issues c = pw, pw+4, ...
        {
            constexpr int NA = MT * 2 * KSA;
            const int pw = warp - kProducerWarp0;
            uint32_t my_bytes = 0;
            for (int c = pw; c <= NA; c += kProdWarps) my_bytes += (c < NA) ? kSlabBytes : (WST ? 0 : Cfg::B_BYTES);
            if (WST && pw == 0) {                              // one-time load of the whole weight image
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(wbar, Cfg::WRES_BYTES);
#pragma unroll
                    for (int s = 0; s < (WST ? WST : 1); ++s)
                        ptx::bulk_g2s(wres + s * Cfg::B_BYTES, p.w_packed + (size_t)s * Cfg::B_BYTES, Cfg::B_BYTES, wbar);
                }
                __syncwarp();
            }
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m = (tile / p.n_tiles) * MT, n = tile % p.n_tiles;
                const uint8_t* a_row = p.a_tape + (size_t)(128 * ((p.dbg & 1) ? 0 : m) + kGuard - 1) * 16;
                const uint8_t* wsrc = p.w_packed + (size_t)n * p.stages * Cfg::B_BYTES;
                for (int s = 0; s < p.stages; ++s, ++it) {
                    const uint32_t slot = it % NSTAGE, ph = (it / NSTAGE) & 1;
                    ptx::mbar_wait_relaxed(&empty[slot], ph ^ 1);
                    if (pw == 0 && s == 0) TG_TRACE(it / p.stages, 0);
                    if (pw == 0 && s == p.stages - 1) TG_TRACE(it / p.stages, 1);
                    uint8_t* st = smem + slot * Cfg::STAGE_BYTES;
                    if ((p.dbg & 4) && it >= NSTAGE) {            // ablation: no TMA traffic after the ring is primed
                        if (ptx::elect_one()) ptx::mbar_arrive(&full[slot]);
                        __syncwarp();
                        continue;
                    }
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(&full[slot], my_bytes);
#pragma unroll
                        for (int c0 = 0; c0 <= NA; c0 += kProdWarps) {
                            const int c = c0 + pw;
                            if (c < NA) {
                                const int mt = c / (2 * KSA), part = (c / KSA) & 1, j = c % KSA;
                                if constexpr (CL != 0) {
                                    // slab c of the stage: fetched by CTA c % 2 of the pair, delivered to both
                                    size_t src_off = part * p.a_part_stride + (size_t)(s * KSA + j) * p.a_kch_stride;
                                    if (F8) src_off = f8_stage_src(s, p.stages, part, j, p.a_part_stride, p.a_kch_stride);
                                    if ((uint32_t)(c & 1) == ptx::cluster_ctarank())
                                        ptx::bulk_g2s_multicast(st + mt * Cfg::A_TILE + part * Cfg::A_PART + j * kSlabBytes,
                                                                a_row + (size_t)mt * 2048 + src_off, kSlabBytes, &full[slot], (uint16_t)0x3);
                                } else
                                if (F8) {
                                    const size_t src_off = f8_stage_src(s, p.stages, part, j, p.a_part_stride, p.a_kch_stride);
                                    ptx::bulk_g2s(st + mt * Cfg::A_TILE + part * Cfg::A_PART + j * kSlabBytes,
                                                  a_row + (size_t)mt * 2048 + src_off, kSlabBytes, &full[slot]);
                                } else {
                                    ptx::bulk_g2s(st + mt * Cfg::A_TILE + part * Cfg::A_PART + j * kSlabBytes,
                                                  a_row + (size_t)mt * 2048 + part * p.a_part_stride + (size_t)(s * KSA + j) * p.a_kch_stride,
                                                  kSlabBytes, &full[slot]);
                                }
                            } else if (c == NA && !WST) {
                                ptx::bulk_g2s(st + Cfg::A_BYTES, wsrc + (size_t)s * Cfg::B_BYTES, Cfg::B_BYTES, &full[slot]);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= kMmaWarp) {
        // ===== MMA issuers: warp kMmaWarp + mt owns accumulator mt =====
        // The tensor pipe queues only an MMA or two ahead of the issuing thread, so every cycle the issuer spends
        // between two stages (mbarrier probe ~90 cycles, fences, elect, descriptor arithmetic) is a pipe bubble
        // (tools/microbench/umma_mix.cu: 80 vs 65 cycles per N=128 MMA).  Hence: the leader lane is elected once,
        // and the probe of the NEXT stage's barrier (and of the next tile's accumulator) sits in the middle of the
        // current stage's MMAs, where the pipe still has queued work.
        {
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, BN);
            const int mt = warp - kMmaWarp;
            const bool leader = ptx::elect_one();
            const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const uint32_t total_stages = (uint32_t)my_tiles * p.stages;
            if (WST) ptx::mbar_wait(wbar, 0);
            uint32_t it = 0;
            if (my_tiles > 0) {
                ptx::mbar_wait(&tempty[0], 1);                       // fresh barrier: passes
                ptx::mbar_wait(&full[0], 0);
                ptx::tc_fence_after_sync();
            }
            for (int tcount = 0; tcount < my_tiles; ++tcount) {
                const uint32_t buf = tcount % NBUF;
                if (mt == 0) TG_TRACE(tcount, 2);
                const uint32_t d = tmem_base + buf * (MT * BN) + mt * BN;
                for (int s = 0; s < p.stages; ++s, ++it) {
                    const uint32_t slot = it % NSTAGE;
                    if (mt == 0 && s == 0) TG_TRACE(tcount, 4);
                    if (mt == 0 && s == p.stages - 1) TG_TRACE(tcount, 5);
                    const uint32_t a0 = ptx::smem_u32(smem + slot * Cfg::STAGE_BYTES) + mt * Cfg::A_TILE;
                    const uint32_t b0 = WST ? ptx::smem_u32(wres) + s * Cfg::B_BYTES : ptx::smem_u32(smem + slot * Cfg::STAGE_BYTES) + Cfg::A_BYTES;
                    if (F8) {
                        constexpr uint32_t id8 = ptx::make_idesc_e4m3_f32(128, BN), id16 = ptx::make_idesc_f16_f32(128, BN);
                        // With a single accumulator buffer the next tile's `tempty` (= this tile's epilogue) cannot
                        // complete before this tile's `tfull` commit at the end of the stage: wait for it at the top of
                        // the next tile instead of mid-stage (the bf16x3 loop below has the mid-stage wait and is kept to
                        // one tile per CTA by launch_layer; tools/simulate_block2_protocol.py).
                        if (NBUF == 1 && s == 0 && tcount > 0) { ptx::mbar_wait(&tempty[0], (tcount & 1) ^ 1); ptx::tc_fence_after_sync(); }
                        auto probe_next = [&]() {                          // as below: probe the next stage mid-stage
                            if (it + 1 < total_stages) {
                                if (NBUF > 1 && s == p.stages - 1) {
                                    const uint32_t nt = tcount + 1;
                                    ptx::mbar_wait(&tempty[nt % NBUF], ((nt / NBUF) & 1) ^ 1);
                                }
                                ptx::mbar_wait(&full[(it + 1) % NSTAGE], ((it + 1) / NSTAGE) & 1);
                                ptx::tc_fence_after_sync();
                            }
                        };
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // sweep 1: (x - f16 x) 2^12 * w 2^3 and x 2^1 * (w - f16 w) 2^14, K = 32 per MMA; sweep 2: the fp16 products
                            const F8Mma m = f8_fc_mma(s, p.stages, i, Cfg::A_PART, Cfg::B_PART, Cfg::B_TAPCH);
                            const uint64_t da = ptx::make_smem_desc(a0 + m.a_off, kSlabBytes, 128);
                            const uint64_t db = ptx::make_smem_desc(b0 + m.b_off, Cfg::B_TAPCH, 128);
                            if (leader) {
                                if (m.e4m3) ptx::umma_e4m3_ss(d, da, db, id8, m.mode);
                                else if (m.mode == 2) ptx::umma_f16_ss_scale_d<kF8ScaleD>(d, da, db, id16);
                                else ptx::umma_bf16_ss(d, da, db, id16, 1u);              // kind::f16; the operand format is in the idesc
                            }
                            if (i == 1) probe_next();
                        }
                    } else {
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
                        const int arow = (TAPS == 1) ? 1 : tap;            // Linear layers read the centre row only
#pragma unroll
                        for (int kk = 0; kk < KSA / 2; ++kk) {
                            const uint32_t b_hi = b0 + (tap * KSA + 2 * kk) * Cfg::B_TAPCH;
                            const uint64_t db_hi = ptx::make_smem_desc(b_hi, Cfg::B_TAPCH, 128);
                            const uint64_t db_lo = ptx::make_smem_desc(b_hi + Cfg::B_PART, Cfg::B_TAPCH, 128);
                            const uint32_t first = (s == 0 && tap == 0 && kk == 0) ? 0u : 1u;
                            const uint32_t a_hi = a0 + (2 * kk) * kSlabBytes + arow * 16;
                            const uint64_t da_hi = ptx::make_smem_desc(a_hi, kSlabBytes, 128);
                            const uint64_t da_lo = ptx::make_smem_desc(a_hi + Cfg::A_PART, kSlabBytes, 128);
                            if (leader) {
                                ptx::umma_bf16_ss(d, da_hi, db_lo, idesc, first);      // small terms first
                                ptx::umma_bf16_ss(d, da_lo, db_hi, idesc, 1u);
                                ptx::umma_bf16_ss(d, da_hi, db_hi, idesc, 1u);
                            }
                            // mid-stage: probe what the NEXT stage needs while MMAs of this one are still queued
                            if (tap == (TAPS - 1) / 2 && kk == (KSA / 2 - 1) / 2 && it + 1 < total_stages) {
                                if (s == p.stages - 1) {                       // next stage opens the next tile
                                    const uint32_t nt = tcount + 1;
                                    ptx::mbar_wait(&tempty[nt % NBUF], ((nt / NBUF) & 1) ^ 1);
                                }
                                ptx::mbar_wait(&full[(it + 1) % NSTAGE], ((it + 1) / NSTAGE) & 1);
                                ptx::tc_fence_after_sync();
                            }
                        }
                    }
                    }
                    if (leader) {
                        if constexpr (CL != 0) ptx::umma_commit_multicast(&empty[slot], (uint16_t)0x3);   // ... in both CTAs of the pair
                        else
                        ptx::umma_commit(&empty[slot]);          // this issuer's MMAs on the slot have retired
                        if (s == p.stages - 1) ptx::umma_commit(&tfull[buf]);   // this accumulator is complete
                    }
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias / ReLU / pool -> bf16 hi/lo -> next layer's tape =====
        // warp e: TMEM lane quadrant q = e % 4 (hardware restriction), column half h = e / 4.
        constexpr int HALF = BN / 2;
        const int q = warp & 3, h = warp >> 2;
        const int row_in_tile = q * 32 + lane;
        float* my_bias = s_bias + warp * HALF;               // warp-private copy of this warp's bias slice
        const float acc_scale = F8 ? __ldg(p.acc_scale) : 1.f;   // F8: the weights were packed times a power of two
        uint32_t tcount = 0;
        int last_n = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int m0 = (tile / p.n_tiles) * MT, n = tile % p.n_tiles;
            const uint32_t buf = tcount % NBUF, tph = (tcount / NBUF) & 1;
            const int n0 = n * BN + h * HALF;                // first output feature this warp owns
            if (n != last_n) {                               // stage this warp's bias slice (warp-private copy)
                __syncwarp();
                for (int i = lane; i < HALF; i += 32) my_bias[i] = __ldg(p.bias + n0 + i);
                if (EPI == EPI_FC_LOGITS) {
                    // fc.6 rows n*BN .. n*BN+BN-1, staged once by the eight epilogue warps together (they walk the same
                    // tile sequence, so they meet at this named barrier the same number of times)
                    asm volatile("bar.sync 2, 256;" ::: "memory");       // everyone is done with the previous n-tile's rows
                    const float4* src = reinterpret_cast<const float4*>(p.w3t) + (size_t)n * BN * 4;
                    for (int i = warp * 32 + lane; i < BN * 4; i += kEpiWarps * 32) s_w3[i] = __ldg(src + i);
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                }
                __syncwarp();
                last_n = n;
            }

            if (warp == 0) TG_TRACE(tcount, 6);
            ptx::mbar_wait_relaxed(&tfull[buf], tph);
            if (warp == 0) TG_TRACE(tcount, 7);
            ptx::tc_fence_after_sync();
            // per-M-tile row bookkeeping
            int rows[MT]; bool valid[MT]; size_t out_off[MT]; bool zero_prev[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int row = 128 * (m0 + mt) + row_in_tile;
                rows[mt] = row; valid[mt] = true; out_off[mt] = 0; zero_prev[mt] = false;
                if (EPI == EPI_TAPE) {
                    const int t = row % p.rw;
                    valid[mt] = t < p.tv;
                    out_off[mt] = (size_t)(row + kGuard) * 16;
                    zero_prev[mt] = (row == 0);
                } else if (EPI == EPI_POOL_TAPE) {
                    const int t = row % p.rw;
                    valid[mt] = (t >> 1) < (p.tv >> 1);
                    const int orow = row >> 1;
                    out_off[mt] = (size_t)(orow + kGuard) * 16;
                    zero_prev[mt] = (orow == 0);
                    if (orow >= p.out_rows_cap) out_off[mt] = (size_t)-1;
                } else if (EPI == EPI_POOL_FC) {
                    const int w = row / p.rw, t = row % p.rw;
                    const int to = t >> 1;
                    valid[mt] = to < (p.tv >> 1);
                    out_off[mt] = (valid[mt] && w < p.out_rows_cap) ? (size_t)(to * (BN / 8)) * p.out_kch_stride + (size_t)(w + kGuard) * 16 : (size_t)-1;
                } else if (EPI == EPI_FC_TAPE) {
                    out_off[mt] = (size_t)(row + kGuard) * 16;
                }
            }
            constexpr int CPM = HALF / 32;                   // 32-column chunks per M-tile for this warp
            constexpr int NCH = MT * CPM;
            const uint32_t taddr0 = tmem_base + buf * (MT * BN) + h * HALF + ((uint32_t)(q * 32) << 16);
            auto chunk_addr = [&](int ci) { return taddr0 + (ci / CPM) * BN + (ci % CPM) * 32; };

            float lg[16];                                    // EPI_FC_LOGITS: this thread's share of its row's 16 logits
#pragma unroll
            for (int o = 0; o < 16; ++o) lg[o] = 0.f;
            // one 32-column chunk: bias / ReLU / pool / guard -> bf16 hi/lo -> store
            auto process = [&](const uint32_t (&v)[32], int ci) {
                const int mt = ci / CPM, c0 = (ci % CPM) * 32;
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(my_bias + c0 + i);     // smem broadcast
                    if (F8) {
                        y[i] = relu_nan(fmaf(__uint_as_float(v[i]), acc_scale, b4.x));
                        y[i + 1] = relu_nan(fmaf(__uint_as_float(v[i + 1]), acc_scale, b4.y));
                        y[i + 2] = relu_nan(fmaf(__uint_as_float(v[i + 2]), acc_scale, b4.z));
                        y[i + 3] = relu_nan(fmaf(__uint_as_float(v[i + 3]), acc_scale, b4.w));
                        continue;
                    }
                    y[i] = relu_nan(__uint_as_float(v[i]) + b4.x);
                    y[i + 1] = relu_nan(__uint_as_float(v[i + 1]) + b4.y);
                    y[i + 2] = relu_nan(__uint_as_float(v[i + 2]) + b4.z);
                    y[i + 3] = relu_nan(__uint_as_float(v[i + 3]) + b4.w);
                }
                if (EPI == EPI_FC_LOGITS) {
                    // fc.6 folded into fc.3's epilogue: H2 never leaves the registers (smem reads are warp broadcasts)
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float4* wr = s_w3 + (h * HALF + c0 + i) * 4;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 w = wr[j];
                            lg[4 * j + 0] = fmaf(y[i], w.x, lg[4 * j + 0]);
                            lg[4 * j + 1] = fmaf(y[i], w.y, lg[4 * j + 1]);
                            lg[4 * j + 2] = fmaf(y[i], w.z, lg[4 * j + 2]);
                            lg[4 * j + 3] = fmaf(y[i], w.w, lg[4 * j + 3]);
                        }
                    }
                    return;
                }
                if (EPI == EPI_FC_F32) {
                    if (rows[mt] < p.n_valid) {
                        float* dst = p.out_f32 + (size_t)rows[mt] * p.N + n0 + c0;
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4*>(dst + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
                    }
                    return;
                }
                if (F8 && EPI == EPI_FC_TAPE) {
                    f8_range_note(y, 32, p.f8_status, 4);
                    // next layer's operand in the fp16 + e4m3 format: fp16 chunks in tape part 0, the two e4m3 images
                    // (N / 16 chunks each: lo8, then hi8) in tape part 1
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint4 fa, fb, lo8, hi8;
                        split16_f16f8(y + hh * 16, fa, fb, lo8, hi8);
                        const F8Dst d = f8_tape_dst((n0 + c0) / 16 + hh, p.N, p.out_part_stride, p.out_kch_stride);
                        uint8_t* row = p.out + out_off[mt];
                        *reinterpret_cast<uint4*>(row + d.f16) = fa;
                        *reinterpret_cast<uint4*>(row + d.f16 + p.out_kch_stride) = fb;
                        *reinterpret_cast<uint4*>(row + d.lo8) = lo8;
                        *reinterpret_cast<uint4*>(row + d.hi8) = hi8;
                    }
                    return;
                }
                if (EPI == EPI_POOL_TAPE || EPI == EPI_POOL_FC) {
                    // MaxPool1d(2,2): rows (2i, 2i+1) are adjacent lanes (src/contact_cnn.py:24-25,42-43)
#pragma unroll
                    for (int i = 0; i < 32; ++i) y[i] = max_nan(y[i], __shfl_xor_sync(0xffffffffu, y[i], 1));
                }
                if (EPI == EPI_TAPE || EPI == EPI_POOL_TAPE) {
                    if (!valid[mt]) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) y[i] = 0.f;                  // guard rows stay zero
                    }
                }
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 hi, lo;
                    split8(y + qd * 8, hi, lo);
                    const int kch = (n0 + c0) / 8 + qd;
                    if (EPI == EPI_TAPE || EPI == EPI_FC_TAPE) {
                        uint8_t* dst = p.out + (size_t)kch * p.out_kch_stride + out_off[mt];
                        *reinterpret_cast<uint4*>(dst) = hi;
                        *reinterpret_cast<uint4*>(dst + p.out_part_stride) = lo;
                        if (EPI == EPI_TAPE && zero_prev[mt]) {
                            *reinterpret_cast<uint4*>(dst - 16) = make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(dst + p.out_part_stride - 16) = make_uint4(0, 0, 0, 0);
                        }
                    } else {
                        // pooled: even lane writes the hi part, odd lane the lo part of the pooled row
                        if (out_off[mt] != (size_t)-1) {
                            uint8_t* dst = p.out + (size_t)kch * p.out_kch_stride + out_off[mt] + ((lane & 1) ? p.out_part_stride : 0);
                            *reinterpret_cast<uint4*>(dst) = (lane & 1) ? lo : hi;
                            if (EPI == EPI_POOL_TAPE && zero_prev[mt]) *reinterpret_cast<uint4*>(dst - 16) = make_uint4(0, 0, 0, 0);
                        }
                    }
                }
            };

            // software pipeline over the chunks: the TMEM load of chunk i+1 is in flight while chunk i is processed
            if (!(p.dbg & 2)) {
                uint32_t va[32], vb[32];
                ptx::tmem_ld32(chunk_addr(0), va);
#pragma unroll
                for (int ci = 0; ci < NCH; ci += 2) {
                    ptx::tmem_ld_wait();
                    if (ci + 1 < NCH) ptx::tmem_ld32(chunk_addr(ci + 1), vb);
                    process(va, ci);
                    if (ci + 1 < NCH) {
                        ptx::tmem_ld_wait();
                        if (ci + 2 < NCH) ptx::tmem_ld32(chunk_addr(ci + 2), va);
                        process(vb, ci + 1);
                    }
                }
            }
            if (EPI == EPI_FC_LOGITS && !(p.dbg & 2) && rows[0] < p.n_valid) {
                float4* dst = reinterpret_cast<float4*>(p.out_f32 + ((size_t)(n * 2 + h) * p.n_valid + rows[0]) * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(lg[4 * j], lg[4 * j + 1], lg[4 * j + 2], lg[4 * j + 3]);
            }
            if (warp == 0) TG_TRACE(tcount, 8);
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[buf]);
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if constexpr (CL != 0) ptx::cluster_sync();                 // the peer may still multicast into this CTA or signal its barriers
    if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// K0 (tensor-core part): weights -> bf16 hi/lo shared-memory images, one block per (n_tile, stage)
// kind 0: conv  W[n][cin][3]      K index (tap, c), c padded to 8*KSA*stages
// kind 1: fc.0  W[n][c*37 + t]    K index k' = t*128 + c      (src/contact_cnn.py:64)
// kind 2: fc.3  W[n][k]
// ---------------------------------------------------------------------------------------------
__global__ void pack_b_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int n_tiles, int stages,
                              int BN, int TAPS, int KSA, int kind, int cin) {
    const size_t per_part = (size_t)TAPS * KSA * BN * 8;            // elements in one part of one block
    const size_t total = (size_t)n_tiles * stages * per_part;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx;
        const int e = r % 8; r /= 8;
        const int nn = r % BN; r /= BN;
        const int j = r % KSA; r /= KSA;
        const int tap = r % TAPS; r /= TAPS;
        const int s = r % stages; r /= stages;
        const int nt = (int)r;
        const int n = nt * BN + nn;
        const int c = (s * KSA + j) * 8 + e;
        float v;
        if (kind == 0) v = (c < cin) ? W[((size_t)n * cin + c) * 3 + tap] : 0.f;
        else if (kind == 1) { const int t = c / 128, ch = c % 128; v = W[(size_t)n * 4736 + ch * 37 + t]; }
        else v = W[(size_t)n * cin + c];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        const size_t blk = ((size_t)nt * stages + s) * (2 * per_part * 2);      // bytes
        const size_t within = ((((size_t)tap * KSA + j) * BN + nn) * 8 + e) * 2;
        *reinterpret_cast<__nv_bfloat16*>(out + blk + within) = h;
        *reinterpret_cast<__nv_bfloat16*>(out + blk + per_part * 2 + within) = l;
    }
}

// ---------------------------------------------------------------------------------------------
// K0, fp16 + e4m3 format (option "fc_f16f8"; Linear layers): per n-tile, `stages` blocks of 8 * BN * 16 bytes in
// the order the F8 kernel streams them: stages/2 correction blocks [w8 | wl8] (4 chunks of 16 e4m3 each) covering
// 64 K-elements apiece, then stages/2 main blocks (8 chunks of 8 fp16).  Weights are multiplied by the layer's
// power-of-two scale sw first (max |w sw| in [1, 2)), so small weights stay clear of the fp16 subnormals.
// kind 3: fc.0 (K index k' = t*128 + c), kind 4: fc.3.
// ---------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ W, size_t n, unsigned int* __restrict__ out_bits) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float a = fabsf(W[i]);
        if (a <= 3.0e38f) m = fmaxf(m, a);                     // NaN / inf weights do not set the scale
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));      // non-negative floats order like their bits
}
// scale[0] = sw = 2^-floor(log2 max|w|), scale[1] = 1 / sw; scale[2] holds the absmax bits on entry
__global__ void weight_scale_kernel(float* __restrict__ scale) {
    const float m = __uint_as_float(reinterpret_cast<const unsigned int*>(scale)[2]);
    int e = 1;
    if (m > 0.f) frexpf(m, &e);                                // m = f * 2^e, f in [0.5, 1)  ->  floor(log2 m) = e - 1
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    scale[0] = ldexpf(1.f, 1 - e);
    scale[1] = ldexpf(1.f, e - 1);
}
// one weight element (idx = n * K + k) of a Linear layer -> its three images
__host__ __device__ __forceinline__ void pack_b_f16f8_elem(const float* __restrict__ W, uint8_t* __restrict__ out, size_t idx,
                                                           int stages, int BN, int kind, int K, float sw) {
    const size_t blk_bytes = (size_t)8 * BN * 16;
    const int half = stages / 2;                               // K == 64 * half
    {
        const int k = (int)(idx % K);
        const int n = (int)(idx / K);
        const int nt = n / BN, nn = n % BN;
        float v;
        if (kind == 3) { const int t = k / 128, ch = k % 128; v = W[(size_t)n * 4736 + ch * 37 + t]; }
        else v = W[(size_t)n * K + k];
        v *= sw;
        const __half h = __float2half_rn(v);
        const float r = v - __half2float(h);
        const int s = k / 64, kr = k % 64;
        uint8_t* corr = out + ((size_t)nt * stages + s) * blk_bytes;
        uint8_t* mainb = out + ((size_t)nt * stages + half + s) * blk_bytes;
        *reinterpret_cast<__half*>(mainb + (((size_t)(kr / 8) * BN + nn) * 8 + kr % 8) * 2) = h;
        const size_t o8 = ((size_t)(kr / 16) * BN + nn) * 16 + kr % 16;
        corr[o8] = (uint8_t)__nv_cvt_float_to_fp8(v * (float)(1 << kF8EwH), __NV_SATFINITE, __NV_E4M3);
        corr[blk_bytes / 2 + o8] = (uint8_t)__nv_cvt_float_to_fp8(r * (float)(1 << kF8EwL), __NV_SATFINITE, __NV_E4M3);
    }
}
__global__ void pack_b_f16f8_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int n_tiles, int stages,
                                    int BN, int kind, int K, const float* __restrict__ scale) {
    const float sw = scale[0];
    const size_t total = (size_t)n_tiles * BN * K;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
        pack_b_f16f8_elem(W, out, idx, stages, BN, kind, K, sw);
}

// Conv layers in the fp16 + e4m3 format (option "conv_f16f8"): blocks of 192 * cout bytes, each covering 32 input
// channels of all three taps; first the G = cin_pad / 32 e4m3 blocks [img: w8 | wl8][tap][2 chunks of 16][cout][16 B],
// then the G fp16 blocks [tap][4 chunks of 8][cout][8 x 2 B] — the order the two-sweep issuers consume them
// (block2's 24 KB weight ring at cout = 128; block1 keeps its four 12 KB blocks resident).
__host__ __device__ __forceinline__ void pack_conv_f16f8_elem(const float* __restrict__ W, uint8_t* __restrict__ out, int idx,
                                                              int cout, int cin, int cin_pad, float sw) {
    const int G = cin_pad / 32;
    const size_t blk = (size_t)192 * cout;
    {
        const int tap = idx % 3, c = (idx / 3) % cin_pad, n = idx / (3 * cin_pad);
        const float v = (c < cin ? W[((size_t)n * cin + c) * 3 + tap] : 0.f) * sw;
        const __half h = __float2half_rn(v);
        const float r = v - __half2float(h);
        const int g = c / 32, cr = c % 32;
        uint8_t* b8 = out + (size_t)g * blk;
        uint8_t* b16 = out + (size_t)(G + g) * blk;
        *reinterpret_cast<__half*>(b16 + ((size_t)(tap * 4 + cr / 8) * cout + n) * 16 + (cr % 8) * 2) = h;
        const size_t o8 = ((size_t)(tap * 2 + cr / 16) * cout + n) * 16 + cr % 16;
        b8[o8] = (uint8_t)__nv_cvt_float_to_fp8(v * (float)(1 << kF8EwH), __NV_SATFINITE, __NV_E4M3);
        b8[blk / 2 + o8] = (uint8_t)__nv_cvt_float_to_fp8(r * (float)(1 << kF8EwL), __NV_SATFINITE, __NV_E4M3);
    }
}
__global__ void pack_conv_f16f8_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int cout, int cin, int cin_pad,
                                       const float* __restrict__ scale) {
    const float sw = scale[0];
    const int total = cout * cin_pad * 3;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
        pack_conv_f16f8_elem(W, out, idx, cout, cin, cin_pad, sw);
}

// ---------------------------------------------------------------------------------------------
// ingest: fp32 windows -> layer-0 tape (54 channels padded to 64, bf16 hi/lo, guard rows zero)
//   batch mode : x = [B][150][54]                                 (DataLoader batch, src/inference_one_seq.py:24)
//   stream mode: x = [T][54] log; window w = rows first+w .. +149, z-scored with stats[w]
//                (utils/data_handler.py:55-56)
// block = 8 warps = 8 kchunks x 32 consecutive tape rows
// ---------------------------------------------------------------------------------------------
template <bool STREAM>
__global__ void __launch_bounds__(256)
ingest_kernel(const float* __restrict__ x, int64_t first, int n_windows, const float* __restrict__ mean,
              const float* __restrict__ sdev, uint8_t* __restrict__ tape, size_t part_stride, size_t kch_stride) {
    const int kch = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 32 + lane;
    const int w = row / kRW1, t = row % kRW1;
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = 0.f;
    if (w < n_windows && t < 150 && kch < 7) {
        const float* src = STREAM ? x + (size_t)(first + w + t) * 54 + kch * 8 : x + ((size_t)w * 150 + t) * 54 + kch * 8;
        const int nv = (kch == 6) ? 6 : 8;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            if (i < nv) {
                const float2 f = __ldg(reinterpret_cast<const float2*>(src + i));
                y[i] = f.x; y[i + 1] = f.y;
            }
        }
        if (STREAM) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) y[i] = (y[i] - __ldg(mean + (size_t)w * 64 + kch * 8 + i)) / __ldg(sdev + (size_t)w * 64 + kch * 8 + i);
        }
    }
    uint4 hi, lo;
    split8(y, hi, lo);
    uint8_t* dst = tape + (size_t)kch * kch_stride + (size_t)(row + kGuard) * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + part_stride) = lo;
    if (row == 0) {
        *reinterpret_cast<uint4*>(dst - 16) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dst + part_stride - 16) = make_uint4(0, 0, 0, 0);
    }
}

// per-window, per-channel mean and unbiased std in fp32 (utils/data_handler.py:55-56).
// A CTA owns 32 consecutive windows = 181 log rows, read ONCE into shared memory as y = x - pivot (the tile's
// middle row, per channel: removes the channel offset, so the sums below carry no cancellation from it), then
// per-channel prefix sums of y and y^2 along time give every window's two sums as a difference of two prefixes:
//     mean = pivot + sum/150,   var = (sumsq - sum * sum/150) / 149
// instead of 300 strided loads per (window, channel) (the first version of this kernel was LSU-bound, 21 us per
// 4096 windows).  A tile holding a non-finite value (a NaN would poison every later prefix) takes the direct
// two-pass route, window by window, so NaN / inf stay confined to the windows that contain them.
constexpr int kStatWT = 32;
constexpr int kStatRows = kStatWT + 149;
constexpr int kStatSmemBytes = 2 * kStatRows * 54 * 4;

__device__ __forceinline__ void window_stats_direct(const float* __restrict__ src, float& mu, float& sd) {
    float s = 0.f;
    for (int t = 0; t < 150; ++t) s += __ldg(src + (size_t)t * 54);
    mu = s / 150.f;
    float v = 0.f;
    for (int t = 0; t < 150; ++t) { const float d = __ldg(src + (size_t)t * 54) - mu; v = fmaf(d, d, v); }
    sd = sqrtf(v / 149.f);
}

// Tiles are aligned to ABSOLUTE window indices (tile j = windows 32j .. 32j+31 of the log, pivot = log row 32j+90,
// scan from log row 32j) so a window's statistics — and with them its logits — do not depend on which sub-range
// of the log a call asks for: dce_stream(first, n) equals the matching slice of a whole-log call bit for bit.
__global__ void __launch_bounds__(256)
window_stats_kernel(const float* __restrict__ x, int64_t first, int n_windows, int64_t total_rows,
                    float* __restrict__ mean, float* __restrict__ sdev, int reciprocal) {
    extern __shared__ __align__(16) float st_smem[];
    float* S = st_smem;
    float* Q = st_smem + kStatRows * 54;
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row0 = (first / kStatWT + blockIdx.x) * kStatWT;            // first log row (= first window) of this tile
    const int R = (total_rows - row0 < kStatRows) ? (int)(total_rows - row0) : kStatRows;
    const int wlo = (first > row0) ? (int)(first - row0) : 0;                   // windows [wlo, whi) of the tile are wanted
    const int whi = (first + n_windows - row0 < kStatWT) ? (int)(first + n_windows - row0) : kStatWT;
    const float* src = x + (size_t)row0 * 54;
    constexpr int rp = 90;                                                      // pivot row: a wanted window exists, so R >= 150
    int bad = 0;
    for (int i = threadIdx.x; i < R * 54; i += 256) {
        const float y = __ldg(src + i) - __ldg(src + rp * 54 + i % 54);
        bad |= !(fabsf(y) <= 3.0e38f);                 // NaN or inf (also a finite overflow of the difference)
        S[i] = y; Q[i] = y * y;
    }
    bad = __syncthreads_or(bad);
    if (bad) {
        for (int i = wlo * 54 + threadIdx.x; i < whi * 54; i += 256) {
            const int w = i / 54, c = i % 54;
            float mu, sd;
            window_stats_direct(src + (size_t)w * 54 + c, mu, sd);
            const size_t o = (size_t)(row0 + w - first) * 64 + c;
            mean[o] = mu;
            sdev[o] = reciprocal ? 1.f / sd : sd;
        }
        return;
    }
    // inclusive scan along time, per channel: 4 row segments per channel, then the segment offsets
    const int c = threadIdx.x % 54, seg = threadIdx.x / 54;          // threads 216..255 idle here
    const int seg_rows = (R + 3) / 4;
    const int r0 = seg * seg_rows, r1 = (r0 + seg_rows < R) ? r0 + seg_rows : R;
    if (seg < 4) {
        float s = 0.f, q = 0.f;
        for (int r = r0; r < r1; ++r) { s += S[r * 54 + c]; S[r * 54 + c] = s; q += Q[r * 54 + c]; Q[r * 54 + c] = q; }
    }
    __syncthreads();
    float so = 0.f, qo = 0.f;
    if (seg < 4)
        for (int j = 0; j < seg; ++j) {
            const int e = (((j + 1) * seg_rows < R) ? (j + 1) * seg_rows : R) - 1;
            so += S[e * 54 + c]; qo += Q[e * 54 + c];
        }
    __syncthreads();
    if (seg > 0 && seg < 4)
        for (int r = r0; r < r1; ++r) { S[r * 54 + c] += so; Q[r * 54 + c] += qo; }
    __syncthreads();
    for (int i = wlo * 54 + threadIdx.x; i < whi * 54; i += 256) {
        const int w = i / 54, cc = i % 54;
        float s = S[(w + 149) * 54 + cc], q = Q[(w + 149) * 54 + cc];
        if (w > 0) { s -= S[(w - 1) * 54 + cc]; q -= Q[(w - 1) * 54 + cc]; }
        const float my = s / 150.f;
        const float ss = q - s * my;                   // sum of squared deviations
        float mu = __ldg(src + rp * 54 + cc) + my;
        float sd = sqrtf(((ss < 0.f) ? 0.f : ss) / 149.f);
        // the window sits far from the pivot compared with its spread (or is constant): the difference above has
        // lost >10 bits, so take the direct two-pass route for this (window, channel) — rare, and it keeps a
        // constant channel exactly at std = 0 -> NaN, as in the reference
        if (ss < 1e-3f * q) window_stats_direct(src + (size_t)w * 54 + cc, mu, sd);
        const size_t o = (size_t)(row0 + w - first) * 64 + cc;
        mean[o] = mu;
        sdev[o] = reciprocal ? 1.f / sd : sd;          // std == 0 -> inf -> (x - mean) * inf = NaN, as 0/0 in the reference
    }
}

// ---------------------------------------------------------------------------------------------
// host side: packed layout, workspace, launch sequence
// ---------------------------------------------------------------------------------------------
struct LayerCfg { int BN, TAPS, KSA, stages, n_tiles, kind, cin, src; };
//                                   BN  TAPS KSA stages n_tiles kind cin   state_dict index of the weight
constexpr int kNumPacked = 14;
constexpr LayerCfg kLayers[kNumPacked] = {{64, 3, 4, 2, 1, 0, 54, 0},       // block1.0
                                 {64, 3, 4, 2, 1, 0, 64, 2},       // block1.2
                                 {128, 3, 4, 2, 1, 0, 64, 4},      // block2.0
                                 {128, 3, 2, 8, 1, 0, 128, 6},     // block2.2
                                 {256, 1, 4, 148, 8, 1, 4736, 8},  // fc.0
                                 {128, 1, 4, 64, 4, 2, 2048, 10},  // fc.3
                                 {128, 3, 4, 4, 1, 0, 128, 6},     // block2.2 again, in 48 KB blocks (unused; kept for ablations)
                                 {128, 3, 2, 4, 1, 0, 64, 4},      // block2.0 again, in 24 KB blocks for the fused block2 kernel's weight ring
                                 {256, 1, 4, 148, 8, 3, 4736, 8},  // fc.0 in the fp16 + e4m3 format (option "fc_f16f8")
                                 {128, 1, 4, 64, 4, 4, 2048, 10},  // fc.3 in the fp16 + e4m3 format
                                 {64, 3, 2, 4, 1, 5, 54, 0},       // block1.0 in the fp16 + e4m3 format (option "conv_f16f8"): 2 * cin_pad/32 blocks
                                 {64, 3, 2, 4, 1, 5, 64, 2},       // block1.2
                                 {128, 3, 2, 4, 1, 5, 64, 4},      // block2.0
                                 {128, 3, 2, 8, 1, 5, 128, 6}};    // block2.2
inline size_t layer_packed_bytes(const LayerCfg& c) { return (size_t)c.n_tiles * c.stages * 2 * c.TAPS * c.KSA * c.BN * 16; }

struct PackedLayout { size_t w[kNumPacked]; size_t scales; size_t begin, end; };   // scales: [kNumPacked][4] floats {sw, 1/sw, absmax bits, -}; then the
constexpr int kF8StatusWord = 60;                                                  // f8 range-status word: 32-bit word 60 of that 256-byte block
inline PackedLayout make_packed_layout(size_t base) {
    PackedLayout L; L.begin = base; size_t o = base;
    for (int i = 0; i < kNumPacked; ++i) { L.w[i] = o; o = align_up(o + layer_packed_bytes(kLayers[i]), 256); }
    L.scales = o; o = align_up(o + kNumPacked * 16, 256);
    L.end = o;
    return L;
}

inline int pack(char* buf, const PackedLayout& L, const float* const* params, Ctx& ctx) {
    for (int i = 0; i < kNumPacked; ++i) {
        const LayerCfg& c = kLayers[i];
        const size_t total = (size_t)c.n_tiles * c.stages * c.TAPS * c.KSA * c.BN * 8;
        const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        if (c.kind >= 3) {                                     // fp16 + e4m3 format: scale from max |w|, then the images
            float* sc = reinterpret_cast<float*>(buf + L.scales) + i * 4;
            const size_t nw = (size_t)c.n_tiles * c.BN * c.cin * c.TAPS;
            cudaError_t e = cudaMemsetAsync(sc, 0, 16, ctx.stream);
            if (e != cudaSuccess) { ctx.err = e; return DCE_ECUDA; }
            DCE_KL(ctx, "tc_absmax", absmax_kernel<<<1024, 256, 0, ctx.stream>>>(params[c.src], nw, reinterpret_cast<unsigned int*>(sc) + 2));
            DCE_KL(ctx, "tc_weight_scale", weight_scale_kernel<<<1, 1, 0, ctx.stream>>>(sc));
            if (c.kind == 5) {                                 // conv: cin_pad = 16 * stages (2 blocks per 32 channels)
                DCE_KL(ctx, "tc_pack_conv_f16f8", pack_conv_f16f8_kernel<<<96, 256, 0, ctx.stream>>>(
                    params[c.src], reinterpret_cast<uint8_t*>(buf + L.w[i]), c.BN, c.cin, 16 * c.stages, sc));
                continue;
            }
            DCE_KL(ctx, "tc_pack_b_f16f8", pack_b_f16f8_kernel<<<4096, 256, 0, ctx.stream>>>(
                params[c.src], reinterpret_cast<uint8_t*>(buf + L.w[i]), c.n_tiles, c.stages, c.BN, c.kind, c.cin, sc));
            continue;
        }
        DCE_KL(ctx, "tc_pack_b", pack_b_kernel<<<blocks, 256, 0, ctx.stream>>>(
            params[c.src], reinterpret_cast<uint8_t*>(buf + L.w[i]), c.n_tiles, c.stages, c.BN, c.TAPS, c.KSA, c.kind, c.cin));
    }
    return DCE_OK;
}

struct Workspace {
    Tape x0, x1, x2, x3, x4, h1;
    size_t o_x0, o_x1, o_x2, o_x3, o_x4, o_h1, o_h2, o_mean, o_sdev, end;
};
inline Workspace make_workspace(int64_t n) {
    Workspace W; size_t o = 256;          // [0,256): the latency kernel's barrier counters (dce_latency.cuh)
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    W.x0 = make_tape(n * kRW1, 8);   W.o_x0 = take(W.x0.bytes);
    W.x1 = make_tape(n * kRW1, 8);   W.o_x1 = take(W.x1.bytes);
    W.x2 = make_tape(n * kRW2, 8);   W.o_x2 = take(W.x2.bytes);
    W.x3 = make_tape(n * kRW2, 16);  W.o_x3 = take(W.x3.bytes);
    W.x4 = make_tape(n, 592);        W.o_x4 = take(W.x4.bytes);
    W.h1 = make_tape(n, 256);        W.o_h1 = take(W.h1.bytes);
    W.o_h2 = take((size_t)n * 512 * 4);
    W.o_mean = take((size_t)n * 64 * 4);
    W.o_sdev = take((size_t)n * 64 * 4);
    W.end = o;
    return W;
}
inline size_t workspace_bytes(int64_t max_windows) {
    return make_workspace(max_windows < kChunk ? max_windows : kChunk).end;
}

template <int BN, int TAPS, int KSA, int NSTAGE, int EPI, int MT = 1, int WST = 0, int F8 = 0, int CL = 0>
inline int launch_layer(Ctx& ctx, const char* name, int sm_count, const TapGemmParams& p) {
    using Cfg = TapGemmCfg<BN, TAPS, KSA, NSTAGE, MT, WST>;
    auto kern = tapgemm_kernel<BN, TAPS, KSA, NSTAGE, EPI, MT, WST, F8, CL>;
    if (WST && (p.stages != WST || p.n_tiles != 1)) return DCE_EINVAL;
    if (F8 && ((p.stages & 1) || !p.acc_scale)) return DCE_EINVAL;
    constexpr int kSmem = Cfg::SMEM_BYTES + (EPI == EPI_FC_LOGITS ? BN * 64 : 0);      // + [BN][16] fp32 of the next layer
    static_assert(kSmem <= 232448, "exceeds 227 KB");
    static DeviceOnce attr_once;
    if (auto first_ = attr_once.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) { ctx.err = e; return DCE_ECUDA; }
    }
    const int tiles = (p.m_tiles / MT) * p.n_tiles;
    const int grid = tiles < sm_count ? tiles : sm_count;
    // With a single accumulator buffer (fc.0: 2 x 256 columns fill the TMEM) the issuer's mid-stage probe of the next
    // tile's `tempty` would wait for an epilogue that cannot start before the current tile's `tfull` commit: a CTA must
    // not get a second tile (tools/simulate_block2_protocol.py).  128 tiles on a B200's 148 SMs: never the case here.
    if (Cfg::NBUF == 1 && tiles > grid && !F8) return DCE_EUNSUPPORTED;
    if constexpr (CL != 0) {
        // pairs need one tile per CTA, tiles 2j / 2j+1 on the same M-tile, and every pair resident at once
        if (tiles > sm_count || (p.n_tiles & 1)) return DCE_EUNSUPPORTED;
        static int max_clusters[64] = {};
        static DeviceOnce cl_once;
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        if (auto first_ = cl_once.need()) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(sm_count & ~1)); cfg.blockDim = dim3(tapgemm_threads(MT)); cfg.dynamicSmemBytes = kSmem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
            if (e != cudaSuccess) { ctx.err = e; return DCE_ECUDA; }
            max_clusters[dev] = n > 0 ? n : -1;
        }
        if (max_clusters[dev] * 2 < tiles) return DCE_EUNSUPPORTED;
        DCE_KL(ctx, name, { cudaError_t le_ = launch_pdl_cluster(kern, dim3(tiles), dim3(tapgemm_threads(MT)), kSmem, ctx.stream, 2, p); (void)le_; });
    } else {
        DCE_KL(ctx, name, { cudaError_t le_ = launch_pdl(kern, dim3(grid), dim3(tapgemm_threads(MT)), kSmem, ctx.stream, p); (void)le_; });
    }
    return DCE_OK;
}

}  // namespace tc
}  // namespace dce
