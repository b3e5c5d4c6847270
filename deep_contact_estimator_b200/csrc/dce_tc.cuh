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
    CUtensorMap a_map;           // AT = 1 (fc.0): the A operand as a row-major tensor {K elements, rows, hi | lo}; box {32, 128, 1}, 64-byte swizzle
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
    // EPI_FC_LOGITS with tickets != nullptr: the CTA that completes an M-tile's last n-tile (atomic ticket per M-tile) adds the
    // logit shares in their fixed order + bias, takes the argmax and writes logits / class / contact bits: no extra launch
    unsigned* tickets;           // [m_tiles], zero before the call and after it
    const float* b3;             // fc.6 bias [16]
    float* logits; int32_t* cls; uint8_t* bits;      // outputs of this chunk (may be null)
    int N;                       // total output features
    int rw, tv;                  // rows per window / valid rows per window of the INPUT tape (conv modes)
    int n_valid;                 // EPI_FC_F32: valid rows
    long long* trace;            // optional clock64 timeline of CTA 0: [tile][8] (tools/trace_tapgemm.py)
    int dbg;                     // timing ablations (results invalid): 1 = every tile loads the A slabs of tile 0; 2 = skip epilogue stores
};

struct Tape {
    int rows, m_tiles, cap;      // logical rows, 128-row tiles, row capacity incl. guards
    int kch;
    size_t kch_stride, part_stride, bytes;
};
inline Tape make_tape(int64_t rows, int kch) {
    Tape t; t.rows = (int)rows; t.m_tiles = (int)((rows + 127) / 128); t.m_tiles += t.m_tiles & 1;   // even: MT = 2 tiles, CTA pairs
    t.cap = kGuard + t.m_tiles * 128 + 136; t.kch = kch;   // trailing guard: the fused kernels' 124-row tiles read up to 130 rows past a tile start
    t.kch_stride = (size_t)t.cap * 16; t.part_stride = t.kch_stride * kch; t.bytes = align_up(t.part_stride * 2, 256);
    return t;
}

// The fc.0 operand (X4) is row-major instead: [part][row = window][K = 4736] bf16, no guard rows.  block2 writes it in whole
// lines (a pooled row is 256 contiguous bytes), fc.0 reads it through a tensor map (boxes of 128 rows x 64 B).
inline Tape make_rowmajor(int64_t rows, int kch) {
    Tape t; t.rows = (int)rows; t.m_tiles = (int)((rows + 127) / 128); t.m_tiles += t.m_tiles & 1;
    t.cap = t.m_tiles * 128; t.kch = kch;
    t.kch_stride = 16; t.part_stride = (size_t)t.cap * kch * 16; t.bytes = align_up(t.part_stride * 2, 256);
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
// AT = 1: the A operand is row-major in global memory ([part][row][K] bf16: what block2 can write in whole lines) and
// arrives by tensor-map TMA, one box of 128 rows x 64 B (KSA = 4 kchunks) per M-tile and part, in the 64-byte swizzle.
// KS = 2 (fc.3, MT = 1): TWO issuer warps share one tile along K — issuer i takes the K-stages i, i + 2, ... into its own
// accumulator, the epilogue adds the two.  A single issuer at N = 128 spends ~93 cycles per 64-cycle MMA (mbarrier probe,
// descriptors, commit per stage, and the pipe queues only an MMA or two ahead); two issuers fill each other's gaps exactly
// as fc.0's two accumulators do, and every accumulator still sees its MMAs in one fixed order (bit-reproducible).
template <int BN, int TAPS, int KSA, int NSTAGE, int MT = 1, int WST = 0, int AT = 0, int KS = 1>
struct TapGemmCfg {
    static constexpr int A_PART = AT ? 128 * KSA * 16 : KSA * kSlabBytes;
    static constexpr int A_TILE = 2 * A_PART;                 // hi + lo slabs of one M-tile
    static constexpr int A_BYTES = MT * A_TILE;
    static constexpr int B_TAPCH = BN * 16;
    static constexpr int B_PART = TAPS * KSA * B_TAPCH;
    static constexpr int B_BYTES = 2 * B_PART;
    static constexpr int STAGE_BYTES = A_BYTES + (WST ? 0 : B_BYTES);
    static constexpr int WRES_BYTES = WST * B_BYTES;
    // BN = 240 (fc.0: 2048 = 8 x 240 + 128 output features -> 9 n-tiles x 16 M-tile pairs = 144 tiles for 148 SMs instead of
    // 128): the MMA is N = 240, the accumulator keeps a 256-column pitch, and the epilogue masks the columns that do not exist
    static constexpr int ACC_COLS = (BN == 240) ? 256 : BN;
    static constexpr int ACCS = MT * KS;                              // accumulators (= issuer warps) per tile
    static constexpr int NBUF = (2 * ACCS * ACC_COLS <= 512) ? 2 : 1; // accumulator buffers (epilogue / MMA overlap)
    static constexpr int TMEM_COLS = NBUF * ACCS * ACC_COLS;
    static_assert(KS == 1 || (KS == 2 && MT == 1 && WST == 0), "K-split issuers: one M-tile, streamed weights");
    static constexpr int BAR_BYTES = (2 * NSTAGE + 5) * 8 + 8;
    static constexpr int RING_BYTES = NSTAGE * STAGE_BYTES;
    static constexpr int SMEM_BYTES = RING_BYTES + WRES_BYTES + BAR_BYTES + kEpiWarps * (ACC_COLS / 2) * 4;
    static_assert(KSA % 2 == 0, "an MMA consumes two kchunks");
    static_assert(!AT || (KSA == 4 && TAPS == 1 && WST == 0 && STAGE_BYTES % 1024 == 0), "tensor-map A operand: Linear layers, 64-byte rows, swizzle atoms stay aligned");
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns: power of two");
    static_assert(STAGE_BYTES % 16 == 0, "bulk copies are 16-byte granular");
    static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
};

#ifndef DCE_TRACE
#define DCE_TRACE 0
#endif
#define TG_TRACE(k, ev) do { if (DCE_TRACE && p.trace && blockIdx.x == 0 && (k) < 60 && (threadIdx.x & 31) == 0) p.trace[(k) * 16 + (ev)] = clock64(); } while (0)

template <int BN, int TAPS, int KSA, int NSTAGE, int EPI, int MT = 1, int WST = 0, int AT = 0, int KS = 1>
__global__ void __launch_bounds__(tapgemm_threads(MT * KS), 1)
tapgemm_kernel(const __grid_constant__ TapGemmParams p) {
    constexpr int kProducerWarp0 = kEpiWarps + MT * KS;
    using Cfg = TapGemmCfg<BN, TAPS, KSA, NSTAGE, MT, WST, AT, KS>;
    constexpr int ACCS = Cfg::ACCS;
    constexpr int NBUF = Cfg::NBUF;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wres = smem + Cfg::RING_BYTES;                       // resident weight image (WST > 0)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES + Cfg::WRES_BYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* tfull = empty + NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* s_bias = reinterpret_cast<float*>(smem + Cfg::RING_BYTES + Cfg::WRES_BYTES + Cfg::BAR_BYTES);   // [8 warps][BN/2]
    float4* s_w3 = reinterpret_cast<float4*>(s_bias + kEpiWarps * (Cfg::ACC_COLS / 2));    // EPI_FC_LOGITS: [BN columns][16] fc.6 weights of this n-tile
    static_assert(EPI != EPI_FC_LOGITS || MT == 1, "the logit-share epilogue keeps one row per thread");
    static_assert(BN != 240 || EPI == EPI_FC_TAPE, "the masked 240-column tile is fc.0's");
    constexpr int ACC = Cfg::ACC_COLS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = (p.m_tiles / MT) * p.n_tiles;      // p.m_tiles is a multiple of MT (make_tape)

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { ptx::mbar_init(&full[i], kProdWarps); ptx::mbar_init(&empty[i], MT); }   // KS = 2: a slot belongs to ONE issuer (MT = 1)
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull[b], ACCS); ptx::mbar_init(&tempty[b], kEpiWarps); }
        ptx::mbar_init(wbar, 1);
        ptx::fence_barrier_init();
    }
    pdl_launch_dependents();                                    // the next kernel's prologue may overlap our tail
    if (warp == kMmaWarp) { ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS); ptx::tmem_relinquish(); }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
#if DCE_TRACE
    if (p.trace && blockIdx.x < 60 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[blockIdx.x * 16 + 12] = (long long)t_; }
#endif
    pdl_wait();                                                 // everything the previous kernel wrote is visible from here on
#if DCE_TRACE
    if (p.trace && blockIdx.x < 60 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[blockIdx.x * 16 + 13] = (long long)t_; }
#endif

    if (warp >= kProducerWarp0) {
        // ===== TMA producers (warp-uniform loops; one elected lane per warp issues) =====
        // copy c of a stage: c < NA -> A slab (mt, part, j) of 2080 B; c == NA -> the B block.  Warp pw issues c = pw, pw+4, ...
        {
            constexpr int NA = AT ? MT * 2 : MT * 2 * KSA;       // AT: one tensor-map box per (M-tile, part)
            constexpr uint32_t kACopy = AT ? Cfg::A_PART : kSlabBytes;
            const int pw = warp - kProducerWarp0;
            uint32_t my_bytes = 0;
            for (int c = pw; c <= NA; c += kProdWarps) my_bytes += (c < NA) ? kACopy : (WST ? 0 : Cfg::B_BYTES);
            if (AT && pw == 0 && ptx::elect_one()) ptx::tma_prefetch_desc(&p.a_map);
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
                            if (AT && c < NA) {
                                const int mt = c >> 1, part = c & 1;
                                ptx::tma_load_3d(st + mt * Cfg::A_TILE + part * Cfg::A_PART, &p.a_map, s * (KSA * 8),
                                                 128 * (((p.dbg & 1) ? 0 : m) + mt), part, &full[slot]);
                            } else if (c < NA) {
                                const int mt = c / (2 * KSA), part = (c / KSA) & 1, j = c % KSA;
                                ptx::bulk_g2s(st + mt * Cfg::A_TILE + part * Cfg::A_PART + j * kSlabBytes,
                                              a_row + (size_t)mt * 2048 + part * p.a_part_stride + (size_t)(s * KSA + j) * p.a_kch_stride,
                                              kSlabBytes, &full[slot]);
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
            const int iw = warp - kMmaWarp;                      // issuer = accumulator index
            const int mt = KS > 1 ? 0 : iw, ks = KS > 1 ? iw : 0;
            const bool leader = ptx::elect_one();
            const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const uint32_t total_stages = (uint32_t)my_tiles * p.stages;
            if (WST) ptx::mbar_wait(wbar, 0);
            uint32_t it = ks;
            if (my_tiles > 0) {
                ptx::mbar_wait(&tempty[0], 1);                       // fresh barrier: passes
                ptx::mbar_wait(&full[ks % NSTAGE], 0);
                ptx::tc_fence_after_sync();
            }
            for (int tcount = 0; tcount < my_tiles; ++tcount) {
                const uint32_t buf = tcount % NBUF;
                if (iw == 0) TG_TRACE(tcount, 2);
                const uint32_t d = tmem_base + buf * (ACCS * ACC) + iw * ACC;
                // A single accumulator buffer (fc.0: 2 x 256 columns fill the TMEM): the epilogue that frees it cannot start
                // before this issuer's own `tfull` commit of the previous tile, so the buffer is awaited here, at the top of
                // the tile, not by the mid-stage probe below (which would wait for it BEFORE that commit: a deadlock found
                // by tools/simulate_block2_protocol.py).
                if (NBUF == 1 && tcount > 0) { ptx::mbar_wait(&tempty[0], (tcount & 1) ^ 1); ptx::tc_fence_after_sync(); }
                for (int s = ks; s < p.stages; s += KS, it += KS) {     // p.stages is a multiple of KS
                    const uint32_t slot = it % NSTAGE;
                    const bool last_stage = s + KS >= p.stages;       // this issuer's last stage of the tile
                    if (iw == 0 && s == 0) TG_TRACE(tcount, 4);
                    if (iw == 0 && last_stage) TG_TRACE(tcount, 5);
                    const uint32_t a0 = ptx::smem_u32(smem + slot * Cfg::STAGE_BYTES) + mt * Cfg::A_TILE;
                    const uint32_t b0 = WST ? ptx::smem_u32(wres) + s * Cfg::B_BYTES : ptx::smem_u32(smem + slot * Cfg::STAGE_BYTES) + Cfg::A_BYTES;
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
                        const int arow = (TAPS == 1) ? 1 : tap;            // Linear layers read the centre row only
#pragma unroll
                        for (int kk = 0; kk < KSA / 2; ++kk) {
                            const uint32_t b_hi = b0 + (tap * KSA + 2 * kk) * Cfg::B_TAPCH;
                            const uint64_t db_hi = ptx::make_smem_desc(b_hi, Cfg::B_TAPCH, 128);
                            const uint64_t db_lo = ptx::make_smem_desc(b_hi + Cfg::B_PART, Cfg::B_TAPCH, 128);
                            const uint32_t first = (s == ks && tap == 0 && kk == 0) ? 0u : 1u;
                            const uint32_t a_hi = AT ? a0 + kk * 32 : a0 + (2 * kk) * kSlabBytes + arow * 16;
                            const uint64_t da_hi = AT ? ptx::make_smem_desc_sw64(a_hi) : ptx::make_smem_desc(a_hi, kSlabBytes, 128);
                            const uint64_t da_lo = AT ? ptx::make_smem_desc_sw64(a_hi + Cfg::A_PART) : ptx::make_smem_desc(a_hi + Cfg::A_PART, kSlabBytes, 128);
                            if (leader) {
                                ptx::umma_bf16_ss(d, da_hi, db_lo, idesc, first);      // small terms first
                                ptx::umma_bf16_ss(d, da_lo, db_hi, idesc, 1u);
                                ptx::umma_bf16_ss(d, da_hi, db_hi, idesc, 1u);
                            }
                            // mid-stage: probe what the NEXT stage needs while MMAs of this one are still queued
                            if (tap == (TAPS - 1) / 2 && kk == (KSA / 2 - 1) / 2 && it + KS < total_stages) {
                                if (NBUF > 1 && last_stage) {                  // next stage opens the next tile
                                    const uint32_t nt = tcount + 1;
                                    ptx::mbar_wait(&tempty[nt % NBUF], ((nt / NBUF) & 1) ^ 1);
                                }
                                ptx::mbar_wait(&full[(it + KS) % NSTAGE], ((it + KS) / NSTAGE) & 1);
                                ptx::tc_fence_after_sync();
                            }
                        }
                    }
                    if (leader) {
                        ptx::umma_commit(&empty[slot]);          // this issuer's MMAs on the slot have retired
                        if (last_stage) ptx::umma_commit(&tfull[buf]);          // this accumulator is complete
                    }
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias / ReLU / pool -> bf16 hi/lo -> next layer's tape =====
        // warp e: TMEM lane quadrant q = e % 4 (hardware restriction), column half h = e / 4.
        constexpr int HALF = ACC / 2;
        const int q = warp & 3, h = warp >> 2;
        const int row_in_tile = q * 32 + lane;
        float* my_bias = s_bias + warp * HALF;               // warp-private copy of this warp's bias slice
        uint32_t tcount = 0;
        int last_n = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int m0 = (tile / p.n_tiles) * MT, n = tile % p.n_tiles;
            const uint32_t buf = tcount % NBUF, tph = (tcount / NBUF) & 1;
            const int n0 = n * BN + h * HALF;                // first output feature this warp owns
            if (n != last_n) {                               // stage this warp's bias slice (warp-private copy)
                __syncwarp();
                for (int i = lane; i < HALF; i += 32) my_bias[i] = (h * HALF + i < BN && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
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
                    out_off[mt] = (valid[mt] && w < p.out_rows_cap) ? ((size_t)w * 4736 + (size_t)to * BN) * 2 : (size_t)-1;   // row-major [window][t * 128 + c]
                } else if (EPI == EPI_FC_TAPE) {
                    out_off[mt] = (size_t)(row + kGuard) * 16;
                }
            }
            // the chunk loop below is a real loop (its body must stay in the instruction cache: see there), so the per-M-tile
            // values are picked with selects, never by indexing a register array with a runtime value
            auto pick_i = [&](const int (&a)[MT], int mt) { return MT == 1 ? a[0] : (mt ? a[MT - 1] : a[0]); };
            auto pick_b = [&](const bool (&a)[MT], int mt) { return MT == 1 ? a[0] : (mt ? a[MT - 1] : a[0]); };
            auto pick_z = [&](const size_t (&a)[MT], int mt) { return MT == 1 ? a[0] : (mt ? a[MT - 1] : a[0]); };
            static_assert(MT <= 2, "pick_*: two M-tiles at most");
            constexpr int CPM = HALF / 32;                   // 32-column chunks per M-tile for this warp
            constexpr int NCH = MT * CPM;
            const uint32_t taddr0 = tmem_base + buf * (ACCS * ACC) + h * HALF + ((uint32_t)(q * 32) << 16);
            auto chunk_addr = [&](int ci) { return taddr0 + (ci / CPM) * ACC + (ci % CPM) * 32; };

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
                    y[i] = relu_nan(__uint_as_float(v[i]) + b4.x);
                    y[i + 1] = relu_nan(__uint_as_float(v[i + 1]) + b4.y);
                    y[i + 2] = relu_nan(__uint_as_float(v[i + 2]) + b4.z);
                    y[i + 3] = relu_nan(__uint_as_float(v[i + 3]) + b4.w);
                }
                if (EPI == EPI_FC_LOGITS) {
                    // fc.6 folded into fc.3's epilogue: H2 never leaves the registers (smem reads are warp broadcasts).  This layer's
                    // thread has two chunks, i.e. ONE pass of the loop below: straight-line code.  A short loop over eight columns
                    // (re-used instructions, software-pipelined TMEM loads) was measured too: 14.4 k cycles instead of 8.2 k — the
                    // 256 broadcast LDS.128 per thread want the deep interleaving the unrolled form gives them
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
                    if (pick_i(rows, mt) < p.n_valid) {
                        float* dst = p.out_f32 + (size_t)pick_i(rows, mt) * p.N + n0 + c0;
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4*>(dst + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
                    }
                    return;
                }
                if (EPI == EPI_POOL_TAPE || EPI == EPI_POOL_FC) {
                    // MaxPool1d(2,2): rows (2i, 2i+1) are adjacent lanes (src/contact_cnn.py:24-25,42-43)
#pragma unroll
                    for (int i = 0; i < 32; ++i) y[i] = max_nan(y[i], __shfl_xor_sync(0xffffffffu, y[i], 1));
                }
                if (EPI == EPI_TAPE || EPI == EPI_POOL_TAPE) {
                    if (!pick_b(valid, mt)) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) y[i] = 0.f;                  // guard rows stay zero
                    }
                }
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 hi, lo;
                    split8(y + qd * 8, hi, lo);
                    const int kch = (n0 + c0) / 8 + qd;
                    if (BN == 240 && (h * HALF + c0 + qd * 8 >= BN || n0 + c0 + qd * 8 >= p.N)) continue;   // columns of the pitch / of the last n-tile that do not exist
                    if (EPI == EPI_TAPE || EPI == EPI_FC_TAPE) {
                        uint8_t* dst = p.out + (size_t)kch * p.out_kch_stride + pick_z(out_off, mt);
                        *reinterpret_cast<uint4*>(dst) = hi;
                        *reinterpret_cast<uint4*>(dst + p.out_part_stride) = lo;
                        if (EPI == EPI_TAPE && pick_b(zero_prev, mt)) {
                            *reinterpret_cast<uint4*>(dst - 16) = make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(dst + p.out_part_stride - 16) = make_uint4(0, 0, 0, 0);
                        }
                    } else {
                        // pooled: even lane writes the hi part, odd lane the lo part of the pooled row
                        if (pick_z(out_off, mt) != (size_t)-1) {
                            uint8_t* dst = p.out + (EPI == EPI_POOL_FC ? (size_t)kch * 16 : (size_t)kch * p.out_kch_stride) + pick_z(out_off, mt) + ((lane & 1) ? p.out_part_stride : 0);
                            *reinterpret_cast<uint4*>(dst) = (lane & 1) ? lo : hi;
                            if (EPI == EPI_POOL_TAPE && pick_b(zero_prev, mt)) *reinterpret_cast<uint4*>(dst - 16) = make_uint4(0, 0, 0, 0);
                        }
                    }
                }
            };

            // KS = 2: the second issuer's accumulator (the odd K-stages) sits ACC columns further: added here
            auto fold = [&](uint32_t (&v)[32], int ci) {
                if (KS > 1) {
                    uint32_t w[32];
                    ptx::tmem_ld32(chunk_addr(ci) + ACC, w);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
                }
            };
            // software pipeline over the chunks: the TMEM load of chunk i+1 is in flight while chunk i is processed.  NOT
            // unrolled: a CTA runs this epilogue once or a few times per launch, so unrolled (fc.0: 8 chunks = 42 KB of
            // straight-line code) every instruction line is a cold instruction-cache miss — measured 3 500 cycles per
            // 330-instruction chunk, 28 k cycles per tile, 10 % of the kernel; as a loop the second pass on hits
            if (!(p.dbg & 2)) {
                uint32_t va[32], vb[32];
                ptx::tmem_ld32(chunk_addr(0), va);
#pragma unroll 1
                for (int ci = 0; ci < NCH; ci += 2) {
                    ptx::tmem_ld_wait();
                    if (ci + 1 < NCH) ptx::tmem_ld32(chunk_addr(ci + 1), vb);
                    fold(va, ci);
                    process(va, ci);
                    if (ci + 1 < NCH) {
                        ptx::tmem_ld_wait();
                        if (ci + 2 < NCH) ptx::tmem_ld32(chunk_addr(ci + 2), va);
                        fold(vb, ci + 1);
                        process(vb, ci + 1);
                    }
                }
            }
            if (EPI == EPI_FC_LOGITS && !(p.dbg & 2) && rows[0] < p.n_valid) {
                float4* dst = reinterpret_cast<float4*>(p.out_f32 + ((size_t)(n * 2 + h) * p.n_valid + rows[0]) * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(lg[4 * j], lg[4 * j + 1], lg[4 * j + 2], lg[4 * j + 3]);
            }
            if (EPI == EPI_FC_LOGITS && p.tickets) {
                // last CTA of this M-tile reduces: shares -> logits -> argmax -> contact bits (src/inference_one_seq.py:26-27,59-62)
                __threadfence();                                     // this thread's shares are visible before the ticket
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (warp == 0 && lane == 0) {
                    const unsigned old = atomicAdd(p.tickets + m0, 1u);
                    const bool last = (old == (unsigned)p.n_tiles - 1);
                    if (last) p.tickets[m0] = 0;                      // every arrival is in: the call leaves its tickets at zero
                    *reinterpret_cast<volatile uint32_t*>(tmem_slot + 1) = last ? 1u : 0u;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (*reinterpret_cast<volatile uint32_t*>(tmem_slot + 1) && warp < 4) {
                    __threadfence();                                 // acquire side: the other CTAs' shares
                    const int64_t w = 128 * (int64_t)m0 + warp * 32 + lane;
                    if (w < p.n_valid) {
                        float sacc[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) sacc[j] = 0.f;
                        for (int sidx = 0; sidx < 2 * p.n_tiles; ++sidx) {
                            const float4* src = reinterpret_cast<const float4*>(p.out_f32 + ((size_t)sidx * p.n_valid + w) * 16);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 v = __ldcg(src + j);
                                sacc[4 * j] += v.x; sacc[4 * j + 1] += v.y; sacc[4 * j + 2] += v.z; sacc[4 * j + 3] += v.w;
                            }
                        }
                        float yl[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) yl[j] = __ldg(p.b3 + j) + sacc[j];
                        fp32::argmax_bits_store(yl, w, p.logits, p.cls, p.bits);
                    }
                }
            }
            if (warp == 0) TG_TRACE(tcount, 8);
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[buf]);
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
#if DCE_TRACE
    if (p.trace && blockIdx.x < 60 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[blockIdx.x * 16 + 14] = (long long)t_; }
#endif
    if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// K0 (tensor-core part): weights -> bf16 hi/lo shared-memory images, one block per (n_tile, stage)
// kind 0: conv  W[n][cin][3]      K index (tap, c), c padded to 8*KSA*stages
// kind 1: fc.0  W[n][c*37 + t]    K index k' = t*128 + c      (src/contact_cnn.py:64)
// kind 2: fc.3  W[n][k]
// ---------------------------------------------------------------------------------------------
// stack = 1: the block holds, per (tap, kchunk), [W_hi (BN rows) ; W_lo (BN rows)] as ONE K-major operand of 2*BN rows
// (an N = 2*BN MMA against A_hi yields a_hi*w_hi and a_hi*w_lo side by side; dce_tc_block2s.cuh) instead of a hi image
// followed by a lo image.
__global__ void pack_b_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int n_tiles, int stages,
                              int BN, int TAPS, int KSA, int kind, int cin, int nout, int stack) {
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
        if (n >= nout) v = 0.f;                      // padding columns of a ragged last n-tile
        else if (kind == 0) v = (c < cin) ? W[((size_t)n * cin + c) * 3 + tap] : 0.f;
        else if (kind == 1) { const int t = c / 128, ch = c % 128; v = W[(size_t)n * 4736 + ch * 37 + t]; }
        else v = W[(size_t)n * cin + c];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        const size_t blk = ((size_t)nt * stages + s) * (2 * per_part * 2);      // bytes
        if (stack == 2) {
            // CTA-pair images (block1_kernel, BN = 64): [rank 2][tap][kchunk][96 rows][8].  Rows 0..63: rank 0 W_hi, rank 1 W_lo (each CTA's
            // half of the stacked N = 128 operand); rows 64..95: W_hi rows 32 r .. 32 r + 31 (its half of the N = 64 operand)
            const size_t rank_bytes = (size_t)TAPS * KSA * 96 * 16, cell = (((size_t)tap * KSA + j) * 96) * 16 + e * 2;
            *reinterpret_cast<__nv_bfloat16*>(out + blk + cell + (size_t)nn * 16) = h;
            *reinterpret_cast<__nv_bfloat16*>(out + blk + rank_bytes + cell + (size_t)nn * 16) = l;
            *reinterpret_cast<__nv_bfloat16*>(out + blk + (nn >> 5) * rank_bytes + cell + (size_t)(64 + (nn & 31)) * 16) = h;
            continue;
        }
        if (stack) {
            const size_t within = ((((size_t)tap * KSA + j) * 2 * BN + nn) * 8 + e) * 2;
            *reinterpret_cast<__nv_bfloat16*>(out + blk + within) = h;
            *reinterpret_cast<__nv_bfloat16*>(out + blk + within + (size_t)BN * 16) = l;
            continue;
        }
        const size_t within = ((((size_t)tap * KSA + j) * BN + nn) * 8 + e) * 2;
        *reinterpret_cast<__nv_bfloat16*>(out + blk + within) = h;
        *reinterpret_cast<__nv_bfloat16*>(out + blk + per_part * 2 + within) = l;
    }
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
struct LayerCfg { int BN, TAPS, KSA, stages, n_tiles, kind, cin, src, nout, stack; };
//                                   BN  TAPS KSA stages n_tiles kind cin   state_dict index of the weight
constexpr int kNumPacked = 10;
constexpr LayerCfg kLayers[kNumPacked] = {{64, 3, 4, 2, 1, 0, 54, 0, 64},       // block1.0
                                 {64, 3, 4, 2, 1, 0, 64, 2, 64},       // block1.2
                                 {128, 3, 4, 2, 1, 0, 64, 4, 128},     // block2.0 (layer-wise conv3: resident image)
                                 {128, 3, 2, 8, 1, 0, 128, 6, 128},    // block2.2
                                 {240, 1, 4, 148, 9, 1, 4736, 8, 2048},  // fc.0: 8 n-tiles of 240 + one of 128 (zero-padded to 240)
                                 {128, 1, 4, 64, 4, 2, 2048, 10, 512}, // fc.3
                                 {128, 3, 2, 4, 1, 0, 64, 4, 128, 1},  // block2.0 in 24 KB ring blocks with the stacked [W_hi ; W_lo] operand (block2_kernel)
                                 {128, 3, 2, 8, 1, 0, 128, 6, 128, 1},   // block2.2, stacked
                                 {64, 3, 8, 1, 1, 0, 54, 0, 64, 2},      // block1.0, stacked, one image per CTA of a pair (block1_kernel)
                                 {64, 3, 8, 1, 1, 0, 64, 2, 64, 2}};     // block1.2, the same
constexpr int kLayerConv3Stack = 6, kLayerConv4Stack = 7, kLayerConv1Stack = 8, kLayerConv2Stack = 9;     // images of the fused kernels
inline size_t layer_packed_bytes(const LayerCfg& c) {
    if (c.stack == 2) return (size_t)2 * c.TAPS * c.KSA * 96 * 16;          // two CTA-pair images
    return (size_t)c.n_tiles * c.stages * 2 * c.TAPS * c.KSA * c.BN * 16;
}

struct PackedLayout { size_t w[kNumPacked]; size_t begin, end; };
inline PackedLayout make_packed_layout(size_t base) {
    PackedLayout L; L.begin = base; size_t o = base;
    for (int i = 0; i < kNumPacked; ++i) { L.w[i] = o; o = align_up(o + layer_packed_bytes(kLayers[i]), 256); }
    L.end = o;
    return L;
}

inline int pack(char* buf, const PackedLayout& L, const float* const* params, Ctx& ctx) {
    for (int i = 0; i < kNumPacked; ++i) {
        const LayerCfg& c = kLayers[i];
        const size_t total = (size_t)c.n_tiles * c.stages * c.TAPS * c.KSA * c.BN * 8;
        const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        DCE_KL(ctx, "tc_pack_b", pack_b_kernel<<<blocks, 256, 0, ctx.stream>>>(
            params[c.src], reinterpret_cast<uint8_t*>(buf + L.w[i]), c.n_tiles, c.stages, c.BN, c.TAPS, c.KSA, c.kind, c.cin, c.nout, c.stack));
    }
    return DCE_OK;
}

struct Workspace {
    Tape x0, x1, x2, x3, x4, h1;
    size_t o_x0, o_x1, o_x2, o_x3, o_x4, o_h1, o_h2, o_tickets, o_mean, o_sdev, end;
};
inline Workspace make_workspace(int64_t n) {
    Workspace W; size_t o = 512;          // [0,256): the latency kernel's barrier counters (dce_latency.cuh); [256,512): fc.3's per-M-tile
                                          // tickets (at a fixed place: every other offset depends on the chunk size).  Zero once; every call leaves them at zero.
    W.o_tickets = 256;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    W.x0 = make_tape(n * kRW1, 8);   W.o_x0 = take(W.x0.bytes);
    W.x1 = make_tape(n * kRW1, 8);   W.o_x1 = take(W.x1.bytes);
    W.x2 = make_tape(n * kRW2, 8);   W.o_x2 = take(W.x2.bytes);
    W.x3 = make_tape(n * kRW2, 16);  W.o_x3 = take(W.x3.bytes);
    W.x4 = make_rowmajor(n, 592);    W.o_x4 = take(W.x4.bytes);
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

template <int BN, int TAPS, int KSA, int NSTAGE, int EPI, int MT = 1, int WST = 0, int AT = 0, int KS = 1>
inline int launch_layer(Ctx& ctx, const char* name, int sm_count, const TapGemmParams& p) {
    using Cfg = TapGemmCfg<BN, TAPS, KSA, NSTAGE, MT, WST, AT, KS>;
    auto kern = tapgemm_kernel<BN, TAPS, KSA, NSTAGE, EPI, MT, WST, AT, KS>;
    if (KS > 1 && p.stages % KS) return DCE_EINVAL;
    if (WST && (p.stages != WST || p.n_tiles != 1)) return DCE_EINVAL;
    constexpr int kSmem = Cfg::SMEM_BYTES + (EPI == EPI_FC_LOGITS ? BN * 64 : 0);      // + [BN][16] fp32 of the next layer
    static_assert(kSmem <= 232448, "exceeds 227 KB");
    static DeviceOnce attr_once;
    if (auto first_ = attr_once.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) { first_.fail(); ctx.err = e; return DCE_ECUDA; }
    }
    const int tiles = (p.m_tiles / MT) * p.n_tiles;
    const int grid = tiles < sm_count ? tiles : sm_count;         // persistent: a CTA walks tiles blockIdx.x, + gridDim.x, ...
    DCE_KL(ctx, name, { cudaError_t le_ = launch_pdl(kern, dim3(grid), dim3(tapgemm_threads(MT * KS)), kSmem, ctx.stream, p); (void)le_; });
    return DCE_OK;
}

}  // namespace tc
}  // namespace dce
