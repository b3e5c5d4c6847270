// Fused block1: window ingest (+ optional z-score) -> Conv1d 54->64 + ReLU -> Conv1d 64->64 + ReLU
// -> MaxPool1d(2,2), one persistent kernel; activations between the two convolutions never
// leave the SM.   /root/reference/src/contact_cnn.py:10-26, utils/data_handler.py:55-56
//
// Arithmetic per K-step (stacked B operand): every (tap, kchunk) of a weight image holds [W_hi (64 rows) ; W_lo (64 rows)]
// as ONE K-major operand, so
//      MMA 1 (N = 128):  A_hi x [W_hi ; W_lo]  -> D[:, 0:64] += a_hi w_hi,  D[:, 64:128] += a_hi w_lo
//      MMA 2 (N =  64):  A_lo x  W_hi          -> D[:, 0:64] += a_lo w_hi
// i.e. 14 KB instead of 18 KB of shared-memory operand reads per K-step (these small-N MMAs are bound by the 128 B/clk
// operand fetch, not by the tensor pipe: 64 + 48 cycles instead of 3 x 48); the epilogues add the two accumulator halves.
//
// Tiling ("recompute nothing, shrink the valid range"): a tile produces 124 rows of conv2
// output.  With b = 124*i:
//      slab0 row s  <->  X0 row b-3+s   (s = 0..129)   fp32 input -> bf16 hi/lo, written by converter warps
//      conv1 MMA row k  <->  X1 row b-2+k  (k = 0..127), reads slab0 rows k..k+2; result -> slab1 row k+1
//      conv2 MMA row j  <->  out row b-2+j (j = 0..127), reads slab1 rows j..j+2; rows j in [2,126) are exact
// so each tile is two 128-row UMMA passes for 124 useful rows (96.9 %), pool pairs (j, j+1) with j
// even sit in adjacent lanes (the two lanes of a pair split the pooled columns between them), and both conv weight
// images (2 x 48 KB bf16 hi/lo) stay resident in shared memory for the life of the CTA.  slab0 and slab1 have two
// buffers each; the room for the second slab1 comes from dropping the all-zero kchunk 7 (channels 56..63 do not exist)
// of the slab0 images: the K = 16 MMA that covers kchunks 6 and 7 takes its second kchunk from ONE shared zero chunk
// through its leading-dimension offset.
//
// CTA pairs (tcgen05.mma.cta_group::2): the two CTAs of a cluster, on the two SMs of a TPC, work on two consecutive tiles
// in lockstep.  One M = 256 MMA covers both: each CTA supplies its own 128 rows of A and only HALF of the B operand (MMA 1:
// rank 0 holds W_hi, rank 1 W_lo; MMA 2: each holds 32 of W_hi's 64 rows) from the same shared-memory offsets, and gets its
// own 128 x N slice of D in its own TMEM.  Per SM and K-step the operand fetch drops from 14 to 11 KB and a CTA keeps
// 2 x 36 KB of weights instead of 2 x 48, which is what makes room for the output staging tile below.  Only the leader CTA
// (rank 0) issues MMAs; the peer's two issuer warps relay "my operand slab is written and my accumulator is free" to the
// leader with one remote mbarrier arrive per tile and convolution; tcgen05.commit is multicast to the barriers of both.
//
// Warp roles (16 warps, 1 CTA/SM):
//   warps 0-3  : converters   staged fp32 rows -> (z-score) -> bf16 hi/lo, in place in slab0[buf] (UMMA K-major layout)
//   warps 4-11 : epilogue     epi1: TMEM D1 -> bias/ReLU/guard -> bf16 hi/lo -> slab1 (smem, feeds conv2)
//                             epi2: TMEM D2 -> bias/ReLU/pool/guard -> bf16 hi/lo -> X2 tape (global)
//   warp 13    : loader       raw fp32 rows of a tile -> slab buffer by bulk TMA (<= 2 copies), L2 prefetch ahead
//   warp 12    : conv1 MMA issuer (+ weights via bulk TMA once)
//   warp 14    : conv2 MMA issuer — separate issuers, so the tensor pipe always has the other conv's MMAs queued
//                             while one issuer is between tiles or waiting for epi1 to fill slab1
//   warp 15    : store warp   epi2 leaves the tile's 62 pooled rows in a staging tile in the tape's own [part][kchunk][row]
//                             order: 16 contiguous runs of 992 B, which this warp hands to the copy engine (cp.async.bulk
//                             shared -> global) while the epilogue warps go on.  (Staging them in the slab1 buffer conv2 has
//                             just drained put the copies on epi1's path two tiles later: no gain.)  Issued as
//                             st.global from the epilogue warps the same bytes cost 12 % of the kernel (ablation
//                             block1_dbg=1): the epilogue warps are the kernel's critical resource
#pragma once
#include "dce_tc.cuh"

namespace dce {
namespace tc {

constexpr int kB1Rows = 124;                       // useful conv2 rows per tile
constexpr int kB1Threads = 16 * 32;
constexpr int kB1SlabBytes = 2 * 8 * kSlabBytes;   // [part][8 kchunks][130 rows][16 B] = 33280
constexpr int kB1WRows = 96;                       // per (tap, kchunk) and CTA: 64 rows for MMA 1 (rank 0: W_hi, rank 1: W_lo) + 32 for MMA 2 (W_hi rows 32 r ..)
constexpr int kB1WBytes = 3 * 8 * kB1WRows * 16;   // one CTA's conv weight image: [tap 3][kchunk 8][96 rows][8] bf16 = 36864
constexpr int kB1Slab0 = 2 * 7 * kSlabBytes;        // slab0 buffer: [part][7 kchunks][130 rows][16 B] = 29120 (also stages the raw fp32 rows: <= 28128 B)
constexpr int kB1StageRun = (kB1Rows / 2) * 16;     // 62 pooled rows x 16 B: one (part, kchunk) run of the X2 tape
constexpr int kB1Stage = 16 * kB1StageRun;          // 15872
constexpr int kB1SmemBytes = 2 * kB1WBytes + 2 * kB1Slab0 + kSlabBytes + 2 * kB1SlabBytes + 256 + 2 * 64 * 4 + 2 * 2 * 64 * 4 + kB1Stage;
static_assert(kB1SmemBytes <= 232448, "exceeds 227 KB");

struct Block1Params {
    const float* x;              // batch: [W][150][54]; stream: [T][54]
    int64_t first;               // stream: first window row
    int n_windows;
    const float* mean;           // stream: [W][64]
    const float* rstd;           // 1 / unbiased std
    const uint8_t* w1; const uint8_t* w2;     // packed images (tc::pack layout, layers 0 and 1)
    const float* b1; const float* b2;
    uint8_t* out;                // X2 tape (part 0)
    size_t out_part_stride, out_kch_stride;
    int out_rows_cap;
    int n_tiles;
    int64_t total_rows;          // stream: rows T of the log (to keep the 16-byte-rounded bulk copies in bounds)
    int dbg;                     // timing ablations only (results invalid when non-zero); w1 / w2: [rank 2] images of kB1WBytes
    long long* trace;            // optional: CTA 0 records clock64() per role per tile ([tile][16])
};
#define B1_TRACE(k, ev) do { if (DCE_TRACE && p.trace && blockIdx.x == 0 && (k) < 60 && (threadIdx.x & 31) == 0) p.trace[(k) * 16 + (ev)] = clock64(); } while (0)

__device__ __forceinline__ int pos_mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

// The raw fp32 rows a tile needs form at most two contiguous runs in global memory (one per window
// it touches; the guard rows between windows have no source).  Rows are 216 B, so a run starts on
// an 8-byte boundary only: each run is fetched with ONE bulk TMA copy of the enclosing 16-byte
// aligned range and lands at `off` (its first row at off, i.e. aligned base + phase).
struct TileSegs {
    int s_lo[2], n[2];           // first slab row / row count of each run (n == 0: absent)
    uint32_t off[2];             // byte offset, inside the staging buffer, of the run's first row
    uint32_t dst_al[2];          // 16-byte aligned staging offset the bulk copy writes to
    const char* src_al[2];       // 16-byte aligned global source
    uint32_t bytes[2];           // bulk copy size (multiple of 16)
    bool tail8[2];               // stream mode, run ends at the end of the log on an 8-byte boundary:
                                 //   the last 16-byte chunk is not bulk-copied; its 8 valid bytes are copied by hand
};
template <bool STREAM>
__device__ __forceinline__ TileSegs tile_segs(const Block1Params& p, int r0) {
    TileSegs g;
    const int wf = (r0 >= 0) ? r0 / kRW1 : -1;
    uint32_t cursor = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int wi = wf + j;
        const int lo = r0 > wi * kRW1 ? r0 : wi * kRW1;
        const int hi = (r0 + kSlabRows) < (wi * kRW1 + 150) ? (r0 + kSlabRows) : (wi * kRW1 + 150);
        g.s_lo[j] = lo - r0; g.n[j] = 0; g.off[j] = 0; g.dst_al[j] = 0; g.src_al[j] = nullptr; g.bytes[j] = 0; g.tail8[j] = false;
        if (wi < 0 || wi >= p.n_windows || hi <= lo) continue;
        const int t0 = lo - wi * kRW1;
        const int64_t row = STREAM ? (p.first + wi + t0) : ((int64_t)wi * 150 + t0);
        const char* src = reinterpret_cast<const char*>(p.x) + row * 216;
        const uint32_t ps = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
        const uint32_t len = (uint32_t)(hi - lo) * 216;
        uint32_t bytes = (ps + len + 15) & ~15u;
        if (STREAM && ((ps + len) & 15) && row + (hi - lo) == p.total_rows) { bytes -= 16; g.tail8[j] = true; }
        g.n[j] = hi - lo; g.dst_al[j] = cursor; g.off[j] = cursor + ps; g.src_al[j] = src - ps; g.bytes[j] = bytes;
        cursor += ((ps + len + 15) & ~15u);
    }
    return g;
}

template <bool STREAM>
__global__ void __launch_bounds__(kB1Threads, 1)
block1_kernel(const Block1Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w1s = smem;
    uint8_t* w2s = smem + kB1WBytes;
    uint8_t* slab0 = smem + 2 * kB1WBytes;                  // two buffers of 7 kchunks x hi/lo
    uint8_t* zchunk = slab0 + 2 * kB1Slab0;                // one all-zero kchunk: "kchunk 7" of every slab0 image
    uint8_t* slab1 = zchunk + kSlabBytes;                   // two buffers of 8 kchunks x hi/lo
    uint64_t* bars = reinterpret_cast<uint64_t*>(slab1 + 2 * kB1SlabBytes);
    uint64_t* x0_full = bars;        // [2] 128 converter threads arrive
    uint64_t* x0_empty = bars + 2;   // [2] tcgen05.commit
    uint64_t* d1_full = bars + 4;    // [2] commit
    uint64_t* d1_empty = bars + 6;   // [2] 8 epilogue warps arrive
    uint64_t* d2_full = bars + 8;    // [2] commit
    uint64_t* d2_empty = bars + 10;  // [2] 8 epilogue warps arrive
    uint64_t* x1_full = bars + 12;   // [2] 256 epilogue threads arrive
    uint64_t* x1_empty = bars + 14;  // [2] commit
    uint64_t* wbar = bars + 16;      // weights landed
    uint64_t* raw_full = bars + 17;  // [2] raw fp32 rows of a tile landed (bulk TMA, bytes)
    uint64_t* st_full = bars + 19;   // 256 epilogue threads left the pooled rows of a tile in the staging tile
    uint64_t* st_done = bars + 21;   // the copy engine has read them
    uint64_t* p1_ready = bars + 23;  // [2] leader only: the PEER's slab0[buf] is written and its D1[buf] is free (one remote arrive)
    uint64_t* p2_ready = bars + 25;  // [2] leader only: the peer's slab1[buf] is written and its D2[buf] is free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // b1[64], b2[64]
    float* s_nrm = s_bias + 128;     // stream mode: [2 windows][mean 64 | 1/std 64] of the tile being converted
    uint8_t* stage = reinterpret_cast<uint8_t*>(s_nrm + 256);     // kB1Stage bytes: pooled rows on their way to the X2 tape

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NR = p.n_windows * kRW1;
    // pair j of the grid walks the tile pairs j, j + npairs, ...; rank r of the pair takes tile 2 * (pair tile) + r (an odd
    // tail tile does not exist: its rows lie beyond the chunk, its operands are zeros, nothing of it is stored)
    const uint32_t rank = ptx::cluster_ctarank();
    const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
    const int my_tiles = ((p.n_tiles + 1) / 2 - pair + npairs - 1) / npairs;
    auto tile_of = [&](int k) { return 2 * (pair + k * npairs) + (int)rank; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&p1_ready[i], 1); ptx::mbar_init(&p2_ready[i], 1); }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&x0_full[i], 128); ptx::mbar_init(&x0_empty[i], 1);
            ptx::mbar_init(&d1_full[i], 1);   ptx::mbar_init(&d1_empty[i], 8);
            ptx::mbar_init(&d2_full[i], 1);   ptx::mbar_init(&d2_empty[i], 8);
        }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&x1_full[i], 256); ptx::mbar_init(&x1_empty[i], 1); }
        ptx::mbar_init(st_full, 256); ptx::mbar_init(st_done, 1);
        ptx::mbar_init(wbar, 1);
        ptx::mbar_init(&raw_full[0], 1); ptx::mbar_init(&raw_full[1], 1);
        ptx::fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 12) { ptx::tmem_alloc_pair(tmem_slot, 512); ptx::tmem_relinquish_pair(); }
    // zero the activation slabs once: the shared zero kchunk and the never-written halo rows of slab1 must not hold
    // NaN bit patterns
    for (int i = threadIdx.x; i < (2 * kB1Slab0 + kSlabBytes + 2 * kB1SlabBytes) / 16; i += kB1Threads)
        reinterpret_cast<uint4*>(slab0)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 128) s_bias[threadIdx.x] = __ldg((threadIdx.x < 64 ? p.b1 : p.b2 - 64) + threadIdx.x);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    ptx::cluster_sync();                                        // the peer's barriers exist before anything remote touches them
    // The X2 tape we overwrite may still be read by the previous step's block2, and the stream-mode statistics come from the
    // kernel before us: converters, epilogues and issuers wait.  The conv weights (warp 12) and the raw input rows (loader,
    // warp 13) are written by no kernel of the step: both are requested while the previous kernel drains.
    if (warp == 12 && ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(wbar, 2 * kB1WBytes);
        ptx::bulk_g2s(w1s, p.w1 + (size_t)rank * kB1WBytes, kB1WBytes, wbar);
        ptx::bulk_g2s(w2s, p.w2 + (size_t)rank * kB1WBytes, kB1WBytes, wbar);
    }
    if (warp != 13) pdl_wait();

    if (warp < 4) {
        // ===== converters: fp32 rows -> slab0[buf] =====
        // The loader warp has bulk-copied the tile's raw rows INTO the slab buffer itself; thread t
        // reads row t back (27 x 8 B), the converter warps synchronise on a named barrier, and the
        // bf16 hi/lo image is written in place in the UMMA K-major layout.
        const int tid = threadIdx.x;
        auto row_window = [&](int r, int& w) -> bool {           // false: guard / out-of-range row -> zeros
            if (r < 0 || r >= NR) return false;
            w = r / kRW1;
            return r - w * kRW1 < 150;
        };
        for (int k = 0; k < my_tiles; ++k) {
            const int r0 = tile_of(k) * kB1Rows - 3;
            const uint32_t buf = k & 1;
            if (warp == 0) B1_TRACE(k, 0);
            const TileSegs sg = tile_segs<STREAM>(p, r0);
            ptx::mbar_wait_relaxed(&raw_full[buf], (k >> 1) & 1);                           // raw rows of tile k have landed
            if (warp == 0) B1_TRACE(k, 1);
            if (p.dbg & 4) { ptx::mbar_arrive(&x0_full[buf]); continue; }
            const int wbase = (r0 > 0 ? r0 : 0) / kRW1;                              // first window this tile touches
            if (STREAM) {                                                            // its z-score constants -> smem
                const int wi = wbase + (tid >> 6), c = tid & 63;
                const bool in = wi < p.n_windows;
                s_nrm[(tid >> 6) * 128 + c] = in ? __ldg(p.mean + (size_t)wi * 64 + c) : 0.f;
                s_nrm[(tid >> 6) * 128 + 64 + c] = in ? __ldg(p.rstd + (size_t)wi * 64 + c) : 1.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            if (warp == 0) B1_TRACE(k, 2);
            auto row_ptr = [&](int srow) -> const uint8_t* {                         // where slab row srow's raw data sits
                const int j = (sg.n[1] > 0 && srow >= sg.s_lo[1]) ? 1 : 0;
                return slab0 + buf * kB1Slab0 + sg.off[j] + (srow - sg.s_lo[j]) * 216;
            };
            uint8_t* dst0 = slab0 + buf * kB1Slab0;
            float2 f[27];
            int w0 = 0;
            const bool v0 = row_window(r0 + tid, w0);
            const uint8_t* src_row0 = row_ptr(tid);
#pragma unroll
            for (int i = 0; i < 27; ++i)
                f[i] = v0 ? *reinterpret_cast<const float2*>(src_row0 + 8 * i) : make_float2(0.f, 0.f);
            float2 g[4];
            int w1 = 0;
            const int s1 = 128 + tid / 7, kch1 = tid % 7;
            const bool v1 = (tid < 14) && row_window(r0 + s1, w1);
            const uint8_t* src_row1 = row_ptr(s1);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                g[i] = (v1 && (kch1 < 6 || i < 3)) ? *reinterpret_cast<const float2*>(src_row1 + kch1 * 32 + 8 * i)
                                                   : make_float2(0.f, 0.f);
            if (STREAM) {
                if (v0) {
                    const float2* mu = reinterpret_cast<const float2*>(s_nrm + (w0 - wbase) * 128);
                    const float2* sd = mu + 32;
#pragma unroll
                    for (int i = 0; i < 27; ++i) {                 // utils/data_handler.py:55-56: (x - mean) / std, as (x - mean) * (1 / std)
                        const float2 m2 = mu[i], s2 = sd[i];
                        f[i].x = (f[i].x - m2.x) * s2.x; f[i].y = (f[i].y - m2.y) * s2.y;
                    }
                }
                if (v1) {
                    const float2* mu = reinterpret_cast<const float2*>(s_nrm + (w1 - wbase) * 128 + kch1 * 8);
                    const float2* sd = mu + 32;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (kch1 < 6 || i < 3) {
                            const float2 m2 = mu[i], s2 = sd[i];
                            g[i].x = (g[i].x - m2.x) * s2.x; g[i].y = (g[i].y - m2.y) * s2.y;
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");                           // all raw reads done: overwrite in place
            if (warp == 0) B1_TRACE(k, 15);
#pragma unroll
            for (int kch = 0; kch < 7; ++kch) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool have = (kch < 6 || i < 3);              // channels 54, 55 do not exist
                    y[2 * i] = have ? f[(kch * 4 + i) % 27].x : 0.f;
                    y[2 * i + 1] = have ? f[(kch * 4 + i) % 27].y : 0.f;
                }
                uint4 hi, lo;
                split8(y, hi, lo);
                uint8_t* d = dst0 + kch * kSlabBytes + tid * 16;
                *reinterpret_cast<uint4*>(d) = hi;
                *reinterpret_cast<uint4*>(d + 7 * kSlabBytes) = lo;
            }
            if (tid < 14) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) { y[2 * i] = g[i].x; y[2 * i + 1] = g[i].y; }
                uint4 hi, lo;
                split8(y, hi, lo);
                uint8_t* d = dst0 + kch1 * kSlabBytes + s1 * 16;
                *reinterpret_cast<uint4*>(d) = hi;
                *reinterpret_cast<uint4*>(d + 7 * kSlabBytes) = lo;
            }
            ptx::fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async proxy
            ptx::mbar_arrive(&x0_full[buf]);
            if (warp == 0) B1_TRACE(k, 3);
        }
    } else if (warp == 13) {
        // ===== loader: raw fp32 rows of tile k -> slab0[k & 1] by bulk TMA, as soon as conv1(k-2) has drained it =====
        const uint64_t pol_once = ptx::policy_evict_first();     // the input windows are read exactly once: first out of the L2
        for (int k = 0; k < my_tiles; ++k) {
            const int r0 = tile_of(k) * kB1Rows - 3;
            const uint32_t buf = k & 1;
            const TileSegs sg = tile_segs<STREAM>(p, r0);
            ptx::mbar_wait_relaxed(&x0_empty[buf], ((k >> 1) & 1) ^ 1, 256);
            if (ptx::elect_one()) {
                B1_TRACE(k, 14);
                uint8_t* stage = slab0 + buf * kB1Slab0;
                uint32_t total = 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (sg.n[j] > 0 && sg.tail8[j]) {        // 8 valid bytes at the very end of the log, copied by hand
                        const unsigned long long v = *reinterpret_cast<const unsigned long long*>(sg.src_al[j] + sg.bytes[j]);
                        *reinterpret_cast<unsigned long long*>(stage + sg.dst_al[j] + sg.bytes[j]) = v;
                    }
                    total += sg.bytes[j];
                }
                if (total > 0) {
                    ptx::mbar_arrive_expect_tx(&raw_full[buf], total);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        if (sg.bytes[j] > 0) {
                            if (p.dbg & 16) ptx::bulk_g2s(stage + sg.dst_al[j], sg.src_al[j], sg.bytes[j], &raw_full[buf]);
                            else ptx::bulk_g2s_hint(stage + sg.dst_al[j], sg.src_al[j], sg.bytes[j], &raw_full[buf], pol_once);
                        }
                } else {
                    ptx::mbar_arrive(&raw_full[buf]);
                }
                // warm L2 with the tile that will use this buffer next
                if (k + 2 < my_tiles) {
                    const TileSegs nx = tile_segs<STREAM>(p, tile_of(k + 2) * kB1Rows - 3);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        if (nx.bytes[j] > 0)
                        {
                            if (p.dbg & 16) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx.src_al[j]), "r"(nx.bytes[j]) : "memory");
                            else ptx::bulk_prefetch_l2_hint(nx.src_al[j], nx.bytes[j], pol_once);
                        }
                }
            }
            __syncwarp();
        }
    } else if (warp == 15) {
        // ===== store warp: the 62 pooled rows of tile k, staged as 16 runs (part, kchunk) of 992 B -> X2 tape =====
        for (int k = 0; k < my_tiles; ++k) {
            const int tile = tile_of(k);
            ptx::mbar_wait_relaxed(st_full, k & 1);
            int rows = p.out_rows_cap - tile * (kB1Rows / 2);
            rows = rows < kB1Rows / 2 ? rows : kB1Rows / 2;
            if (lane < 16 && rows > 0 && !(p.dbg & 1)) {
                const int part = lane >> 3, kc = lane & 7;
                ptx::bulk_s2g(p.out + (part ? p.out_part_stride : 0) + (size_t)kc * p.out_kch_stride + (size_t)(tile * (kB1Rows / 2) + kGuard) * 16,
                              stage + lane * kB1StageRun, (uint32_t)rows * 16);
                ptx::bulk_commit_group();
                ptx::bulk_wait_group_read0();                         // the staging tile may be overwritten
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(st_done);
        }
        if (lane < 16) ptx::bulk_wait_group0();                        // every write has been performed before the CTA exits
    } else if (warp == 12 || warp == 14) {
        // ===== MMA issuers (warp-uniform control flow; one elected lane issues): warp 12 = conv1, warp 14 = conv2.
        // Two issuers so that one's serial code between tiles (mbarrier probes, fences, descriptors: ~500 cycles, more
        // than the tensor pipe's short queue covers) is filled by the other's MMAs; each accumulator keeps one issuer.
        {
            ptx::mbar_wait(wbar, 0);
            if (rank != 0) {
                // ===== peer CTA: no MMA is issued here.  Each issuer warp tells the leader, once per tile, that this CTA's operand slab
                // is written and its accumulator buffer is free
                const bool one = ptx::elect_one();
                for (int k = 0; k < my_tiles; ++k) {
                    const uint32_t buf = k & 1, ph = (k >> 1) & 1;
                    if (warp == 12) { ptx::mbar_wait(&x0_full[buf], ph); ptx::mbar_wait(&d1_empty[buf], ph ^ 1); }
                    else            { ptx::mbar_wait(&x1_full[buf], ph); ptx::mbar_wait(&d2_empty[buf], ph ^ 1); }
                    if (one) ptx::mbar_arrive_remote(ptx::mapa(ptx::smem_u32(warp == 12 ? &p1_ready[buf] : &p2_ready[buf]), 0));
                    __syncwarp();
                }
            } else {
            constexpr uint32_t idesc128 = ptx::make_idesc_bf16_f32(256, 128);
            constexpr uint32_t idesc64 = ptx::make_idesc_bf16_f32(256, 64);
            const uint32_t w1a = ptx::smem_u32(w1s), w2a = ptx::smem_u32(w2s);
            const uint32_t s0a = ptx::smem_u32(slab0), s1a = ptx::smem_u32(slab1), za = ptx::smem_u32(zchunk);

            // 24 MMAs of M = 256 (both CTAs): 3 taps x 4 kchunk pairs x (A_hi x [W_hi ; W_lo] at N = 128, A_lo x W_hi at N = 64); then
            // the two completion commits, multicast to both CTAs.  KCH = kchunks per part of the A image; with KCH = 7 the pair
            // (6, 7) takes kchunk 7 from the shared zero chunk through its leading-dimension offset.
            auto conv_mmas = [&](uint32_t a_base, int kch, uint32_t w_base, uint32_t d, uint64_t* bar_a, uint64_t* bar_b) {
                if (ptx::elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t a_hi = a_base + (2 * kk) * kSlabBytes + tap * 16;
                            const uint32_t a_lo = a_hi + (uint32_t)kch * kSlabBytes;
                            const bool zpair = (kch == 7 && kk == 3);
                            const uint64_t da_hi = ptx::make_smem_desc(a_hi, zpair ? (za + tap * 16 - a_hi) : (uint32_t)kSlabBytes, 128);
                            const uint64_t da_lo = ptx::make_smem_desc(a_lo, zpair ? (za + tap * 16 - a_lo) : (uint32_t)kSlabBytes, 128);
                            const uint32_t wb = w_base + (tap * 8 + 2 * kk) * (kB1WRows * 16);
                            const uint64_t db1 = ptx::make_smem_desc(wb, kB1WRows * 16, 128);               // this CTA's 64 rows of [W_hi ; W_lo]
                            const uint64_t db2 = ptx::make_smem_desc(wb + 64 * 16, kB1WRows * 16, 128);     // this CTA's 32 rows of W_hi
                            ptx::umma_bf16_ss_pair(d, da_hi, db1, idesc128, (tap | kk) ? 1u : 0u);
                            ptx::umma_bf16_ss_pair(d, da_lo, db2, idesc64, 1u);
                        }
                    }
                    ptx::umma_commit_pair(bar_a, (uint16_t)0x3);
                    ptx::umma_commit_pair(bar_b, (uint16_t)0x3);
                }
                __syncwarp();
            };
            auto issue_c1 = [&](int k) {
                const uint32_t buf = k & 1, ph = (k >> 1) & 1;
                ptx::mbar_wait(&x0_full[buf], ph);
                ptx::mbar_wait(&d1_empty[buf], ph ^ 1);
                ptx::mbar_wait_cluster(&p1_ready[buf], ph);
                ptx::tc_fence_after_sync();
                B1_TRACE(k, 4);
                conv_mmas(s0a + buf * kB1Slab0, 7, w1a, tmem_base + buf * 128, &x0_empty[buf], &d1_full[buf]);
            };
            auto issue_c2 = [&](int k) {
                const uint32_t buf = k & 1, ph = (k >> 1) & 1;
                ptx::mbar_wait(&x1_full[buf], ph);
                ptx::mbar_wait(&d2_empty[buf], ph ^ 1);
                ptx::mbar_wait_cluster(&p2_ready[buf], ph);
                ptx::tc_fence_after_sync();
                B1_TRACE(k, 5);
                conv_mmas(s1a + buf * kB1SlabBytes, 8, w2a, tmem_base + 256 + buf * 128, &x1_empty[buf], &d2_full[buf]);
            };
            if (warp == 12) { for (int k = 0; k < my_tiles; ++k) issue_c1(k); }
            else            { for (int k = 0; k < my_tiles; ++k) issue_c2(k); }
            }
        }
    } else {
        // ===== epilogue warps 4..11 =====
        const int q = warp & 3, h = (warp - 4) >> 2;          // TMEM lane quadrant, column half
        const int rit = q * 32 + lane;                        // MMA row this thread owns
        const int odd = lane & 1;
        const float* bias1 = s_bias + h * 32;
        const float* bias2 = s_bias + 64 + h * 32 + odd * 16;
        const uint32_t tq = tmem_base + h * 32 + ((uint32_t)(q * 32) << 16);

        auto epi1 = [&](int k) {
            const int tile = tile_of(k);
            const uint32_t buf = k & 1, ph = (k >> 1) & 1;
            const int r = tile * kB1Rows - 2 + rit;           // X1 row
            const bool valid = r >= 0 && pos_mod(r, kRW1) < 150;
            if (warp == 4) B1_TRACE(k, 6);
            ptx::mbar_wait_relaxed(&d1_full[buf], ph);
            if (warp == 4) B1_TRACE(k, 7);
            ptx::tc_fence_after_sync();
            uint32_t v[32], u[32];
            ptx::tmem_ld32(tq + buf * 128, v);                // a_hi w_hi + a_lo w_hi
            ptx::tmem_ld32(tq + buf * 128 + 64, u);           // a_hi w_lo
            ptx::tmem_ld_wait();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&d1_empty[buf]);  // accumulator is in registers now
            uint4 hi[4], lo[4];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float y[8];
                const float4 b0 = *reinterpret_cast<const float4*>(bias1 + qd * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(bias1 + qd * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float sum = __uint_as_float(v[qd * 8 + i]) + __uint_as_float(u[qd * 8 + i]);
                    y[i] = valid ? relu_nan(sum + bb[i]) : 0.f;
                }
                split8(y, hi[qd], lo[qd]);
            }
            if (warp == 4) B1_TRACE(k, 8);
            ptx::mbar_wait_relaxed(&x1_empty[buf], ph ^ 1);   // conv2 of tile k-2 has finished reading this slab1 buffer
            if (warp == 4) B1_TRACE(k, 9);
            if (p.dbg & 2) { ptx::mbar_arrive(&x1_full[buf]); return; }
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                uint8_t* d = slab1 + buf * kB1SlabBytes + (h * 4 + qd) * kSlabBytes + (rit + 1) * 16;
                *reinterpret_cast<uint4*>(d) = hi[qd];
                *reinterpret_cast<uint4*>(d + 8 * kSlabBytes) = lo[qd];
            }
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&x1_full[buf]);
            if (warp == 4) B1_TRACE(k, 10);
        };
        auto epi2 = [&](int k) {
            const int tile = tile_of(k);
            const uint32_t buf = k & 1, ph = (k >> 1) & 1;
            const int r = tile * kB1Rows - 2 + rit;           // conv2 output row (X1 row space)
            const bool valid = r >= 0 && (pos_mod(r, kRW1) >> 1) < 75;
            const bool store = rit >= 2 && rit < 126;         // pooled rows 0..61 of the tile
            if (warp == 4) B1_TRACE(k, 11);
            ptx::mbar_wait_relaxed(&d2_full[buf], ph);
            if (warp == 4) B1_TRACE(k, 12);
            ptx::tc_fence_after_sync();
            uint32_t v[32], u[32];
            ptx::tmem_ld32(tq + 256 + buf * 128, v);
            ptx::tmem_ld32(tq + 256 + buf * 128 + 64, u);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&d2_empty[buf]);
            if (p.dbg & 8) { ptx::mbar_arrive(st_full); return; }
            // MaxPool1d(2,2) first (bias and ReLU commute with max): the two lanes of a pool pair exchange halves, the even
            // lane finishes columns [0,16) of this warp's 32, the odd lane [16,32)
            float y[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float s0 = __uint_as_float(v[i]) + __uint_as_float(u[i]);
                const float s1 = __uint_as_float(v[16 + i]) + __uint_as_float(u[16 + i]);
                const float give = odd ? s0 : s1, keep = odd ? s1 : s0;
                y[i] = max_nan(keep, __shfl_xor_sync(0xffffffffu, give, 1));
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias2 + i);
                y[i] = valid ? relu_nan(y[i] + b4.x) : 0.f;
                y[i + 1] = valid ? relu_nan(y[i + 1] + b4.y) : 0.f;
                y[i + 2] = valid ? relu_nan(y[i + 2] + b4.z) : 0.f;
                y[i + 3] = valid ? relu_nan(y[i + 3] + b4.w) : 0.f;
            }
            uint4 ghi[2], glo[2];
#pragma unroll
            for (int qd = 0; qd < 2; ++qd) split8(y + qd * 8, ghi[qd], glo[qd]);
            ptx::mbar_wait_relaxed(st_done, (k & 1) ^ 1);     // the previous tile's rows have left the staging tile
            if (store) {
                uint8_t* base = stage + (h * 4 + odd * 2) * kB1StageRun + ((rit >> 1) - 1) * 16;      // [part][kchunk][pooled row][16 B]
#pragma unroll
                for (int qd = 0; qd < 2; ++qd) {
                    *reinterpret_cast<uint4*>(base + qd * kB1StageRun) = ghi[qd];
                    *reinterpret_cast<uint4*>(base + qd * kB1StageRun + 8 * kB1StageRun) = glo[qd];
                }
            }
            if (tile == 0 && rit < 2) {
                // the tape's leading guard row (X2 row -1: conv3's zero padding in front of window 0).  The tape's place in the
                // workspace depends on the chunk size, so the row may hold another call's data: written with every tile 0
                uint8_t* g = p.out + (size_t)(kGuard - 1) * 16 + (size_t)(h * 4 + odd * 2) * p.out_kch_stride;
#pragma unroll
                for (int qd = 0; qd < 2; ++qd) {
                    *reinterpret_cast<uint4*>(g + (size_t)qd * p.out_kch_stride) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(g + (size_t)qd * p.out_kch_stride + p.out_part_stride) = make_uint4(0, 0, 0, 0);
                }
            }
            ptx::fence_proxy_async_smem();                    // generic-proxy writes -> visible to the copy engine
            ptx::mbar_arrive(st_full);
            if (warp == 4) B1_TRACE(k, 13);
        };
        for (int k = 0; k < my_tiles; ++k) {
            epi1(k);
            if (k > 0) epi2(k - 1);
        }
        if (my_tiles > 0) epi2(my_tiles - 1);
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync();              // no CTA leaves (or frees its TMEM) while its peer's MMAs / arrives may still reach it
    if (warp == 12) ptx::tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace tc
}  // namespace dce
