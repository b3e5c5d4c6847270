// Fused block2: Conv1d 64->128 + ReLU -> Conv1d 128->128 + ReLU -> MaxPool1d(2,2) -> flatten, one persistent
// kernel; the 128-channel activation between the two convolutions (X3: 159 MB per 4096 windows as bf16 hi/lo,
// the largest tensor of the whole path) never leaves the SM.   /root/reference/src/contact_cnn.py:28-44,64
//
// Same tiling idea as block1 (dce_tc_block1.cuh): a tile is 124 rows of conv4 output; with b = 124*i
//      slabA row s  <->  X2 row b-3+s  (s = 0..129)   bulk-TMA straight from the X2 tape (16 x 2080 B)
//      conv3 MMA row k <-> X3 row b-2+k, reads slabA rows k..k+2; result -> slabB row k+1 (smem, 128 channels)
//      conv4 MMA row j <-> out row b-2+j, reads slabB rows j..j+2; rows j in [2,126) are exact
// The two weight images (96 KB + 192 KB) cannot stay resident next to a 66 KB slabB, so they stream from L2
// through a 4-stage ring of 24 KB blocks in the fixed order conv3(k+1), conv4(k), ...
//
// Arithmetic per K-step (stacked B operand): the weight image of every (tap, kchunk) holds [W_hi (128 rows) ; W_lo (128 rows)]
// as ONE K-major operand, so a K-step is
//      MMA 1 (N = 256):  A_hi x [W_hi ; W_lo]  -> D[:, 0:128] += a_hi w_hi,  D[:, 128:256] += a_hi w_lo
//      MMA 2 (N = 128):  A_lo x  W_hi          -> D[:, 0:128] += a_lo w_hi
// instead of three N = 128 MMAs: the same tensor-pipe time (128 + 64 cycles), 20 KB instead of 24 KB of shared-memory
// operand reads, two issue slots instead of three.  The epilogues add the two accumulator halves.  Each accumulator is
// 256 TMEM columns, so D3 and D4 are single-buffered: an epilogue pulls the whole accumulator into registers and hands
// it back before it does any arithmetic.
//
// Output: the fc.0 operand is ROW-MAJOR ([part][window][k' = t * 128 + c] bf16), so a pooled row is 256 contiguous bytes of
// hi and of lo and the pooled rows a tile holds of one window are one contiguous run.  Epilogue 2 writes its rows into a
// dense shared-memory staging tile ([part][62 rows][256 B]; every lane rotates the order of its four 16-byte granules by
// its row number, which keeps the writes conflict-free without padding) and a store warp hands every (window, part) run
// to the copy engine as ONE bulk copy: at most six per tile (the engine spends ~40 cycles per copy whatever its size).  (The first version of this kernel wrote a [kchunk][window][8] tape: 1 984
// isolated 16-byte granules per tile.  The LSU takes ~4.4 cycles per such granule whoever issues it — 8 700 cycles per
// tile, more than the tile's MMAs — and the TMA engine's tensor-store scatter is no faster: profiles/r02_stacked_operand.md.)
//
// Warps (12, 1 CTA/SM): 0-7 epilogue (e1: D3 -> slabB; e2: D4 -> pooled rows -> staging), 8 MMA issuer,
// 9 weight producer, 10 slabA loader, 11 store warp.
#pragma once
#include "dce_tc.cuh"
#include "dce_tc_block1.cuh"

namespace dce {
namespace tc {

constexpr int kB2Rows = 124;
constexpr int kB2Threads = 12 * 32;
constexpr int kB2SlabA = 2 * 8 * kSlabBytes;        // 33280: [part][8 kchunks][130][16 B]
constexpr int kB2SlabB = 2 * 16 * kSlabBytes;       // 66560: [part][16 kchunks][130][16 B]
constexpr int kB2WBlock = 24576;                    // [tap][2 kchunks][W_hi 128 rows | W_lo 128 rows][8] bf16: 6 MMAs
constexpr int kB2Ring = 4;                          // weight ring stages
constexpr int kB2StagePart = 62 * 256;              // staging: [part][62 pooled rows][128 channels] bf16, dense
constexpr int kB2Stage = 2 * kB2StagePart;
constexpr int kB2SmemBytes = kB2SlabA + kB2SlabB + kB2Ring * kB2WBlock + 256 + 2 * 128 * 4 + kB2Stage;
static_assert(kB2SmemBytes <= 232448, "exceeds 227 KB");

struct Block2Params {
    const uint8_t* x2; size_t x2_part_stride, x2_kch_stride;
    int n_windows;
    const uint8_t* w3; const uint8_t* w4;           // 4 and 8 blocks of kB2WBlock
    const float* b3; const float* b4;
    uint8_t* out; size_t out_part_stride; int out_rows_cap;     // row-major fc.0 operand: part 0 (hi), rows of 4736 bf16
    int n_tiles;
    long long* trace;            // optional clock64 timeline of CTA 0 (DCE_TRACE builds)
    int dbg;                     // timing ablations (results invalid): 1 = no weight copies once the ring is primed; 2 = no output stores;
                                 // 4 = slabA is loaded for the first tile only; 16 = epilogue 2 stops after its TMEM load; 32 = epilogue 1 writes nothing
};

#define B2_TRACE(k, ev) do { if (DCE_TRACE && p.trace && blockIdx.x == 0 && (k) < 60 && (threadIdx.x & 31) == 0) p.trace[(k) * 16 + (ev)] = clock64(); } while (0)

__global__ void __launch_bounds__(kB2Threads, 1)
block2_kernel(const Block2Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* slabA = smem;
    uint8_t* slabB = smem + kB2SlabA;
    uint8_t* ring = slabB + kB2SlabB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kB2Ring * kB2WBlock);
    uint64_t* wfull = bars;          // [4] weight block landed (tx bytes)
    uint64_t* wempty = bars + 4;     // [4] tcgen05.commit
    uint64_t* a_full = bars + 8;     // slabA landed
    uint64_t* a_empty = bars + 9;    // conv3 has finished reading slabA
    uint64_t* d3_full = bars + 10;
    uint64_t* d3_empty = bars + 11;  // 8 epilogue warps hold D3 in registers
    uint64_t* x3_full = bars + 12;   // 256 epilogue threads wrote slabB
    uint64_t* x3_empty = bars + 13;  // conv4 has finished reading slabB
    uint64_t* d4_full = bars + 14;
    uint64_t* d4_empty = bars + 15;
    uint64_t* st_full = bars + 16;   // 256 epilogue threads wrote the staging tile
    uint64_t* st_empty = bars + 17;  // the copy engine has read it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // b3[128], b4[128]
    uint8_t* stage = reinterpret_cast<uint8_t*>(s_bias + 256);                          // kB2Stage bytes

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // Tiles are walked from the END of the X2 tape: block1 wrote it front to back, so its tail is what the L2 still holds
    // (an LRU cache scanned in writing order by a reader that trails the writer by more than its capacity hits nothing).
    auto tile_of = [&](int k) { const int t = (int)blockIdx.x + k * (int)gridDim.x; return (p.dbg & 64) ? t : p.n_tiles - 1 - t; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < kB2Ring; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1); }
        ptx::mbar_init(d3_full, 1); ptx::mbar_init(d3_empty, 8);
        ptx::mbar_init(d4_full, 1); ptx::mbar_init(d4_empty, 8);
        ptx::mbar_init(a_full, 1); ptx::mbar_init(a_empty, 1);
        ptx::mbar_init(x3_full, 256); ptx::mbar_init(x3_empty, 1);
        ptx::mbar_init(st_full, 256); ptx::mbar_init(st_empty, 1);
        ptx::fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 8) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    for (int i = threadIdx.x; i < kB2SlabB / 16; i += kB2Threads)            // halo rows of slabB are never written
        reinterpret_cast<uint4*>(slabB)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 256) s_bias[threadIdx.x] = __ldg((threadIdx.x < 128 ? p.b3 : p.b4 - 128) + threadIdx.x);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // the weight producer reads only the packed weights, which no kernel of the step writes: it fills the ring while the
    // previous kernel drains; everyone else waits for the X2 tape
    if (warp != 9) pdl_wait();

    if (warp == 9) {
        // ===== weight producer: blocks in the order the issuer consumes them: c3(0); then c3(k+1), c4(k) =====
        uint32_t it = 0;
        auto stream_blocks = [&](const uint8_t* w, int nblk) {
            for (int s = 0; s < nblk; ++s, ++it) {
                const uint32_t slot = it % kB2Ring, ph = (it / kB2Ring) & 1;
                ptx::mbar_wait_relaxed(&wempty[slot], ph ^ 1);
                if (ptx::elect_one()) {
                    if ((p.dbg & 1) && it >= kB2Ring) ptx::mbar_arrive(&wfull[slot]);
                    else {
                        ptx::mbar_arrive_expect_tx(&wfull[slot], kB2WBlock);
                        ptx::bulk_g2s(ring + slot * kB2WBlock, w + (size_t)s * kB2WBlock, kB2WBlock, &wfull[slot]);
                    }
                }
                __syncwarp();
            }
        };
        if (my_tiles > 0) stream_blocks(p.w3, 4);
        for (int k = 0; k < my_tiles; ++k) {
            if (k + 1 < my_tiles) stream_blocks(p.w3, 4);
            stream_blocks(p.w4, 8);
        }
    } else if (warp == 10) {
        // ===== slabA loader: 130 rows x 8 kchunks x hi/lo of the X2 tape per tile =====
        for (int k = 0; k < my_tiles; ++k) {
            const int b = tile_of(k) * kB2Rows;
            ptx::mbar_wait_relaxed(a_empty, (k & 1) ^ 1);                       // conv3(k-1) has drained slabA
            if ((p.dbg & 4) && k > 0) {
                if (ptx::elect_one()) ptx::mbar_arrive(a_full);
                __syncwarp();
                continue;
            }
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(a_full, kB2SlabA);
                const uint8_t* src = p.x2 + (size_t)(b - 3 + kGuard) * 16;
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    ptx::bulk_g2s(slabA + c * kSlabBytes, src + (c >> 3) * p.x2_part_stride + (size_t)(c & 7) * p.x2_kch_stride,
                                  kSlabBytes, a_full);
            }
            __syncwarp();
        }
    } else if (warp == 11) {
        // ===== store warp: staging tile -> row-major fc.0 operand.  A tile spans at most three windows; the pooled rows it
        // holds of one window are contiguous in the staging tile and in global memory: lane l < 6 copies (window l / 2, part l % 2)
        const int NR = p.n_windows * kRW2;
        for (int k = 0; k < my_tiles; ++k) {
            const int tile = tile_of(k);
            ptx::mbar_wait_relaxed(st_full, k & 1);
            if (!(p.dbg & 2) && lane < 6) {
                const int r0 = tile * kB2Rows;                         // conv4 row of pooled row 0 (rit = 2)
                const int w = r0 / kRW2 + (lane >> 1), part = lane & 1;
                int lo = w * kRW2 > r0 ? w * kRW2 : r0;                // rows [lo, hi) of window w in this tile, pooled rows 0..36 only
                int hi = w * kRW2 + 74 < r0 + kB2Rows ? w * kRW2 + 74 : r0 + kB2Rows;
                if (hi > NR) hi = NR;
                if (hi > lo && w < p.out_rows_cap) {
                    const int pr = (lo - r0) >> 1, to = (lo - w * kRW2) >> 1, n = (hi - lo) >> 1;
                    ptx::bulk_s2g(p.out + (part ? p.out_part_stride : 0) + ((size_t)w * 4736 + (size_t)to * 128) * 2,
                                  stage + part * kB2StagePart + pr * 256, (uint32_t)n * 256);
                }
                ptx::bulk_commit_group();
                ptx::bulk_wait_group_read0();                          // the staging tile has been read
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(st_empty);
        }
        if (lane < 6) ptx::bulk_wait_group0();                         // every write has been performed before the CTA exits
    } else if (warp == 8) {
        // ===== MMA issuer (leader elected once) =====
        constexpr uint32_t idesc256 = ptx::make_idesc_bf16_f32(128, 256);
        constexpr uint32_t idesc128 = ptx::make_idesc_bf16_f32(128, 128);
        const bool leader = ptx::elect_one();
        const uint32_t sa = ptx::smem_u32(slabA), sb = ptx::smem_u32(slabB), rg = ptx::smem_u32(ring);
        uint32_t it = 0;
        // one 24 KB weight block = 2 kchunks of K for all 3 taps, [tap][kchunk][W_hi 128 rows | W_lo 128 rows][8]: 6 MMAs.
        // The next block's barrier is probed in the middle of this block's MMAs.
        const uint32_t total_blocks = (uint32_t)my_tiles * 12;
        if (my_tiles > 0) { ptx::mbar_wait(&wfull[0], 0); ptx::tc_fence_after_sync(); }
        // Descriptors are built by ADDING 16-byte units to the low word of a constant descriptor (start addresses stay below
        // 2^18, so no carry reaches the LBO field at bit 16): one uniform add per descriptor instead of shift / mask / or — the
        // thread that issues the MMAs is the kernel's critical path, the pipe queues only an MMA or two ahead of it
        const uint64_t kDA = ptx::make_smem_desc(0, kSlabBytes, 128), kDB = ptx::make_smem_desc(0, 4096, 128);
        const uint32_t sa16 = sa >> 4, sb16 = sb >> 4, rg16 = rg >> 4;
        auto stage_mmas = [&](uint32_t slab16, int kch_total, int s, uint32_t d, bool first_stage) {
            const uint32_t slot = it % kB2Ring;
            const uint32_t b16 = rg16 + slot * (kB2WBlock >> 4);
            const uint32_t a16 = slab16 + (uint32_t)(s * 2) * (kSlabBytes >> 4);
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
                const uint64_t db = kDB + (uint64_t)(b16 + tap * (8192 >> 4));
                const uint64_t da_hi = kDA + (uint64_t)(a16 + tap);
                const uint64_t da_lo = kDA + (uint64_t)(a16 + tap + (uint32_t)kch_total * (kSlabBytes >> 4));
                if (leader) {
                    ptx::umma_bf16_ss(d, da_hi, db, idesc256, (first_stage && tap == 0) ? 0u : 1u);
                    ptx::umma_bf16_ss(d, da_lo, db, idesc128, 1u);
                }
                if (tap == 1 && it + 1 < total_blocks) {          // probe the next block while this one's MMAs are queued
                    ptx::mbar_wait(&wfull[(it + 1) % kB2Ring], ((it + 1) / kB2Ring) & 1);
                    ptx::tc_fence_after_sync();
                }
            }
            if (leader) ptx::umma_commit(&wempty[slot]);
            ++it;
        };
        auto issue_c3 = [&](int k) {
            B2_TRACE(k, 0);
            ptx::mbar_wait(a_full, k & 1);
            ptx::mbar_wait(d3_empty, (k & 1) ^ 1);               // epilogue 1 of tile k-1 holds D3 in registers
            ptx::tc_fence_after_sync();
            B2_TRACE(k, 1);
            for (int s = 0; s < 4; ++s) stage_mmas(sa16, 8, s, tmem_base, s == 0);
            if (leader) { ptx::umma_commit(a_empty); ptx::umma_commit(d3_full); }
        };
        auto issue_c4 = [&](int k) {
            B2_TRACE(k, 2);
            ptx::mbar_wait(x3_full, k & 1);
            ptx::mbar_wait(d4_empty, (k & 1) ^ 1);               // epilogue 2 of tile k-1 holds D4 in registers
            ptx::tc_fence_after_sync();
            B2_TRACE(k, 3);
            for (int s = 0; s < 8; ++s) stage_mmas(sb16, 16, s, tmem_base + 256, s == 0);
            if (leader) { ptx::umma_commit(x3_empty); ptx::umma_commit(d4_full); }
            B2_TRACE(k, 4);
        };
        if (my_tiles > 0) issue_c3(0);
        for (int k = 0; k < my_tiles; ++k) {
            if (k + 1 < my_tiles) issue_c3(k + 1);
            issue_c4(k);
        }
    } else {
        // ===== epilogue warps 0..7 =====
        const int q = warp & 3, h = warp >> 2;                // TMEM lane quadrant, 64-column half
        const int rit = q * 32 + lane;
        const int odd = lane & 1;
        const float* bias3 = s_bias + h * 64;
        const float* bias4 = s_bias + 128 + h * 64 + odd * 32;
        const uint32_t tq = tmem_base + h * 64 + ((uint32_t)(q * 32) << 16);

        // epilogue 1, first half: D3 -> registers -> bias / ReLU / guard rows -> bf16 hi/lo images of this thread's row
        auto epi1_load = [&](int k, uint4 (&hi)[8], uint4 (&lo)[8]) {
            const int tile = tile_of(k);
            const int r = tile * kB2Rows - 2 + rit;            // X3 row
            const bool valid = r >= 0 && pos_mod(r, kRW2) < 75;
            if (warp == 0) B2_TRACE(k, 5);
            ptx::mbar_wait_relaxed(d3_full, k & 1);
            if (warp == 0) B2_TRACE(k, 6);
            ptx::tc_fence_after_sync();
            uint32_t va[32], vb[32], vc[32], vd[32];
            ptx::tmem_ld32(tq, va);
            ptx::tmem_ld32(tq + 32, vb);
            ptx::tmem_ld32(tq + 128, vc);
            ptx::tmem_ld32(tq + 160, vd);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(d3_empty);         // the accumulator is in registers now
            if (warp == 0) B2_TRACE(k, 7);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t (&u)[32] = c ? vb : va;
                const uint32_t (&w)[32] = c ? vd : vc;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    float y[8];
                    const float4 b0 = *reinterpret_cast<const float4*>(bias3 + c * 32 + qd * 8);
                    const float4 b1 = *reinterpret_cast<const float4*>(bias3 + c * 32 + qd * 8 + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float s = __uint_as_float(u[qd * 8 + i]) + __uint_as_float(w[qd * 8 + i]);
                        y[i] = valid ? relu_nan(s + bb[i]) : 0.f;
                    }
                    split8(y, hi[c * 4 + qd], lo[c * 4 + qd]);
                }
            }
        };
        // second half: once conv4 of the previous tile has finished reading slabB, 16 stores
        auto epi1_store = [&](int k, const uint4 (&hi)[8], const uint4 (&lo)[8]) {
            ptx::mbar_wait(x3_empty, (k & 1) ^ 1);             // critical path: tight poll
            if (warp == 0) B2_TRACE(k, 8);
            if (!(p.dbg & 32)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint8_t* d = slabB + (h * 8 + j) * kSlabBytes + (rit + 1) * 16;
                    *reinterpret_cast<uint4*>(d) = hi[j];
                    *reinterpret_cast<uint4*>(d + 16 * kSlabBytes) = lo[j];
                }
            }
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(x3_full);
            if (warp == 0) B2_TRACE(k, 9);
        };
        // epilogue 2: D4 -> registers (all 64 + 64 columns at once, then the accumulator is free) -> pool: the two lanes
        // of a pool pair exchange halves, so every lane finishes 32 pooled columns -> bias / ReLU (both commute with
        // max) -> bf16 hi/lo -> fc.0 operand tape
        auto epi2 = [&](int k) {
            const bool store = rit >= 2 && rit < 126;         // the store warp drops rows past the end of the chunk / pooled rows >= 37
            if (warp == 0) B2_TRACE(k, 10);
            ptx::mbar_wait_relaxed(d4_full, k & 1);
            if (warp == 0) B2_TRACE(k, 11);
            ptx::tc_fence_after_sync();
            uint32_t va[32], vb[32], vc[32], vd[32];
            ptx::tmem_ld32(tq + 256, va);
            ptx::tmem_ld32(tq + 256 + 32, vb);
            ptx::tmem_ld32(tq + 256 + 128, vc);
            ptx::tmem_ld32(tq + 256 + 160, vd);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(d4_empty);
            if (warp == 0) B2_TRACE(k, 13);
            if (p.dbg & 16) { ptx::mbar_arrive(st_full); return; }   // ablation: epilogue 2 ends with the TMEM load
            float m[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float s0 = __uint_as_float(va[i]) + __uint_as_float(vc[i]);     // columns h*64 + i
                const float s1 = __uint_as_float(vb[i]) + __uint_as_float(vd[i]);     // columns h*64 + 32 + i
                const float give = odd ? s0 : s1, keep = odd ? s1 : s0;               // even lanes finish the low 32, odd the high 32
                m[i] = max_nan(keep, __shfl_xor_sync(0xffffffffu, give, 1));          // MaxPool1d(2,2)
            }
            if (warp == 0) B2_TRACE(k, 14);
            const int prl = (rit >> 1) - 1;                    // pooled row of the tile
            uint8_t* base = stage + prl * 256 + (h * 8 + odd * 4) * 16;     // channels h*64 + odd*32 ..
            uint4 ghi[4], glo[4];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float y[8];
                const float4 b0 = *reinterpret_cast<const float4*>(bias4 + qd * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(bias4 + qd * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = relu_nan(m[qd * 8 + i] + bb[i]);
                split8(y, ghi[qd], glo[qd]);
            }
            if (warp == 0) B2_TRACE(k, 15);
            ptx::mbar_wait_relaxed(st_empty, (k & 1) ^ 1);     // the previous tile's rows have left the staging tile
            if (store) {
                // instruction j writes granule (j + row) % 4: the 16 rows of a warp then cover all 32 banks four times (the
                // optimum for 512 bytes) although consecutive rows are 256 B apart
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = (j + prl) & 3;
                    const uint4 vh = c == 0 ? ghi[0] : c == 1 ? ghi[1] : c == 2 ? ghi[2] : ghi[3];
                    const uint4 vl = c == 0 ? glo[0] : c == 1 ? glo[1] : c == 2 ? glo[2] : glo[3];
                    *reinterpret_cast<uint4*>(base + c * 16) = vh;
                    *reinterpret_cast<uint4*>(base + kB2StagePart + c * 16) = vl;
                }
            }
            ptx::fence_proxy_async_smem();                     // generic-proxy writes -> visible to the copy engine
            ptx::mbar_arrive(st_full);
        };
        uint4 hi[8], lo[8];
        for (int k = 0; k < my_tiles; ++k) {
            epi1_load(k, hi, lo);
            epi1_store(k, hi, lo);
            if (k > 0) { epi2(k - 1); if (warp == 0) B2_TRACE(k - 1, 12); }
        }
        if (my_tiles > 0) epi2(my_tiles - 1);
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace dce
