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
// Warps (11, 1 CTA/SM): 0-7 epilogue (e1: D3 -> slabB; e2: D4 -> pooled fc.0 operand tape), 8 MMA issuer,
// 9 weight producer, 10 slabA loader.
#pragma once
#include "dce_tc.cuh"
#include "dce_tc_block1.cuh"

namespace dce {
namespace tc {

constexpr int kB2Rows = 124;
constexpr int kB2Threads = 11 * 32;
constexpr int kB2SlabA = 2 * 8 * kSlabBytes;        // 33280: [part][8 kchunks][130][16 B]
constexpr int kB2SlabB = 2 * 16 * kSlabBytes;       // 66560: [part][16 kchunks][130][16 B]
constexpr int kB2WBlock = 24576;                    // [part][tap][2 kchunks][128][8] bf16: 9 MMAs
constexpr int kB2Ring = 4;                          // weight ring stages
constexpr int kB2SmemBytes = kB2SlabA + kB2SlabB + kB2Ring * kB2WBlock + 256 + 2 * 128 * 4;

struct Block2Params {
    const uint8_t* x2; size_t x2_part_stride, x2_kch_stride;
    int n_windows;
    const uint8_t* w3; const uint8_t* w4;           // 4 and 8 blocks of kB2WBlock
    const float* b3; const float* b4;
    uint8_t* out; size_t out_part_stride, out_kch_stride; int out_rows_cap;
    int n_tiles;
    long long* trace;            // optional clock64 timeline of CTA 0 (DCE_TRACE builds)
    int dbg;                     // timing ablations (results invalid): 1 = no weight copies once the ring is primed; 2 = no output stores;
                                 // 4 = slabA is loaded for the first tile only; 16 = epilogue 2 stops after its TMEM load; 32 = epilogue 1 writes nothing
};

#define B2_TRACE(k, ev) do { if (DCE_TRACE && p.trace && blockIdx.x == 0 && (k) < 60 && (threadIdx.x & 31) == 0) p.trace[(k) * 16 + (ev)] = clock64(); } while (0)

__global__ void __launch_bounds__(kB2Threads, 1)
block2_kernel(const Block2Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* slabA = smem;
    uint8_t* slabB = smem + kB2SlabA;
    uint8_t* ring = slabB + kB2SlabB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kB2Ring * kB2WBlock);
    uint64_t* wfull = bars;          // [4] weight block landed (tx bytes)
    uint64_t* wempty = bars + 4;     // [4] tcgen05.commit
    uint64_t* a_full = bars + 8;     // slabA landed
    uint64_t* a_empty = bars + 9;    // conv3 has finished reading slabA
    uint64_t* d3_full = bars + 10;   // [2]
    uint64_t* d3_empty = bars + 12;  // [2] 8 epilogue warps
    uint64_t* x3_full = bars + 14;   // 256 epilogue threads wrote slabB
    uint64_t* x3_empty = bars + 15;  // conv4 has finished reading slabB
    uint64_t* d4_full = bars + 16;   // [2]
    uint64_t* d4_empty = bars + 18;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // b3[128], b4[128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kB2Ring; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&d3_full[i], 1); ptx::mbar_init(&d3_empty[i], 8);
            ptx::mbar_init(&d4_full[i], 1); ptx::mbar_init(&d4_empty[i], 8);
        }
        ptx::mbar_init(a_full, 1); ptx::mbar_init(a_empty, 1);
        ptx::mbar_init(x3_full, 256); ptx::mbar_init(x3_empty, 1);
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
    // everything the previous kernel wrote (the X2 tape) is visible after this; the weight producer reads only the packed
    // weights, which no kernel of the step writes: it fills the ring while the previous kernel drains
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
            const int b = (int)(blockIdx.x + k * gridDim.x) * kB2Rows;
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
    } else if (warp == 8) {
        // ===== MMA issuer (leader elected once) =====
        constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128);
        const bool leader = ptx::elect_one();
        const uint32_t sa = ptx::smem_u32(slabA), sb = ptx::smem_u32(slabB), rg = ptx::smem_u32(ring);
        uint32_t it = 0;
        // one 24 KB weight block = 2 kchunks of K for all 3 taps: 9 MMAs.  The block's barrier was probed in the middle
        // of the previous block's MMAs (the pipe queues only an MMA or two ahead of this thread).
        const uint32_t total_blocks = (uint32_t)my_tiles * 12;
        if (my_tiles > 0) { ptx::mbar_wait(&wfull[0], 0); ptx::tc_fence_after_sync(); }
        auto stage_mmas = [&](uint32_t slab, int kch_total, int s, uint32_t d, bool first_stage) {
            const uint32_t slot = it % kB2Ring;
            const uint32_t b0 = rg + slot * kB2WBlock;
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
                const uint32_t b_hi = b0 + tap * 2 * 2048;
                const uint32_t a_hi = slab + (uint32_t)(s * 2) * kSlabBytes + tap * 16;
                const uint64_t db_hi = ptx::make_smem_desc(b_hi, 2048, 128);
                const uint64_t db_lo = ptx::make_smem_desc(b_hi + 12288, 2048, 128);
                const uint64_t da_hi = ptx::make_smem_desc(a_hi, kSlabBytes, 128);
                const uint64_t da_lo = ptx::make_smem_desc(a_hi + (uint32_t)kch_total * kSlabBytes, kSlabBytes, 128);
                if (leader) {
                    ptx::umma_bf16_ss(d, da_hi, db_lo, idesc, (first_stage && tap == 0) ? 0u : 1u);
                    ptx::umma_bf16_ss(d, da_lo, db_hi, idesc, 1u);
                    ptx::umma_bf16_ss(d, da_hi, db_hi, idesc, 1u);
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
            const uint32_t buf = k & 1, ph = (k >> 1) & 1;
            B2_TRACE(k, 0);
            ptx::mbar_wait(a_full, k & 1);
            ptx::mbar_wait(&d3_empty[buf], ph ^ 1);
            ptx::tc_fence_after_sync();
            B2_TRACE(k, 1);
            for (int s = 0; s < 4; ++s) stage_mmas(sa, 8, s, tmem_base + buf * 128, s == 0);
            if (leader) { ptx::umma_commit(a_empty); ptx::umma_commit(&d3_full[buf]); }
        };
        auto issue_c4 = [&](int k) {
            const uint32_t buf = k & 1, ph = (k >> 1) & 1;
            B2_TRACE(k, 2);
            ptx::mbar_wait(x3_full, k & 1);
            ptx::mbar_wait(&d4_empty[buf], ph ^ 1);
            ptx::tc_fence_after_sync();
            B2_TRACE(k, 3);
            for (int s = 0; s < 8; ++s) stage_mmas(sb, 16, s, tmem_base + 256 + buf * 128, s == 0);
            if (leader) { ptx::umma_commit(x3_empty); ptx::umma_commit(&d4_full[buf]); }
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
        const float* bias3 = s_bias + h * 64;
        const float* bias4 = s_bias + 128 + h * 64;
        const int NR = p.n_windows * kRW2;

        auto epi1 = [&](int k) {
            const int tile = blockIdx.x + k * gridDim.x;
            const uint32_t buf = k & 1, ph = (k >> 1) & 1;
            const int r = tile * kB2Rows - 2 + rit;            // X3 row
            const bool valid = r >= 0 && pos_mod(r, kRW2) < 75;
            if (warp == 0) B2_TRACE(k, 5);
            ptx::mbar_wait_relaxed(&d3_full[buf], ph);
            if (warp == 0) B2_TRACE(k, 6);
            ptx::tc_fence_after_sync();
            uint32_t va[32], vb[32];
            const uint32_t ta = tmem_base + buf * 128 + h * 64 + ((uint32_t)(q * 32) << 16);
            ptx::tmem_ld32(ta, va);
            ptx::tmem_ld32(ta + 32, vb);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&d3_empty[buf]);   // accumulator is in registers now
            if (warp == 0) B2_TRACE(k, 7);
            ptx::mbar_wait(x3_empty, (k & 1) ^ 1);             // conv4 of the previous tile has finished reading slabB (critical path: tight poll)
            if (warp == 0) B2_TRACE(k, 8);
            if (p.dbg & 32) { ptx::mbar_arrive(x3_full); return; }   // ablation: epilogue 1 writes nothing
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t (&v)[32] = c ? vb : va;
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias3 + c * 32 + i);
                    y[i] = valid ? relu_nan(__uint_as_float(v[i]) + b4.x) : 0.f;
                    y[i + 1] = valid ? relu_nan(__uint_as_float(v[i + 1]) + b4.y) : 0.f;
                    y[i + 2] = valid ? relu_nan(__uint_as_float(v[i + 2]) + b4.z) : 0.f;
                    y[i + 3] = valid ? relu_nan(__uint_as_float(v[i + 3]) + b4.w) : 0.f;
                }
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 hi, lo;
                    split8(y + qd * 8, hi, lo);
                    uint8_t* d = slabB + (h * 8 + c * 4 + qd) * kSlabBytes + (rit + 1) * 16;
                    *reinterpret_cast<uint4*>(d) = hi;
                    *reinterpret_cast<uint4*>(d + 16 * kSlabBytes) = lo;
                }
            }
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(x3_full);
            if (warp == 0) B2_TRACE(k, 9);
        };
        auto epi2 = [&](int k) {
            const int tile = blockIdx.x + k * gridDim.x;
            const uint32_t buf = k & 1, ph = (k >> 1) & 1;
            const int r = tile * kB2Rows - 2 + rit;            // conv4 output row (X3 row space)
            int w = 0, to = 0;
            bool store = r >= 0 && r < NR && rit >= 2 && rit < 126 && !(p.dbg & 2);
            if (store) { w = r / kRW2; to = (r - w * kRW2) >> 1; store = to < 37 && w < p.out_rows_cap; }
            if (warp == 0) B2_TRACE(k, 10);
            ptx::mbar_wait_relaxed(&d4_full[buf], ph);
            if (warp == 0) B2_TRACE(k, 11);
            ptx::tc_fence_after_sync();
            uint32_t va[32], vb[32];
            const uint32_t ta = tmem_base + 256 + buf * 128 + h * 64 + ((uint32_t)(q * 32) << 16);
            ptx::tmem_ld32(ta, va);
            ptx::tmem_ld32(ta + 32, vb);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&d4_empty[buf]);
            if (p.dbg & 16) return;                            // ablation: epilogue 2 ends with the TMEM load
            uint8_t* base = p.out + (size_t)(to * 16) * p.out_kch_stride + (size_t)(w + kGuard) * 16 + ((lane & 1) ? p.out_part_stride : 0);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t (&v)[32] = c ? vb : va;
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias4 + c * 32 + i);
                    y[i] = relu_nan(__uint_as_float(v[i]) + b4.x);
                    y[i + 1] = relu_nan(__uint_as_float(v[i + 1]) + b4.y);
                    y[i + 2] = relu_nan(__uint_as_float(v[i + 2]) + b4.z);
                    y[i + 3] = relu_nan(__uint_as_float(v[i + 3]) + b4.w);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = max_nan(y[i], __shfl_xor_sync(0xffffffffu, y[i], 1));   // MaxPool1d(2,2)
                if (store) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        uint4 hi, lo;
                        split8(y + qd * 8, hi, lo);
                        *reinterpret_cast<uint4*>(base + (size_t)(h * 8 + c * 4 + qd) * p.out_kch_stride) = (lane & 1) ? lo : hi;
                    }
                }
            }
        };
        for (int k = 0; k < my_tiles; ++k) {
            epi1(k);
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
