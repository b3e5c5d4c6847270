// Fused block2 with a STACKED B operand: conv3 + ReLU -> conv4 + ReLU -> pool -> flatten, as dce_tc_block2.cuh, but the
// weight image of every (tap, kchunk) holds [W_hi (128 rows) ; W_lo (128 rows)] as ONE K-major operand, so a K-step is
//      MMA 1 (N = 256):  A_hi x [W_hi ; W_lo]  -> D[:, 0:128] += a_hi w_hi,  D[:, 128:256] += a_hi w_lo
//      MMA 2 (N = 128):  A_lo x  W_hi          -> D[:, 0:128] += a_lo w_hi
// instead of three N = 128 MMAs: the same tensor-pipe time (128 + 64 cycles), but 20 KB instead of 24 KB of shared-memory
// operand reads per K-step (A_hi is fetched once for both of its products; three N = 128 MMAs keep the 128 B/clk port
// 100 % busy, which starves the epilogue's own shared-memory traffic), and two issue slots instead of three, one of
// them twice as long (the pipe queues only an MMA or two ahead of the issuing thread).  The epilogues add the two
// accumulator halves.  Each accumulator is 256 TMEM columns, so D3 and D4 are single-buffered: the epilogues pull a
// whole accumulator into registers at once and hand it back before they do any arithmetic.
//      /root/reference/src/contact_cnn.py:28-44,64
#pragma once
#include "dce_tc_block2.cuh"

namespace dce {
namespace tc {

__global__ void __launch_bounds__(kB2Threads, 1)
block2s_kernel(const Block2Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // b3[128], b4[128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kB2Ring; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1); }
        ptx::mbar_init(d3_full, 1); ptx::mbar_init(d3_empty, 8);
        ptx::mbar_init(d4_full, 1); ptx::mbar_init(d4_empty, 8);
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
        constexpr uint32_t idesc256 = ptx::make_idesc_bf16_f32(128, 256);
        constexpr uint32_t idesc128 = ptx::make_idesc_bf16_f32(128, 128);
        const bool leader = ptx::elect_one();
        const uint32_t sa = ptx::smem_u32(slabA), sb = ptx::smem_u32(slabB), rg = ptx::smem_u32(ring);
        uint32_t it = 0;
        // one 24 KB weight block = 2 kchunks of K for all 3 taps, [tap][kchunk][W_hi 128 rows | W_lo 128 rows][8]: 6 MMAs.
        // The next block's barrier is probed in the middle of this block's MMAs.
        const uint32_t total_blocks = (uint32_t)my_tiles * 12;
        if (my_tiles > 0) { ptx::mbar_wait(&wfull[0], 0); ptx::tc_fence_after_sync(); }
        auto stage_mmas = [&](uint32_t slab, int kch_total, int s, uint32_t d, bool first_stage) {
            const uint32_t slot = it % kB2Ring;
            const uint32_t b0 = rg + slot * kB2WBlock;
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
                const uint32_t a_hi = slab + (uint32_t)(s * 2) * kSlabBytes + tap * 16;
                const uint64_t db = ptx::make_smem_desc(b0 + tap * 8192, 4096, 128);
                const uint64_t da_hi = ptx::make_smem_desc(a_hi, kSlabBytes, 128);
                const uint64_t da_lo = ptx::make_smem_desc(a_hi + (uint32_t)kch_total * kSlabBytes, kSlabBytes, 128);
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
            for (int s = 0; s < 4; ++s) stage_mmas(sa, 8, s, tmem_base, s == 0);
            if (leader) { ptx::umma_commit(a_empty); ptx::umma_commit(d3_full); }
        };
        auto issue_c4 = [&](int k) {
            B2_TRACE(k, 2);
            ptx::mbar_wait(x3_full, k & 1);
            ptx::mbar_wait(d4_empty, (k & 1) ^ 1);               // epilogue 2 of tile k-1 holds D4 in registers
            ptx::tc_fence_after_sync();
            B2_TRACE(k, 3);
            for (int s = 0; s < 8; ++s) stage_mmas(sb, 16, s, tmem_base + 256, s == 0);
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
        const int NR = p.n_windows * kRW2;
        const uint32_t tq = tmem_base + h * 64 + ((uint32_t)(q * 32) << 16);

        // epilogue 1, first half: D3 -> registers -> bias / ReLU / guard rows -> bf16 hi/lo images of this thread's row
        auto epi1_load = [&](int k, uint4 (&hi)[8], uint4 (&lo)[8]) {
            const int tile = blockIdx.x + k * gridDim.x;
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
            const int tile = blockIdx.x + k * gridDim.x;
            const int r = tile * kB2Rows - 2 + rit;            // conv4 output row (X3 row space)
            int w = 0, to = 0;
            bool store = r >= 0 && r < NR && rit >= 2 && rit < 126 && !(p.dbg & 2);
            if (store) { w = r / kRW2; to = (r - w * kRW2) >> 1; store = to < 37 && w < p.out_rows_cap; }
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
            if (p.dbg & 16) return;                            // ablation: epilogue 2 ends with the TMEM load
            float m[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float s0 = __uint_as_float(va[i]) + __uint_as_float(vc[i]);     // columns h*64 + i
                const float s1 = __uint_as_float(vb[i]) + __uint_as_float(vd[i]);     // columns h*64 + 32 + i
                const float give = odd ? s0 : s1, keep = odd ? s1 : s0;               // even lanes finish the low 32, odd the high 32
                m[i] = max_nan(keep, __shfl_xor_sync(0xffffffffu, give, 1));          // MaxPool1d(2,2)
            }
            if (warp == 0) B2_TRACE(k, 14);
            uint8_t* base = p.out + (size_t)(to * 16 + h * 8 + odd * 4) * p.out_kch_stride + (size_t)(w + kGuard) * 16;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float y[8];
                const float4 b0 = *reinterpret_cast<const float4*>(bias4 + qd * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(bias4 + qd * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = relu_nan(m[qd * 8 + i] + bb[i]);
                uint4 hi, lo;
                split8(y, hi, lo);
                if (qd == 0 && warp == 0) B2_TRACE(k, 15);
                if (store) {
                    *reinterpret_cast<uint4*>(base + (size_t)qd * p.out_kch_stride) = hi;
                    *reinterpret_cast<uint4*>(base + (size_t)qd * p.out_kch_stride + p.out_part_stride) = lo;
                }
            }
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
