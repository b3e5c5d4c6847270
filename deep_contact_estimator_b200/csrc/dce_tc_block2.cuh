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
    const float* inv_sw3; const float* inv_sw4;     // F8IN: 1 / (power-of-two weight scale) of conv3 / conv4
    unsigned int* f8_status;                        // F8OUT / F8IN: range diagnostic word (f8_range_note), or nullptr
};

#define B2_TRACE(k, ev) do { if (DCE_TRACE && p.trace && blockIdx.x == 0 && (k) < 60 && (threadIdx.x & 31) == 0) p.trace[(k) * 16 + (ev)] = clock64(); } while (0)

// F8OUT: write the fc.0 operand in the fp16 + e4m3 format (dce_tc.cuh: split16_f16f8) instead of bf16 hi/lo.
// F8IN : X2, X3 (slabB) and both weight images are in that format too (option "conv_f16f8").  A slab of C
//        channels then holds C/8 fp16 chunks, C/16 lo8 chunks and C/16 hi8 chunks (the same 16 bytes per row and
//        chunk, the same slab bytes), a 24 KB weight block covers 32 input channels of all three taps — first the
//        conv's e4m3 blocks [w8 | wl8], then its fp16 blocks (pack_conv_f16f8_kernel) — and a block is 6 MMAs
//        instead of 9.  Barriers, ring, TMEM and tiling are untouched.
// CL = 2 or 4 (option "block2_cluster"): thread-block clusters of CL CTAs share the weight stream.  Every CTA consumes the
//        SAME sequence of 24 KB weight blocks, so block i is fetched from L2 once per cluster — by the producer of CTA
//        i % CL, with .multicast::cluster into the same ring slot of all CL CTAs — instead of once per CTA (L2 -> SM
//        traffic per tile 321 KB -> 33 + 288 / CL KB).  A slot is refilled when the MMAs of ALL CL CTAs have drained it:
//        every issuer's tcgen05.commit arrives on the `wempty` barrier of every CTA (count CL).  The CTAs of a cluster
//        therefore walk the same number of tiles (`rounds`); a CTA without a tile in the last round runs it on stale
//        operands with all stores predicated off.
template <bool F8OUT, bool F8IN = false, int CL = 0>
__global__ void __launch_bounds__(kB2Threads, 1)
block2_kernel(const Block2Params p) {
    static_assert(F8OUT || !F8IN, "F8IN implies F8OUT");
    static_assert(CL == 0 || CL == 2 || CL == 4, "cluster size (0: no clusters)");
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
    const int real_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // tiles this CTA walks: with clusters, the same for all CTAs (the last one may be a dummy for some of them)
    const int my_tiles = (CL > 1) ? (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x : real_tiles;
    constexpr uint16_t kClMask = (uint16_t)((1u << CL) - 1);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kB2Ring; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], CL ? CL : 1); }
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
    if constexpr (CL > 1) ptx::cluster_sync();      // every CTA's barriers exist before anything remote touches them
    pdl_wait();

    if (warp == 9) {
        // ===== weight producer: blocks in the order the issuer consumes them: c3(0); then c3(k+1), c4(k) =====
        uint32_t it = 0;
        auto stream_blocks = [&](const uint8_t* w, int nblk) {
            for (int s = 0; s < nblk; ++s, ++it) {
                const uint32_t slot = it % kB2Ring, ph = (it / kB2Ring) & 1;
                ptx::mbar_wait_relaxed(&wempty[slot], ph ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&wfull[slot], kB2WBlock);
                    if constexpr (CL > 1) {
                        // all CL producers are here for the same block `it`; one of them fetches it for everybody
                        if (it % CL == ptx::cluster_ctarank())
                            ptx::bulk_g2s_multicast(ring + slot * kB2WBlock, w + (size_t)s * kB2WBlock, kB2WBlock, &wfull[slot], kClMask);
                    } else {
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
            if (CL > 1 && k >= real_tiles) {                                    // dummy round: nothing to load
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
            if (F8IN) {
                constexpr uint32_t id8 = ptx::make_idesc_e4m3_f32(128, 128), id16 = ptx::make_idesc_f16_f32(128, 128);
                const int half = kch_total >> 2;                   // e4m3 blocks of this conv (= its fp16 blocks)
#pragma unroll
                for (int tap = 0; tap < 3; ++tap) {
                    // e4m3 block s < half: the two correction products for input channels [32 s, 32 s + 32), one K = 32 MMA each;
                    // fp16 block: the main products of channels [32 (s - half), + 32), two K = 16 MMAs (f8_conv_mma)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const F8Mma m = f8_conv_mma(s, half, tap, i, kch_total * 8, 128);
                        const uint64_t da = ptx::make_smem_desc(slab + m.a_off, kSlabBytes, 128);
                        const uint64_t db = ptx::make_smem_desc(b0 + m.b_off, 2048, 128);
                        if (leader) {
                            if (m.e4m3) ptx::umma_e4m3_ss(d, da, db, id8, m.mode);
                            else if (m.mode == 2) ptx::umma_f16_ss_scale_d<kF8ScaleD>(d, da, db, id16);
                            else ptx::umma_bf16_ss(d, da, db, id16, 1u);           // kind::f16; fp16 operands per the idesc
                        }
                    }
                    if (tap == 1 && it + 1 < total_blocks) {
                        ptx::mbar_wait(&wfull[(it + 1) % kB2Ring], ((it + 1) / kB2Ring) & 1);
                        ptx::tc_fence_after_sync();
                    }
                }
            } else {
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
            }
            if (leader) {
                if constexpr (CL > 1) ptx::umma_commit_multicast(&wempty[slot], kClMask);   // the slot is free when all CL CTAs have drained it
                else ptx::umma_commit(&wempty[slot]);
            }
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
        const float inv3 = F8IN ? __ldg(p.inv_sw3) : 1.f, inv4 = F8IN ? __ldg(p.inv_sw4) : 1.f;

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
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t (&v)[32] = c ? vb : va;
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias3 + c * 32 + i);
                    if (F8IN) {
                        y[i] = valid ? relu_nan(fmaf(__uint_as_float(v[i]), inv3, b4.x)) : 0.f;
                        y[i + 1] = valid ? relu_nan(fmaf(__uint_as_float(v[i + 1]), inv3, b4.y)) : 0.f;
                        y[i + 2] = valid ? relu_nan(fmaf(__uint_as_float(v[i + 2]), inv3, b4.z)) : 0.f;
                        y[i + 3] = valid ? relu_nan(fmaf(__uint_as_float(v[i + 3]), inv3, b4.w)) : 0.f;
                        continue;
                    }
                    y[i] = valid ? relu_nan(__uint_as_float(v[i]) + b4.x) : 0.f;
                    y[i + 1] = valid ? relu_nan(__uint_as_float(v[i + 1]) + b4.y) : 0.f;
                    y[i + 2] = valid ? relu_nan(__uint_as_float(v[i + 2]) + b4.z) : 0.f;
                    y[i + 3] = valid ? relu_nan(__uint_as_float(v[i + 3]) + b4.w) : 0.f;
                }
                if (F8IN) {
                    f8_range_note(y, 32, p.f8_status, 2);
                    // slabB: fp16 chunks 0..15, lo8 chunks 16..23, hi8 chunks 24..31 (16 channels per e4m3 chunk)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint4 fa, fb, lo8, hi8;
                        split16_f16f8(y + hh * 16, fa, fb, lo8, hi8);
                        const F8Dst d = f8_slab_dst(h * 4 + c * 2 + hh, 128);        // 16-channel group of the 128 X3 channels
                        uint8_t* row = slabB + (rit + 1) * 16;
                        *reinterpret_cast<uint4*>(row + d.f16) = fa;
                        *reinterpret_cast<uint4*>(row + d.f16 + kSlabBytes) = fb;
                        *reinterpret_cast<uint4*>(row + d.lo8) = lo8;
                        *reinterpret_cast<uint4*>(row + d.hi8) = hi8;
                    }
                } else {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 hi, lo;
                    split8(y + qd * 8, hi, lo);
                    uint8_t* d = slabB + (h * 8 + c * 4 + qd) * kSlabBytes + (rit + 1) * 16;
                    *reinterpret_cast<uint4*>(d) = hi;
                    *reinterpret_cast<uint4*>(d + 16 * kSlabBytes) = lo;
                }
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
            bool store = r >= 0 && r < NR && rit >= 2 && rit < 126;
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
            uint8_t* base = p.out + (size_t)(to * 16) * p.out_kch_stride + (size_t)(w + kGuard) * 16 + ((lane & 1) ? p.out_part_stride : 0);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t (&v)[32] = c ? vb : va;
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias4 + c * 32 + i);
                    if (F8IN) {
                        y[i] = relu_nan(fmaf(__uint_as_float(v[i]), inv4, b4.x));
                        y[i + 1] = relu_nan(fmaf(__uint_as_float(v[i + 1]), inv4, b4.y));
                        y[i + 2] = relu_nan(fmaf(__uint_as_float(v[i + 2]), inv4, b4.z));
                        y[i + 3] = relu_nan(fmaf(__uint_as_float(v[i + 3]), inv4, b4.w));
                        continue;
                    }
                    y[i] = relu_nan(__uint_as_float(v[i]) + b4.x);
                    y[i + 1] = relu_nan(__uint_as_float(v[i + 1]) + b4.y);
                    y[i + 2] = relu_nan(__uint_as_float(v[i + 2]) + b4.z);
                    y[i + 3] = relu_nan(__uint_as_float(v[i + 3]) + b4.w);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = max_nan(y[i], __shfl_xor_sync(0xffffffffu, y[i], 1));   // MaxPool1d(2,2)
                if (F8OUT) {
                    // k' = to*128 + channel.  Even lane: the four fp16 chunks (tape part 0, chunk to*16 + channel/8);
                    // odd lane: the e4m3 images (tape part 1: lo8 chunks [0, 296), hi8 chunks [296, 592), chunk to*8 + channel/16)
                    if (store) {
                        f8_range_note(y, 32, p.f8_status, 3);
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            uint4 fa, fb, lo8, hi8;
                            split16_f16f8(y + hh * 16, fa, fb, lo8, hi8);
                            // fc.0 operand row of window w: K index k' = to*128 + channel, 4736 "channels" in all
                            const F8Dst d = f8_tape_dst(to * 8 + h * 4 + c * 2 + hh, 4736, p.out_part_stride, p.out_kch_stride);
                            uint8_t* row = p.out + (size_t)(w + kGuard) * 16;
                            if (lane & 1) {
                                *reinterpret_cast<uint4*>(row + d.lo8) = lo8;
                                *reinterpret_cast<uint4*>(row + d.hi8) = hi8;
                            } else {
                                *reinterpret_cast<uint4*>(row + d.f16) = fa;
                                *reinterpret_cast<uint4*>(row + d.f16 + p.out_kch_stride) = fb;
                            }
                        }
                    }
                } else if (store) {
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
    if constexpr (CL > 1) ptx::cluster_sync();      // no CTA leaves while a peer may still multicast into it or signal its barriers
    if (warp == 8) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace dce
