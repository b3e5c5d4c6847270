// The Linear layers (fc.0, fc.3 + fc.6) on CTA PAIRS: tcgen05.mma.cta_group::2.
//
// fc.0 and fc.3 are bound by operand delivery, not by the tensor pipe (profiles/r02_experimental_modes.md: a K-stage
// of fc.0 takes the same ~1 us whether it issues 12 or 8 MMAs; every SM ingests 65 KB per stage, ~8.5 TB/s chip-wide).
// A pair of CTAs on the two SMs of a TPC executes ONE M = 256 MMA per K-step: each CTA supplies its own 128 rows of A
// and only ITS HALF of the B block (N / 2 weight columns) from the same shared-memory offsets, and gets its own
// 128 x N slice of D in its own TMEM.  Per SM and K-stage the weight bytes halve (fc.0: 33 + 32 KB -> 33 + 16 KB),
// and so do the shared-memory reads of B per MMA; the arithmetic, its order and therefore the results are
// bit-identical to the single-CTA kernel (tapgemm_kernel, dce_tc.cuh), which stays as the layer-wise ablation.
//
// Pair tile: 2 * MT M-tiles of 128 rows (CTA r owns M-tiles (2 pm + r) MT .. + MT - 1) x BN output features.
// Roles per CTA as in tapgemm_kernel (8 epilogue warps, MT issuer warps, 4 bulk-TMA producer warps); what differs:
//   * only the LEADER CTA (cluster rank 0) issues MMAs.  A K-stage is ready when its operands have landed in BOTH
//     CTAs: bulk copies complete on the mbarrier of the CTA they land in, so the peer's first issuer warp relays its
//     `full` barrier to the leader's `pfull` barrier with a remote arrive (release / acquire at cluster scope);
//   * ring slots and accumulators are released in both CTAs at once by multicast tcgen05.commit;
//   * the peer's epilogue warps hand their accumulator buffer back with remote arrives on the leader's `tempty`.
#pragma once
#include "dce_tc.cuh"

namespace dce {
namespace tc {

template <int BN, int KSA, int NSTAGE, int MT>
struct PairCfg {
    static constexpr int BH = BN / 2;                          // weight columns each CTA holds
    static constexpr int A_PART = KSA * kSlabBytes;
    static constexpr int A_TILE = 2 * A_PART;                  // hi + lo slabs of one M-tile
    static constexpr int A_BYTES = MT * A_TILE;
    static constexpr int B_TAPCH = BH * 16;
    static constexpr int B_PART = KSA * B_TAPCH;
    static constexpr int B_BYTES = 2 * B_PART;                 // this CTA's half block, packed as an n-tile of BH columns
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NBUF = (2 * MT * BN <= 512) ? 2 : 1;
    static constexpr int TMEM_COLS = NBUF * MT * BN;
    static constexpr int BAR_BYTES = ((3 * NSTAGE + 4) * 8 + 8 + 15) / 16 * 16;     // the fp32 bias copies behind it are read as float4
    static constexpr int RING_BYTES = NSTAGE * STAGE_BYTES;
    static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + kEpiWarps * (BN / 2) * 4;
    static_assert(KSA % 2 == 0, "an MMA consumes two kchunks");
    static_assert(BN % 32 == 0 && BN <= 256, "M = 256 MMAs take N up to 256");
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns: power of two");
    static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
};

// p.n_tiles = N / BN pair tiles along N; p.w_packed holds 2 * p.n_tiles packed n-tiles of BN / 2 columns
// (pack_b_kernel with BN / 2): CTA r of pair tile n streams packed n-tile 2 n + r.
template <int BN, int KSA, int NSTAGE, int EPI, int MT>
__global__ void __launch_bounds__(tapgemm_threads(MT), 1)
fcpair_kernel(const TapGemmParams p) {
    static_assert(EPI == EPI_FC_TAPE || EPI == EPI_FC_LOGITS, "Linear layers");
    static_assert(EPI != EPI_FC_LOGITS || MT == 1, "the logit-share epilogue keeps one row per thread");
    constexpr int kProducerWarp0 = kEpiWarps + MT;
    using Cfg = PairCfg<BN, KSA, NSTAGE, MT>;
    constexpr int NBUF = Cfg::NBUF;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES);   // this CTA's operands of a stage have landed
    uint64_t* pfull = full + NSTAGE;                                         // leader only: the PEER's operands have landed
    uint64_t* empty = pfull + NSTAGE;                                        // the pair's MMAs have drained the slot
    uint64_t* tfull = empty + NSTAGE;                                        // [2] accumulators complete
    uint64_t* tempty = tfull + 2;                                            // [2] leader only: both CTAs' epilogues are done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>(smem + Cfg::RING_BYTES + Cfg::BAR_BYTES);   // [8 warps][BN/2]
    float4* s_w3 = reinterpret_cast<float4*>(s_bias + kEpiWarps * (BN / 2));            // EPI_FC_LOGITS: [BN][16] fc.6 rows of this n-tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
    const int pm_tiles = (p.m_tiles + 2 * MT - 1) / (2 * MT);
    const int total_tiles = pm_tiles * p.n_tiles;
    const int my_tiles = (total_tiles - pair + npairs - 1) / npairs;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { ptx::mbar_init(&full[i], kProdWarps); ptx::mbar_init(&pfull[i], 1); ptx::mbar_init(&empty[i], MT); }
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull[b], MT); ptx::mbar_init(&tempty[b], 2 * kEpiWarps); }
        ptx::fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == kMmaWarp) { ptx::tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); ptx::tmem_relinquish_pair(); }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    ptx::cluster_sync();                                        // the peer's barriers exist before anything remote touches them
    pdl_wait();                                                 // everything the previous kernel wrote is visible from here on

    if (warp >= kProducerWarp0) {
        // ===== bulk-TMA producers: this CTA's A slabs (MT M-tiles x hi/lo x KSA) and its half B block per stage =====
        constexpr int NA = MT * 2 * KSA;
        const int pw = warp - kProducerWarp0;
        uint32_t my_bytes = 0;
        for (int c = pw; c <= NA; c += kProdWarps) my_bytes += (c < NA) ? kSlabBytes : Cfg::B_BYTES;
        uint32_t it = 0;
        for (int tile = pair; tile < total_tiles; tile += npairs) {
            const int pm = tile / p.n_tiles, n = tile % p.n_tiles;
            int m = (2 * pm + (int)rank) * MT;                          // first M-tile of this CTA
            if (m + MT > p.m_tiles) m = 0;                              // odd tail: in-bounds operands, results never stored
            const uint8_t* a_row = p.a_tape + (size_t)(128 * m + kGuard - 1) * 16;
            const uint8_t* wsrc = p.w_packed + (size_t)(2 * n + (int)rank) * p.stages * Cfg::B_BYTES;
            for (int s = 0; s < p.stages; ++s, ++it) {
                const uint32_t slot = it % NSTAGE, ph = (it / NSTAGE) & 1;
                ptx::mbar_wait_relaxed(&empty[slot], ph ^ 1);
                uint8_t* st = smem + slot * Cfg::STAGE_BYTES;
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&full[slot], my_bytes);
#pragma unroll
                    for (int c0 = 0; c0 <= NA; c0 += kProdWarps) {
                        const int c = c0 + pw;
                        if (c < NA) {
                            const int mt = c / (2 * KSA), part = (c / KSA) & 1, j = c % KSA;
                            ptx::bulk_g2s(st + mt * Cfg::A_TILE + part * Cfg::A_PART + j * kSlabBytes,
                                          a_row + (size_t)mt * 2048 + part * p.a_part_stride + (size_t)(s * KSA + j) * p.a_kch_stride,
                                          kSlabBytes, &full[slot]);
                        } else if (c == NA) {
                            ptx::bulk_g2s(st + Cfg::A_BYTES, wsrc + (size_t)s * Cfg::B_BYTES, Cfg::B_BYTES, &full[slot]);
                        }
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp >= kMmaWarp) {
        const int mt = warp - kMmaWarp;
        const uint32_t total_stages = (uint32_t)my_tiles * p.stages;
        if (rank != 0) {
            // ===== peer CTA: relay `full` to the leader (one warp; the peer issues no MMA) =====
            if (mt == 0) {
                const bool leader_lane = ptx::elect_one();
                for (uint32_t it = 0; it < total_stages; ++it) {
                    const uint32_t slot = it % NSTAGE;
                    ptx::mbar_wait(&full[slot], (it / NSTAGE) & 1);
                    if (leader_lane) ptx::mbar_arrive_remote(ptx::mapa(ptx::smem_u32(&pfull[slot]), 0));
                    __syncwarp();
                }
            }
        } else {
            // ===== leader CTA: MMA issuers, warp kMmaWarp + mt owns accumulator mt of BOTH CTAs =====
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(256, BN);
            const bool leader = ptx::elect_one();
            uint32_t it = 0;
            if (my_tiles > 0) {
                ptx::mbar_wait(&full[0], 0);
                ptx::mbar_wait_cluster(&pfull[0], 0);
                ptx::tc_fence_after_sync();
            }
            for (int tcount = 0; tcount < my_tiles; ++tcount) {
                const uint32_t buf = tcount % NBUF;
                const uint32_t d = tmem_base + buf * (MT * BN) + mt * BN;
                // the accumulator buffer is free when the epilogues of both CTAs have read it (fresh barrier: passes).  With one
                // buffer that cannot happen before this issuer's own `tfull` commit of the previous tile: wait here, not mid-stage.
                if (tcount >= NBUF) { ptx::mbar_wait_cluster(&tempty[buf], ((tcount / NBUF) & 1) ^ 1); ptx::tc_fence_after_sync(); }
                for (int s = 0; s < p.stages; ++s, ++it) {
                    const uint32_t slot = it % NSTAGE;
                    const uint32_t a0 = ptx::smem_u32(smem + slot * Cfg::STAGE_BYTES) + mt * Cfg::A_TILE;
                    const uint32_t b0 = ptx::smem_u32(smem + slot * Cfg::STAGE_BYTES) + Cfg::A_BYTES;
#pragma unroll
                    for (int kk = 0; kk < KSA / 2; ++kk) {
                        const uint32_t b_hi = b0 + (2 * kk) * Cfg::B_TAPCH;
                        const uint64_t db_hi = ptx::make_smem_desc(b_hi, Cfg::B_TAPCH, 128);
                        const uint64_t db_lo = ptx::make_smem_desc(b_hi + Cfg::B_PART, Cfg::B_TAPCH, 128);
                        const uint32_t first = (s == 0 && kk == 0) ? 0u : 1u;
                        const uint32_t a_hi = a0 + (2 * kk) * kSlabBytes + 16;              // Linear layers read the centre row
                        const uint64_t da_hi = ptx::make_smem_desc(a_hi, kSlabBytes, 128);
                        const uint64_t da_lo = ptx::make_smem_desc(a_hi + Cfg::A_PART, kSlabBytes, 128);
                        if (leader) {
                            ptx::umma_bf16_ss_pair(d, da_hi, db_lo, idesc, first);          // small terms first (as tapgemm_kernel)
                            ptx::umma_bf16_ss_pair(d, da_lo, db_hi, idesc, 1u);
                            ptx::umma_bf16_ss_pair(d, da_hi, db_hi, idesc, 1u);
                        }
                        // mid-stage: probe what the NEXT stage needs while MMAs of this one are still queued
                        if (kk == (KSA / 2 - 1) / 2 && it + 1 < total_stages) {
                            ptx::mbar_wait(&full[(it + 1) % NSTAGE], ((it + 1) / NSTAGE) & 1);
                            ptx::mbar_wait_cluster(&pfull[(it + 1) % NSTAGE], ((it + 1) / NSTAGE) & 1);
                            ptx::tc_fence_after_sync();
                        }
                    }
                    if (leader) {
                        ptx::umma_commit_pair(&empty[slot], (uint16_t)0x3);              // the slot is free in both CTAs
                        if (s == p.stages - 1) ptx::umma_commit_pair(&tfull[buf], (uint16_t)0x3);   // this accumulator is complete in both
                    }
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs): TMEM -> registers -> bias / ReLU -> bf16 hi/lo tape, or fc.6 logit shares =====
        constexpr int HALF = BN / 2;
        const int q = warp & 3, h = warp >> 2;
        const int row_in_tile = q * 32 + lane;
        float* my_bias = s_bias + warp * HALF;
        const uint32_t tempty_leader[2] = {ptx::mapa(ptx::smem_u32(&tempty[0]), 0), ptx::mapa(ptx::smem_u32(&tempty[1]), 0)};
        uint32_t tcount = 0;
        int last_n = -1;
        for (int tile = pair; tile < total_tiles; tile += npairs, ++tcount) {
            const int pm = tile / p.n_tiles, n = tile % p.n_tiles;
            const int m0 = (2 * pm + (int)rank) * MT;
            const bool live = m0 + MT <= p.m_tiles;              // false: this CTA's half of an odd tail tile
            const uint32_t buf = tcount % NBUF, tph = (tcount / NBUF) & 1;
            const int n0 = n * BN + h * HALF;
            if (n != last_n) {
                __syncwarp();
                for (int i = lane; i < HALF; i += 32) my_bias[i] = __ldg(p.bias + n0 + i);
                if (EPI == EPI_FC_LOGITS) {
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    const float4* src = reinterpret_cast<const float4*>(p.w3t) + (size_t)n * BN * 4;
                    for (int i = warp * 32 + lane; i < BN * 4; i += kEpiWarps * 32) s_w3[i] = __ldg(src + i);
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                }
                __syncwarp();
                last_n = n;
            }
            ptx::mbar_wait_relaxed(&tfull[buf], tph);
            ptx::tc_fence_after_sync();
            constexpr int CPM = HALF / 32;
            constexpr int NCH = MT * CPM;
            const uint32_t taddr0 = tmem_base + buf * (MT * BN) + h * HALF + ((uint32_t)(q * 32) << 16);
            auto chunk_addr = [&](int ci) { return taddr0 + (ci / CPM) * BN + (ci % CPM) * 32; };
            float lg[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) lg[o] = 0.f;
            auto process = [&](const uint32_t (&v)[32], int ci) {
                const int mt = ci / CPM, c0 = (ci % CPM) * 32;
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(my_bias + c0 + i);
                    y[i] = relu_nan(__uint_as_float(v[i]) + b4.x);
                    y[i + 1] = relu_nan(__uint_as_float(v[i + 1]) + b4.y);
                    y[i + 2] = relu_nan(__uint_as_float(v[i + 2]) + b4.z);
                    y[i + 3] = relu_nan(__uint_as_float(v[i + 3]) + b4.w);
                }
                if (EPI == EPI_FC_LOGITS) {
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
                if (!live) return;
                const size_t out_off = (size_t)(128 * (m0 + mt) + row_in_tile + kGuard) * 16;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 hi, lo;
                    split8(y + qd * 8, hi, lo);
                    uint8_t* dst = p.out + (size_t)((n0 + c0) / 8 + qd) * p.out_kch_stride + out_off;
                    *reinterpret_cast<uint4*>(dst) = hi;
                    *reinterpret_cast<uint4*>(dst + p.out_part_stride) = lo;
                }
            };
            {
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
            if (EPI == EPI_FC_LOGITS) {
                const int row = 128 * m0 + row_in_tile;
                if (live && row < p.n_valid) {
                    float4* dst = reinterpret_cast<float4*>(p.out_f32 + ((size_t)(n * 2 + h) * p.n_valid + row) * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(lg[4 * j], lg[4 * j + 1], lg[4 * j + 2], lg[4 * j + 3]);
                }
            }
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_remote(tempty_leader[buf]);       // (the leader's own warps arrive through the same path)
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync();              // no CTA leaves (or frees its TMEM) while its peer's MMAs / arrives may still reach it
    if (warp == kMmaWarp) ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int KSA, int NSTAGE, int EPI, int MT>
inline int launch_fcpair(Ctx& ctx, const char* name, int sm_count, const TapGemmParams& p) {
    using Cfg = PairCfg<BN, KSA, NSTAGE, MT>;
    auto kern = fcpair_kernel<BN, KSA, NSTAGE, EPI, MT>;
    constexpr int kSmem = Cfg::SMEM_BYTES + (EPI == EPI_FC_LOGITS ? BN * 64 : 0);
    static_assert(kSmem <= 232448, "exceeds 227 KB");
    static DeviceOnce attr_once;
    static int max_pairs[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (auto first_ = attr_once.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(sm_count & ~1)); cfg.blockDim = dim3(tapgemm_threads(MT)); cfg.dynamicSmemBytes = kSmem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess || n < 1) { first_.fail(); ctx.err = (e != cudaSuccess) ? e : cudaErrorLaunchOutOfResources; return DCE_ECUDA; }
        max_pairs[dev] = n;
    }
    const int tiles = ((p.m_tiles + 2 * MT - 1) / (2 * MT)) * p.n_tiles;
    int pairs = tiles;
    if (pairs > sm_count / 2) pairs = sm_count / 2;
    if (pairs > max_pairs[dev]) pairs = max_pairs[dev];
    if (pairs < 1) pairs = 1;
    DCE_KL(ctx, name, {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * pairs)); cfg.blockDim = dim3(tapgemm_threads(MT)); cfg.dynamicSmemBytes = kSmem; cfg.stream = ctx.stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        at[1].id = cudaLaunchAttributeClusterDimension;
        at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        cudaError_t le_ = cudaLaunchKernelEx(&cfg, kern, p); (void)le_; });
    return DCE_OK;
}

}  // namespace tc
}  // namespace dce
