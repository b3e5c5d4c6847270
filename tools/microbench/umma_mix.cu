// Microbenchmark 2: N=128 MMA stream (conv3/conv4-like: 3 taps x hi/lo, two accumulators) with optional
// concurrent (a) bulk-TMA writes into shared memory, (b) TMEM loads by 8 warps, (c) global stores.
//   flags bit0: TMA slab copies (16 x 2080 B per 18 MMAs)   bit1: + 24.6 KB weight block per 18 MMAs
//         bit2: epilogue warps do tcgen05.ld of the other accumulator in a loop   bit3: epilogue also stores to global
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../deep_contact_estimator_b200/csrc/dce_tc_ptx.cuh"
using namespace dce::ptx;

constexpr int SLAB = 2080;
__global__ void __launch_bounds__(416, 1) mix(int flags, int reps, const uint8_t* gsrc, uint8_t* gdst, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, tbar[8], stop_flag;
    __shared__ int progress;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 416) {
        uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
        // flags bit4: random bf16 operands in [-2, 2] instead of zeros (switching power)
        uint32_t w = (flags & 16) ? ((h & 0x807f807fu) | 0x3f803f80u) : 0u;
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(w, w * 3u | ((flags & 16) ? 0x3f003f00u : 0u) & 0xbfffbfffu, w, w);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&tbar[i], 1); fence_barrier_init(); *(volatile uint64_t*)&stop_flag = 0; *(volatile int*)&progress = 0; }
    if (warp == 8) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem(); tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tm = slot;
    uint8_t* a_reg = smem;                       // 4 stages x 16640 A
    uint8_t* b_reg = smem + 4 * 16640;           // weights 98304 resident-like
    uint8_t* tma_dst = smem + 4 * 16640 + 49152; // scratch for concurrent TMA writes
    const int n_iss = (flags & 512) ? 2 : 1;
    const int iss = (warp == 8) ? 0 : ((warp == 11 && (flags & 512)) ? 1 : -1);
    if (iss >= 0) {
        const uint32_t idesc = make_idesc_bf16_f32(128, 128);
        const uint32_t a0b = smem_u32(a_reg), b0b = smem_u32(b_reg);
        long long t0 = clock64();
        if (flags & 256) {
            // per-stage structure of the real kernel: the stage's 'full' barrier was completed long ago by another warp
            for (int r = iss; r < reps; r += n_iss) {
                mbar_wait(&tbar[4 + (r & 3)], (r >> 2) & 1);
                tc_fence_after_sync();
                if (n_iss > 1) { while (*(volatile int*)&progress < r) { } }     // token: the previous stage has been issued
                const uint32_t a0 = a0b + (r & 3) * 16640, b0 = b0b + (r & 1) * 24576;
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 3; ++tap) {
                        const uint32_t b_hi = b0 + tap * 2 * 2048;
                        const uint64_t db_hi = make_smem_desc(b_hi, 2048, 128), db_lo = make_smem_desc(b_hi + 12288, 2048, 128);
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const uint32_t a_hi = a0 + mt * 8320 + tap * 16;
                            const uint64_t da_hi = make_smem_desc(a_hi, SLAB, 128), da_lo = make_smem_desc(a_hi + 4160, SLAB, 128);
                            const uint32_t d = tm + mt * 128;
                            umma_bf16_ss(d, da_hi, db_lo, idesc, 1u);
                            umma_bf16_ss(d, da_lo, db_hi, idesc, 1u);
                            umma_bf16_ss(d, da_hi, db_hi, idesc, 1u);
                        }
                    }
                    umma_commit(&tbar[1]);
                }
                __syncwarp();
                if (lane == 0) *(volatile int*)&progress = r + 1;
            }
            if (iss == (reps - 1) % n_iss && elect_one()) umma_commit(&bar);
        } else if (flags & 1024) {
            // leader elected once; the probe of the NEXT stage's barrier sits in the middle of this stage's MMAs,
            // where the tensor pipe still has queued work
            const bool leader = elect_one();
            mbar_wait(&tbar[4], 0);
            tc_fence_after_sync();
            for (int r = 0; r < reps; ++r) {
                const uint32_t a0 = a0b + (r & 3) * 16640, b0 = b0b + (r & 1) * 24576;
#pragma unroll
                for (int tap = 0; tap < 3; ++tap) {
                    const uint32_t b_hi = b0 + tap * 2 * 2048;
                    const uint64_t db_hi = make_smem_desc(b_hi, 2048, 128), db_lo = make_smem_desc(b_hi + 12288, 2048, 128);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t a_hi = a0 + mt * 8320 + tap * 16;
                        const uint64_t da_hi = make_smem_desc(a_hi, SLAB, 128), da_lo = make_smem_desc(a_hi + 4160, SLAB, 128);
                        const uint32_t d = tm + mt * 128;
                        if (leader) {
                            umma_bf16_ss(d, da_hi, db_lo, idesc, 1u);
                            umma_bf16_ss(d, da_lo, db_hi, idesc, 1u);
                            umma_bf16_ss(d, da_hi, db_hi, idesc, 1u);
                        }
                    }
                    if (tap == 1 && r + 1 < reps) { mbar_wait(&tbar[4 + ((r + 1) & 3)], ((r + 1) >> 2) & 1); tc_fence_after_sync(); }
                }
                if (leader) umma_commit(&tbar[1]);
                if (lane == 0) *(volatile int*)&progress = r + 1;
            }
            if (leader) umma_commit(&bar);
        } else if (elect_one()) {
            for (int r = 0; r < reps; ++r) {
                if ((flags & 32) && r) umma_commit(&tbar[1]);          // bit5: a tcgen05.commit after every 18 MMAs
                if ((flags & 64) && r) { umma_commit(&tbar[1]); umma_commit(&tbar[2]); }
                if (flags & 128) tc_fence_after_sync();
                const uint32_t a0 = a0b + (r & 3) * 16640, b0 = b0b + (r & 1) * 24576;
#pragma unroll
                for (int tap = 0; tap < 3; ++tap) {
                    const uint32_t b_hi = b0 + tap * 2 * 2048;
                    const uint64_t db_hi = make_smem_desc(b_hi, 2048, 128), db_lo = make_smem_desc(b_hi + 12288, 2048, 128);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t a_hi = a0 + mt * 8320 + tap * 16;
                        const uint64_t da_hi = make_smem_desc(a_hi, SLAB, 128), da_lo = make_smem_desc(a_hi + 4160, SLAB, 128);
                        const uint32_t d = tm + mt * 128;
                        umma_bf16_ss(d, da_hi, db_lo, idesc, 1u);
                        umma_bf16_ss(d, da_lo, db_hi, idesc, 1u);
                        umma_bf16_ss(d, da_hi, db_hi, idesc, 1u);
                    }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (lane == 0 && iss == 0) { *(volatile uint64_t*)&stop_flag = 1; if (blockIdx.x == 0) out[0] = t1 - t0; }
    } else if (warp == 9 && (flags & 3)) {
        // concurrent TMA writes, paced like a real producer: one "stage" per ~18 MMAs
        uint32_t ph = 0; int i = 0;
        while (*(volatile uint64_t*)&stop_flag == 0) {
            const uint32_t bytes = ((flags & 1) ? 16 * SLAB : 0) + ((flags & 2) ? 24576 : 0);
            if (elect_one()) {
                mbar_arrive_expect_tx(&tbar[0], bytes);
                if (flags & 1) for (int j = 0; j < 16; ++j) bulk_g2s(tma_dst + j * SLAB, gsrc + ((i * 16 + j) % 512) * 4096 + blockIdx.x * 64, SLAB, &tbar[0]);
                if (flags & 2) bulk_g2s(tma_dst + 33792, gsrc + 4 * 1024 * 1024 + (i % 4) * 24576, 24576, &tbar[0]);
            }
            __syncwarp();
            mbar_wait(&tbar[0], ph); ph ^= 1; ++i;
        }
    } else if (warp == 10 && (flags & (256 | 1024))) {
        // keep the four 'full' barriers completed ahead of the consumer (one arrival completes a phase)
        for (int r = 0; r < reps; ++r) {
            while (*(volatile int*)&progress < r - 3) { }           // at most 4 phases ahead: one per barrier
            if (elect_one()) mbar_arrive(&tbar[4 + (r & 3)]);
            __syncwarp();
        }
    } else if (warp < 8 && (flags & 4)) {
        const int q = warp & 3, h = warp >> 2;
        float acc = 0.f;
        while (*(volatile uint64_t*)&stop_flag == 0) {
            uint32_t v[32];
            tmem_ld32(tm + 256 + h * 64 + ((uint32_t)(q * 32) << 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += __uint_as_float(v[k]);
            if (flags & 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    reinterpret_cast<uint4*>(gdst)[((size_t)blockIdx.x * 256 + threadIdx.x) * 8 + k] = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
        }
        if (acc == 123.f) out[1] = 1;
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 8) tmem_dealloc(tm, 512);
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 16);
    uint8_t *gsrc, *gdst; cudaMalloc(&gsrc, 8 << 20); cudaMemset(gsrc, 0, 8 << 20); cudaMalloc(&gdst, 148 * 256 * 128 + 1024);
    cudaFuncSetAttribute(mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 256;
    for (int flags : {256, 1024}) {
        for (int it = 0; it < 2; ++it) { mix<<<148, 416, 200 * 1024>>>(flags, reps, gsrc, gdst, d_out); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; } }
        long long h = 0; cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        printf("flags=%2d  cycles per MMA (N=128) = %.1f\n", flags, (double)h / (reps * 18));
    }
    return 0;
}
