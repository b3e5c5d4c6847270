// Microbenchmark: cycles per tcgen05.mma (M=128 per CTA, bf16, fp32 accumulate) as a function of
// N, operand layout (SWIZZLE_NONE K-major with various LBO / start shifts, SWIZZLE_128B K-major)
// and cta_group (1 or 2).  Operand contents are irrelevant (zeros).  One CTA (pair) per SM, all SMs busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../deep_contact_estimator_b200/csrc/dce_tc_ptx.cuh"
using namespace dce::ptx;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                          // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // SBO = 8 rows * 128 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)((addr >> 7) & 7) << 49;          // base offset
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Cfg { int N; int mode; int lbo_a; int shift; int reps; };   // mode 0: no swizzle, 1: SW128

template <int CG>
__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        if (CG == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
        else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    tc_fence_after_sync();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (warp == 1 && rank == 0) {
        const int M = CG == 2 ? 256 : 128;
        const uint32_t idesc = make_idesc_bf16_f32(M, c.N);
        const uint32_t a0 = smem_u32(smem) + c.shift, b0 = smem_u32(smem) + 100 * 1024;
        const int nb = CG == 2 ? c.N / 2 : c.N;          // B rows held by this CTA
        t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < c.reps; ++r) {
                // 4 K-steps per "stage" with distinct addresses, like the real kernels
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint64_t da, db;
                    if (c.mode == 0) {
                        da = make_smem_desc(a0 + 2 * k * c.lbo_a, c.lbo_a, 128);
                        db = make_smem_desc(b0 + 2 * k * nb * 16, nb * 16, 128);
                    } else {
                        da = desc_sw128(a0 + k * 32);
                        db = desc_sw128(b0 + k * 32);
                    }
                    if (CG == 1) umma_bf16_ss(tm, da, db, idesc, 1u); else mma2(tm, da, db, idesc, 1u);
                }
            }
            if (CG == 1) umma_commit(&bar);
            else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (warp == 0) {
        if (CG == 1) tmem_dealloc(tm, 512);
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

template <int CG>
double run(Cfg c, long long* d_out) {
    auto k = bench<CG>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int i = 0; i < 2; ++i) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, c, d_out);
        if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); return -1; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("run error %s\n", cudaGetErrorString(e)); return -1; }
    }
    long long h = 0;
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    return (double)h / (c.reps * 4);
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 8);
    const int reps = 512;
    printf("cycles per MMA (M=128 per CTA, K=16), all 148 SMs busy\n");
    for (int N : {64, 128, 256}) {
        printf("N=%3d  cg1 noswz lbo=2080        : %7.1f\n", N, run<1>({N, 0, 2080, 0, reps}, d_out));
        printf("N=%3d  cg1 noswz lbo=2080 shift16: %7.1f\n", N, run<1>({N, 0, 2080, 16, reps}, d_out));
        printf("N=%3d  cg1 noswz lbo=2048        : %7.1f\n", N, run<1>({N, 0, 2048, 0, reps}, d_out));
        printf("N=%3d  cg1 noswz lbo=2176        : %7.1f\n", N, run<1>({N, 0, 2176, 0, reps}, d_out));
        printf("N=%3d  cg1 sw128                 : %7.1f\n", N, run<1>({N, 1, 0, 0, reps}, d_out));
        printf("N=%3d  cg1 sw128 shift 128       : %7.1f\n", N, run<1>({N, 1, 0, 128, reps}, d_out));
        printf("N=%3d  cg2 noswz lbo=2080 (M=256): %7.1f\n", N, run<2>({N, 0, 2080, 0, reps}, d_out));
        printf("N=%3d  cg2 sw128          (M=256): %7.1f\n", N, run<2>({N, 1, 0, 0, reps}, d_out));
    }
    return 0;
}
