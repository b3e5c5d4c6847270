// Microbenchmark 4: issue cost (cycles per warp-instruction per SM sub-partition) of the CUDA-core instructions the
// converters and epilogues are made of: cvt.rn.bf16x2.f32 (F2FP), the bf16 -> fp32 unpack, FADD, max.NaN, SHFL.BFLY,
// LDS.128 (broadcast and per-lane), STS.128.  One CTA per SM, W warps per sub-partition, 8 independent chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rate alu_rate.cu && ./alu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) rate(int reps, long long* out, float* sink) {
    __shared__ __align__(16) float sm[512 * 4 + 64];
    for (int i = threadIdx.x; i < 512 * 4 + 64; i += blockDim.x) sm[i] = (float)i * 0.001f;
    __syncthreads();
    float a[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 1.0f + 0.001f * (threadIdx.x + i); u[i] = threadIdx.x * 2654435761u + i; }
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) {          // cvt.rn.bf16x2.f32
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
            } else if (OP == 1) {   // fadd
                asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
            } else if (OP == 2) {   // max.NaN
                asm volatile("max.NaN.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(__uint_as_float(u[i])));
            } else if (OP == 3) {   // shfl.bfly
                asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u[i]));
            } else if (OP == 4) {   // prmt
                asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            } else if (OP == 5) {   // lds.128 broadcast
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sbase + 16 * i));
                a[i] += v.x;
            } else if (OP == 6) {   // sts.128 per-lane (conflict-free)
                asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(sbase + 16 * threadIdx.x), "f"(a[i]) : "memory");
            } else if (OP == 7) {   // and + shift (integer truncation split)
                asm volatile("and.b32 %0, %0, 0xffff0000;" : "+r"(u[i]));
            } else if (OP == 8) {   // the whole split of one pair: cvt hi, unpack x2, sub x2, cvt lo
                uint32_t h;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a[i]), "f"(a[(i + 1) & 7]));
                const float h1 = __uint_as_float(h & 0xffff0000u), h0 = __uint_as_float(h << 16);
                uint32_t l;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(a[i] - h1), "f"(a[(i + 1) & 7] - h0));
                u[i] ^= l + h;
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, long long* d_out, float* d_sink) {
    const int reps = 2000;
    for (int warps : {4, 8, 16}) {
        rate<OP><<<148, warps * 32>>>(reps, d_out, d_sink);
        rate<OP><<<148, warps * 32>>>(reps, d_out, d_sink);
        long long c = 0;
        cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
        const double per_warp_instr = (double)c / (reps * 8.0);
        printf("%-28s %2d warps/SM (%d per sub-partition): %6.2f cycles per instruction of one warp -> %5.2f cycles per warp-instruction per sub-partition\n",
               name, warps, warps / 4, per_warp_instr, per_warp_instr / (warps / 4));
    }
}

int main() {
    long long* d_out; float* d_sink;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4);
    run<0>("cvt.rn.bf16x2.f32", d_out, d_sink);
    run<1>("add.f32", d_out, d_sink);
    run<2>("max.NaN.f32", d_out, d_sink);
    run<3>("shfl.bfly", d_out, d_sink);
    run<4>("prmt", d_out, d_sink);
    run<5>("ld.shared.v4 (broadcast)", d_out, d_sink);
    run<6>("st.shared.v4 (per lane)", d_out, d_sink);
    run<7>("and.b32", d_out, d_sink);
    run<8>("split of one pair (hi+lo)", d_out, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
