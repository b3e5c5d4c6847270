// Small-batch (latency mode, B <= 4) fully-connected layers: fp32 GEMV kernels that stream the
// fp32 weight images ([K][N], N contiguous) once per call across many CTAs.  At batch 1 the
// FC layers are a pure weight-bandwidth problem (38.8 MB for fc.0; SURVEY.md §7 "Latency mode"):
// a 128-row UMMA tile would waste 127/128 of the tensor work and serialise the weight stream
// over 8 CTAs, so the small path keeps the convolutions on the tensor cores and runs
// fc.0 / fc.3 as deterministic split-N GEMVs (no atomics: same bits every call).
//   /root/reference/src/contact_cnn.py:47-57
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace dce {
namespace small {

constexpr int kMaxB = 4;

// y[b][n0 .. n0+NPC) = relu(sum_k x[b][k] * w[k][n] + bias[n]);  NPC outputs per CTA, 256 threads
// X_TAPE: x comes from the bf16 hi/lo fc.0 operand ([part][row][K], row-major)
template <int K, int N, int NPC, bool X_TAPE>
__global__ void __launch_bounds__(256)
gemv_bias_relu_kernel(const void* __restrict__ xin, size_t part_stride, size_t kch_stride, int B,
                      const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y) {
    extern __shared__ __align__(16) float xs[];               // [B][K]
    __shared__ float red[8][kMaxB * NPC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * NPC;
    if (X_TAPE) {
        const uint8_t* tape = static_cast<const uint8_t*>(xin);
        for (int i = tid; i < B * (K / 8); i += 256) {
            const int b = i / (K / 8), kch = i % (K / 8);
            const uint8_t* src = tape + (size_t)kch * kch_stride + (size_t)b * (K * 2);      // row-major [row][K] bf16, kch_stride = 16
            const uint4 hi = *reinterpret_cast<const uint4*>(src);
            const uint4 lo = *reinterpret_cast<const uint4*>(src + part_stride);
            const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h[e]));
                const float2 lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l[e]));
                xs[b * K + kch * 8 + 2 * e] = hf.x + lf.x;
                xs[b * K + kch * 8 + 2 * e + 1] = hf.y + lf.y;
            }
        }
    } else {
        const float* x = static_cast<const float*>(xin);
        for (int i = tid; i < B * K; i += 256) xs[i] = x[i];
    }
    __syncthreads();

    float acc[kMaxB][NPC];
#pragma unroll
    for (int b = 0; b < kMaxB; ++b)
#pragma unroll
        for (int j = 0; j < NPC; ++j) acc[b][j] = 0.f;
#pragma unroll 2
    for (int k = tid; k < K; k += 256) {
        float wv[NPC];
#pragma unroll
        for (int j = 0; j < NPC; j += 4) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * N + n0 + j));
            wv[j] = w4.x; wv[j + 1] = w4.y; wv[j + 2] = w4.z; wv[j + 3] = w4.w;
        }
#pragma unroll
        for (int b = 0; b < kMaxB; ++b) {
            if (b < B) {
                const float xv = xs[b * K + k];
#pragma unroll
                for (int j = 0; j < NPC; ++j) acc[b][j] = fmaf(xv, wv[j], acc[b][j]);
            }
        }
    }
    // fixed-order reduction: lanes (shuffle tree), then the 8 warps in order
#pragma unroll
    for (int b = 0; b < kMaxB; ++b)
#pragma unroll
        for (int j = 0; j < NPC; ++j) {
            float v = acc[b][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][b * NPC + j] = v;
        }
    __syncthreads();
    if (tid < B * NPC) {
        float s = 0.f;
#pragma unroll
        for (int wdx = 0; wdx < 8; ++wdx) s += red[wdx][tid];
        const int b = tid / NPC, j = tid % NPC;
        s += __ldg(bias + n0 + j);
        y[(size_t)b * N + n0 + j] = (s < 0.f) ? 0.f : s;          // NaN-propagating ReLU
    }
}

}  // namespace small
}  // namespace dce
