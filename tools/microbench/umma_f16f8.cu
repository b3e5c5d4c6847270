// Microbenchmark + numerics check for the "fp16 main + e4m3 corrections" operand splitting (DESIGN.md §8,
// tools/emulate_split_precision.py): can two bf16-pass equivalents replace the three passes of bf16x3?
//
//   x * w  ~=  f16(x) * f16(w)  +  e4m3(x - f16(x)) * e4m3(w)  +  e4m3(x) * e4m3(w - f16(w))
//
// The two correction products are kind::f8f6f4 MMAs (K = 32 per instruction: twice the MACs of a K = 16
// kind::f16 MMA in the same cycles) with STATIC power-of-two operand scales, accumulated FIRST, in a sweep of
// their own over K, into the same fp32 TMEM accumulator at scale 2^15; the first kind::f16 MMA of the second
// sweep rescales the accumulator with its scale-input-d immediate (D = A*B + D * 2^-15), so one accumulator
// serves both kinds and the TMEM budget of the kernels does not change.
//
// Part 1 (numerics): one CTA computes D[128][N] = X[128][K] * W[N][K]^T three ways — bf16x3, f16+e4m3, plain
// f16 — and the host reports the norm-wise error of each against float64.
// Part 2 (rate): cycles per 32 K-elements of one M = 128 tile for the three issue patterns, all SMs busy.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_f16f8 umma_f16f8.cu && ./umma_f16f8
//   ./umma_f16f8 --cpu   prints only the errors exact arithmetic on the rounded operands gives (no GPU needed)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "../../deep_contact_estimator_b200/csrc/dce_tc_ptx.cuh"
using namespace dce::ptx;

// static operand scales (powers of two): e4m3 covers 2^-9 .. 448
constexpr int kExL = 12;     // e4m3((x - f16(x)) * 2^12): |x| < 16 -> |.| < 32
constexpr int kEwH = 3;      // e4m3(w * sw * 2^3),  w * sw in (-2, 2)
constexpr int kExH = 1;      // e4m3(x * 2^1): saturates at |x| = 224
constexpr int kEwL = 14;     // e4m3((w - f16(w)) * sw * 2^14)
constexpr int kScaleD = 15;  // kExL + kEwH == kExH + kEwL == scale-input-d
static_assert(kExL + kEwH == kScaleD && kExH + kEwL == kScaleD, "both correction products carry the same scale");

__host__ __device__ constexpr uint32_t make_idesc(int M, int N, uint32_t afmt, uint32_t bfmt) {
    return (1u << 4) | (afmt << 7) | (bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t kF16 = 0, kBF16 = 1;      // kind::f16 operand formats
constexpr uint32_t kE4M3 = 0;                // kind::f8f6f4 operand formats

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int S>   // D = A*B + D * 2^-S
__device__ __forceinline__ void umma_f16_scale_d(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, %4;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "n"(S) : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Part 1: numerics.  Operand images in global memory are already in the UMMA SWIZZLE_NONE K-major layout
// [kchunk][row][16 bytes] (16-bit types: 8 elements per chunk, fp8: 16), so a K-slice is a flat copy.
// ---------------------------------------------------------------------------------------------------------
struct Images {
    const uint8_t *a16, *a16lo, *b16, *b16lo;      // mode 0: bf16 hi/lo; modes 1, 2: a16 / b16 hold fp16
    const uint8_t *a8lo, *a8hi, *b8hi, *b8lo;      // mode 1: e4m3
    int K, N;
};
constexpr int KC = 64;                              // K elements staged per step

template <int MODE>   // 0 = bf16x3, 1 = f16 + e4m3 (two sweeps), 2 = plain f16
__global__ void __launch_bounds__(128, 1) gemm_check(Images im, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, N = im.N;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    // smem: four regions of up to 32 KB (a 64-element K-slice of 256 rows in a 16-bit type)
    uint8_t* sA0 = smem; uint8_t* sA1 = smem + 32768; uint8_t* sB0 = smem + 65536; uint8_t* sB1 = smem + 98304;
    auto stage = [&](uint8_t* dst, const uint8_t* src, int rows, int k0, int elem_bytes) {
        // chunks [k0 * elem_bytes / 16, +KC * elem_bytes / 16) of an image with `rows` rows
        const size_t off = (size_t)(k0 * elem_bytes / 16) * rows * 16;
        const int n16 = (KC * elem_bytes / 16) * rows;
        for (int i = threadIdx.x; i < n16; i += 128)
            reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src + off)[i];
    };
    uint32_t phase = 0;
    auto run_slice = [&](auto issue) {
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) { tc_fence_after_sync(); issue(); umma_commit(&bar); }
            __syncwarp();
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after_sync();
        __syncthreads();
    };
    const uint32_t a0 = smem_u32(sA0), a1 = smem_u32(sA1), b0 = smem_u32(sB0), b1 = smem_u32(sB1);
    if (MODE == 1) {
        // sweep 1: the two e4m3 correction products, K = 32 per MMA (two 16-byte chunks)
        for (int k0 = 0; k0 < im.K; k0 += KC) {
            stage(sA0, im.a8lo, 128, k0, 1); stage(sA1, im.a8hi, 128, k0, 1);
            stage(sB0, im.b8hi, N, k0, 1);   stage(sB1, im.b8lo, N, k0, 1);
            run_slice([&] {
                const uint32_t idesc = make_idesc(128, N, kE4M3, kE4M3);
                for (int kk = 0; kk < KC / 32; ++kk) {
                    const uint32_t ao = 2 * kk * 128 * 16, bo = 2 * kk * N * 16;
                    umma_f8(tm, make_smem_desc(a0 + ao, 128 * 16, 128), make_smem_desc(b0 + bo, N * 16, 128), idesc, (k0 | kk) ? 1u : 0u);
                    umma_f8(tm, make_smem_desc(a1 + ao, 128 * 16, 128), make_smem_desc(b1 + bo, N * 16, 128), idesc, 1u);
                }
            });
        }
    }
    // main sweep(s): K = 16 per MMA
    for (int k0 = 0; k0 < im.K; k0 += KC) {
        stage(sA0, im.a16, 128, k0, 2); stage(sB0, im.b16, N, k0, 2);
        if (MODE == 0) { stage(sA1, im.a16lo, 128, k0, 2); stage(sB1, im.b16lo, N, k0, 2); }
        run_slice([&] {
            const uint32_t idesc = make_idesc(128, N, MODE == 0 ? kBF16 : kF16, MODE == 0 ? kBF16 : kF16);
            for (int kk = 0; kk < KC / 16; ++kk) {
                const uint32_t ao = 2 * kk * 128 * 16, bo = 2 * kk * N * 16;
                const uint64_t dah = make_smem_desc(a0 + ao, 128 * 16, 128), dbh = make_smem_desc(b0 + bo, N * 16, 128);
                if (MODE == 0) {
                    const uint64_t dal = make_smem_desc(a1 + ao, 128 * 16, 128), dbl = make_smem_desc(b1 + bo, N * 16, 128);
                    umma_f16(tm, dah, dbl, idesc, (k0 | kk) ? 1u : 0u);
                    umma_f16(tm, dal, dbh, idesc, 1u);
                    umma_f16(tm, dah, dbh, idesc, 1u);
                } else if (MODE == 1 && k0 == 0 && kk == 0) {
                    umma_f16_scale_d<kScaleD>(tm, dah, dbh, idesc);         // corrections come down from scale 2^15
                } else {
                    umma_f16(tm, dah, dbh, idesc, (k0 | kk) ? 1u : 0u);
                }
            }
        });
    }
    // epilogue: warp w owns TMEM lanes 32w .. 32w+31
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

// host-side image builders -------------------------------------------------------------------------------
static uint16_t h_bf16(float v) { __nv_bfloat16 b = __float2bfloat16_rn(v); uint16_t r; memcpy(&r, &b, 2); return r; }
static float f_bf16(uint16_t r) { __nv_bfloat16 b; memcpy(&b, &r, 2); return __bfloat162float(b); }
static uint16_t h_f16(float v) { __half b = __float2half_rn(v); uint16_t r; memcpy(&r, &b, 2); return r; }
static float f_f16(uint16_t r) { __half b; memcpy(&b, &r, 2); return __half2float(b); }
static uint8_t h_e4m3(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }

template <class T>
static void put(std::vector<uint8_t>& img, int rows, int r, int k, T v) {      // [k / (16/sizeof T)][row][16 B]
    const int per = 16 / (int)sizeof(T);
    memcpy(&img[((size_t)(k / per) * rows + r) * 16 + (k % per) * sizeof(T)], &v, sizeof(T));
}
static uint8_t* upload(const std::vector<uint8_t>& v) {
    uint8_t* d; cudaMalloc(&d, v.size()); cudaMemcpy(d, v.data(), v.size(), cudaMemcpyHostToDevice); return d;
}
static float gauss() {
    float u1 = (rand() + 1.0f) / (RAND_MAX + 2.0f), u2 = rand() / (float)RAND_MAX;
    return sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2);
}

static void numerics(int K, int N, bool relu_inputs, bool cpu_only) {
    std::vector<float> x((size_t)128 * K), w((size_t)N * K);
    const float bound = 1.f / sqrtf((float)K);
    float wmax = 0.f;
    for (auto& v : x) { v = gauss(); if (relu_inputs && v < 0.f) v = 0.f; }
    for (auto& v : w) { v = (rand() / (float)RAND_MAX * 2.f - 1.f) * bound; wmax = fmaxf(wmax, fabsf(v)); }
    const float sw = exp2f(-floorf(log2f(wmax)));                    // per-layer, computed at pack time
    std::vector<uint8_t> a_bh((size_t)K * 128 * 2), a_bl(a_bh.size()), a_h(a_bh.size()), a8l((size_t)K * 128), a8h(a8l.size());
    std::vector<uint8_t> b_bh((size_t)K * N * 2), b_bl(b_bh.size()), b_h(b_bh.size()), b8h((size_t)K * N), b8l(b8h.size());
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < K; ++k) {
            const float v = x[(size_t)r * K + k];
            const uint16_t bh = h_bf16(v); put(a_bh, 128, r, k, bh); put(a_bl, 128, r, k, h_bf16(v - f_bf16(bh)));
            const uint16_t hh = h_f16(v);  put(a_h, 128, r, k, hh);
            put(a8l, 128, r, k, h_e4m3(ldexpf(v - f_f16(hh), kExL)));
            put(a8h, 128, r, k, h_e4m3(ldexpf(v, kExH)));
        }
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float v = w[(size_t)n * K + k];
            const uint16_t bh = h_bf16(v); put(b_bh, N, n, k, bh); put(b_bl, N, n, k, h_bf16(v - f_bf16(bh)));
            const float vs = v * sw;
            const uint16_t hh = h_f16(vs); put(b_h, N, n, k, hh);
            put(b8h, N, n, k, h_e4m3(ldexpf(vs, kEwH)));
            put(b8l, N, n, k, h_e4m3(ldexpf(vs - f_f16(hh), kEwL)));
        }
    Images bx{}, fx{};
    float* dD = nullptr;
    if (!cpu_only) {
    bx = Images{upload(a_bh), upload(a_bl), upload(b_bh), upload(b_bl), nullptr, nullptr, nullptr, nullptr, K, N};
    fx = Images{upload(a_h), nullptr, upload(b_h), nullptr, upload(a8l), upload(a8h), upload(b8h), upload(b8l), K, N};
    cudaMalloc(&dD, (size_t)128 * N * 4);
    }
    std::vector<double> ref((size_t)128 * N);
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
            double s = 0; for (int k = 0; k < K; ++k) s += (double)x[(size_t)r * K + k] * w[(size_t)n * K + k];
            ref[(size_t)r * N + n] = s;
        }
    // what the rounded operands give in exact arithmetic (double accumulation): the error the GPU result should show,
    // up to fp32 accumulation (~1e-7).  Computed from the same rounding functions the images were built with.
    double model[3] = {0, 0, 0};
    {
        std::vector<float> xh((size_t)128 * K), xl(xh.size()), xf(xh.size()), x8l(xh.size()), x8h(xh.size());
        std::vector<float> wh((size_t)N * K), wl(wh.size()), wf(wh.size()), w8h(wh.size()), w8l(wh.size());
        auto e4 = [](float v) { return (float)__half2float(__nv_cvt_fp8_to_halfraw(h_e4m3(v), __NV_E4M3)); };
        for (size_t i = 0; i < xh.size(); ++i) {
            const float v = x[i];
            xh[i] = f_bf16(h_bf16(v)); xl[i] = f_bf16(h_bf16(v - xh[i]));
            xf[i] = f_f16(h_f16(v)); x8l[i] = e4(ldexpf(v - xf[i], kExL)); x8h[i] = e4(ldexpf(v, kExH));
        }
        for (size_t i = 0; i < wh.size(); ++i) {
            const float v = w[i], vs = v * sw;
            wh[i] = f_bf16(h_bf16(v)); wl[i] = f_bf16(h_bf16(v - wh[i]));
            wf[i] = f_f16(h_f16(vs)); w8h[i] = e4(ldexpf(vs, kEwH)); w8l[i] = e4(ldexpf(vs - wf[i], kEwL));
        }
        for (int r = 0; r < 128; ++r) {
            double num[3] = {0, 0, 0}, den = 0;
            for (int n = 0; n < N; ++n) {
                double b3 = 0, corr = 0, main16 = 0;
                for (int k = 0; k < K; ++k) {
                    const size_t a = (size_t)r * K + k, b = (size_t)n * K + k;
                    b3 += (double)xh[a] * wl[b] + (double)xl[a] * wh[b] + (double)xh[a] * wh[b];
                    corr += (double)x8l[a] * w8h[b] + (double)x8h[a] * w8l[b];
                    main16 += (double)xf[a] * wf[b];
                }
                const double want = ref[(size_t)r * N + n];
                num[0] = fmax(num[0], fabs(b3 - want));
                num[1] = fmax(num[1], fabs((main16 + ldexp(corr, -kScaleD)) / sw - want));
                num[2] = fmax(num[2], fabs(main16 / sw - want));
                den = fmax(den, fabs(want));
            }
            for (int m = 0; m < 3; ++m) model[m] = fmax(model[m], num[m] / den);
        }
    }
    int which = 0;
    auto report = [&](const char* name, float post) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  %-10s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
        std::vector<float> d((size_t)128 * N);
        cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int r = 0; r < 128; ++r) {
            double num = 0, den = 0;
            for (int n = 0; n < N; ++n) { num = fmax(num, fabs(d[(size_t)r * N + n] * post - ref[(size_t)r * N + n])); den = fmax(den, fabs(ref[(size_t)r * N + n])); }
            worst = fmax(worst, num / den);
        }
        printf("  %-10s max over rows of max|d| / max|ref| = %.3e   (exact arithmetic on the rounded operands: %.3e)\n",
               name, worst, model[which++]);
    };
    printf("K = %d, N = %d, %s inputs\n", K, N, relu_inputs ? "ReLU(N(0,1))" : "N(0,1)");
    if (cpu_only) {
        printf("  exact arithmetic on the rounded operands: bf16x3 %.3e   f16+e4m3 %.3e   f16 %.3e\n", model[0], model[1], model[2]);
        return;
    }
    constexpr int kSmem = 4 * 32768;
    cudaFuncSetAttribute(gemm_check<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(gemm_check<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(gemm_check<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    gemm_check<0><<<1, 128, kSmem>>>(bx, dD); report("bf16x3", 1.f);
    gemm_check<1><<<1, 128, kSmem>>>(fx, dD); report("f16+e4m3", 1.f / sw);
    gemm_check<2><<<1, 128, kSmem>>>(fx, dD); report("f16", 1.f / sw);
}

// ---------------------------------------------------------------------------------------------------------
// Part 2: issue rate.  pattern 0: bf16x3 (6 x K16 per 32 k), 1: two-sweep f16 + e4m3 (2 x K16 + 2 x K32 per 32 k,
// the fp8 MMAs in their own sweep), 2: the same four MMAs interleaved per 32 k, 3: fp8 only (1 x K32 per 32 k)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) rate(int N, int pattern, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (warp == 1) {
        const uint32_t i16 = make_idesc(128, N, kF16, kF16), ib = make_idesc(128, N, kBF16, kBF16), i8 = make_idesc(128, N, kE4M3, kE4M3);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 100 * 1024;
        const long long t0 = clock64();
        if (elect_one()) {
            auto da = [&](int k) { return make_smem_desc(a0 + 2 * (k & 7) * 2080, 2080, 128); };
            auto db = [&](int k) { return make_smem_desc(b0 + 2 * (k & 3) * N * 16, N * 16, 128); };
            if (pattern == 1) {
                for (int r = 0; r < reps; ++r) { umma_f8(tm, da(r), db(r), i8, 1u); umma_f8(tm, da(r + 1), db(r + 1), i8, 1u); }
                for (int r = 0; r < reps; ++r) { umma_f16(tm, da(r), db(r), i16, 1u); umma_f16(tm, da(r + 1), db(r + 1), i16, 1u); }
            } else {
                for (int r = 0; r < reps; ++r) {
                    if (pattern == 0) {
#pragma unroll
                        for (int k = 0; k < 6; ++k) umma_f16(tm, da(r + k), db(r + k), ib, 1u);
                    } else if (pattern == 2) {
                        umma_f8(tm, da(r), db(r), i8, 1u); umma_f16(tm, da(r + 1), db(r + 1), i16, 1u);
                        umma_f8(tm, da(r + 2), db(r + 2), i8, 1u); umma_f16(tm, da(r + 3), db(r + 3), i16, 1u);
                    } else {
                        umma_f8(tm, da(r), db(r), i8, 1u);
                    }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

static double run_rate(int N, int pattern, long long* d_out) {
    const int reps = 512;
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int i = 0; i < 2; ++i) {
        rate<<<148, 128, 200 * 1024>>>(N, pattern, reps, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    long long h = 0;
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    return (double)h / reps;
}

int main(int argc, char** argv) {
    const bool cpu_only = argc > 1 && !strcmp(argv[1], "--cpu");      // the model column alone: runs without a GPU
    srand(1);
    numerics(512, 128, false, cpu_only);
    numerics(4736, 256, true, cpu_only);       // fc.0's contraction length, post-ReLU operands
    numerics(192, 64, false, cpu_only);        // a conv-sized contraction
    if (cpu_only) return 0;
    long long* d_out; cudaMalloc(&d_out, 8);
    printf("cycles per 32 K-elements of one M = 128 tile (148 CTAs)\n");
    for (int N : {64, 128, 256})
        printf("N=%3d  bf16x3 %7.1f   f16+e4m3 two sweeps %7.1f   interleaved %7.1f   e4m3 alone %7.1f\n", N,
               run_rate(N, 0, d_out), run_rate(N, 1, d_out), run_rate(N, 2, d_out), run_rate(N, 3, d_out));
    return 0;
}
