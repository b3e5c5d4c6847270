// fp32 FFMA kernels: the DCE_PREC_FP32 arithmetic mode.
//
// This mode keeps every product in fp32 on the CUDA cores.  It is the
// full-precision arm of the library (error vs. the fp64 arbiter ~3e-7, the
// same as the reference's own fp32 forward) and the on-device cross-check for
// the tensor-core mode; its roofline is the fp32 FMA pipe, ~74 TFLOP/s.
//
// Layouts (all channels-last so a conv tap is a contiguous row of the input):
//   conv weights   wp[tap][cin][cout]           from W[cout][cin][tap]   src/contact_cnn.py:11-40
//   fc.0 weight    w1p[t*128 + c][n]            from W[n][c*37 + t]      src/contact_cnn.py:48,64
//   fc.3 weight    w2p[k][n]                    from W[n][k]             src/contact_cnn.py:52
//   fc.6 weight    unchanged [16][512]                                   src/contact_cnn.py:56
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dce {
namespace fp32 {

// ---------------------------------------------------------------------------
// K0 (fp32 part): repack
// ---------------------------------------------------------------------------
__global__ void pack_conv_kernel(const float* __restrict__ w, float* __restrict__ wp, int cout, int cin) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;          // index into wp[tap][cin][cout]
    int n = 3 * cin * cout;
    if (i >= n) return;
    int o = i % cout, c = (i / cout) % cin, k = i / (cout * cin);
    wp[i] = w[(o * cin + c) * 3 + k];
}

// dst[k'][n] = src[n][perm(k')]; flat == 1: perm(k') = k'; else k' = t*128+c -> c*37+t
__global__ void pack_fc_kernel(const float* __restrict__ w, float* __restrict__ wp, int N, int K, int permute) {
    __shared__ float tile[32][33];
    int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y;                 // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        int n = n0 + j, k = k0 + tx;                        // read along k' (dst order)
        float v = 0.f;
        if (n < N && k < K) {
            int ks = k;
            if (permute) { int t = k / 128, c = k % 128; ks = c * 37 + t; }
            v = w[(size_t)n * K + ks];
        }
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int k = k0 + j, n = n0 + tx;
        if (n < N && k < K) wp[(size_t)k * N + n] = tile[tx][j];
    }
}

// ---------------------------------------------------------------------------
// one conv layer on a smem-resident window, thread tile = PT positions x 4 couts
// in : smem rows r = position + 1 (row 0 and rows > T are zero), stride CIN
// out: POOL ? pooled rows : rows, same row convention, stride COUT
// ---------------------------------------------------------------------------
template <int CIN, int CIN_STRIDE, int COUT, int T, bool POOL, bool TO_GLOBAL>
__device__ __forceinline__ void conv3_relu_layer(const float* __restrict__ in, float* __restrict__ out,
                                                 const float* __restrict__ wp, const float* __restrict__ bias) {
    constexpr int PT = 10;                         // positions per thread
    constexpr int CG = COUT / 4;                   // cout groups (float4)
    constexpr int PG = 256 / CG;                   // position groups
    static_assert(PG * PT >= T, "tile does not cover the window");
    const int cg = threadIdx.x % CG, pg = threadIdx.x / CG;
    const int p0 = pg * PT, co = cg * 4;

    float acc[PT][4];
    const float4 b4 = *reinterpret_cast<const float4*>(bias + co);
#pragma unroll
    for (int i = 0; i < PT; ++i) { acc[i][0] = b4.x; acc[i][1] = b4.y; acc[i][2] = b4.z; acc[i][3] = b4.w; }

    const float* xin = in + p0 * CIN_STRIDE;       // row p0 = position p0-1
#pragma unroll 2
    for (int c = 0; c < CIN; ++c) {
        float xv[PT + 2];
#pragma unroll
        for (int i = 0; i < PT + 2; ++i) xv[i] = xin[i * CIN_STRIDE + c];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp + ((size_t)k * CIN + c) * COUT + co));
#pragma unroll
            for (int i = 0; i < PT; ++i) {
                acc[i][0] = fmaf(xv[i + k], w4.x, acc[i][0]);
                acc[i][1] = fmaf(xv[i + k], w4.y, acc[i][1]);
                acc[i][2] = fmaf(xv[i + k], w4.z, acc[i][2]);
                acc[i][3] = fmaf(xv[i + k], w4.w, acc[i][3]);
            }
        }
    }
    if (POOL) {
        // MaxPool1d(2,2) after ReLU (src/contact_cnn.py:22-25): max(relu(a),relu(b)) = relu(max(a,b))
#pragma unroll
        for (int i = 0; i < PT; i += 2) {
            const int tp = (p0 + i) / 2;                     // pooled position
            if (p0 + i + 1 < T) {                            // floor: an unpaired last sample is dropped
                float4 v;
                v.x = fmaxf(fmaxf(acc[i][0], acc[i + 1][0]), 0.f);
                v.y = fmaxf(fmaxf(acc[i][1], acc[i + 1][1]), 0.f);
                v.z = fmaxf(fmaxf(acc[i][2], acc[i + 1][2]), 0.f);
                v.w = fmaxf(fmaxf(acc[i][3], acc[i + 1][3]), 0.f);
                // fmaxf drops NaN operands; the reference propagates them.
                if (acc[i][0] != acc[i][0] || acc[i + 1][0] != acc[i + 1][0]) v.x = __int_as_float(0x7fc00000);
                if (acc[i][1] != acc[i][1] || acc[i + 1][1] != acc[i + 1][1]) v.y = __int_as_float(0x7fc00000);
                if (acc[i][2] != acc[i][2] || acc[i + 1][2] != acc[i + 1][2]) v.z = __int_as_float(0x7fc00000);
                if (acc[i][3] != acc[i][3] || acc[i + 1][3] != acc[i + 1][3]) v.w = __int_as_float(0x7fc00000);
                float* dst = TO_GLOBAL ? out + tp * COUT + co : out + (tp + 1) * COUT + co;
                *reinterpret_cast<float4*>(dst) = v;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            if (p0 + i < T) {
                float4 v;
                v.x = acc[i][0] > 0.f || acc[i][0] != acc[i][0] ? acc[i][0] : 0.f;
                v.y = acc[i][1] > 0.f || acc[i][1] != acc[i][1] ? acc[i][1] : 0.f;
                v.z = acc[i][2] > 0.f || acc[i][2] != acc[i][2] ? acc[i][2] : 0.f;
                v.w = acc[i][3] > 0.f || acc[i][3] != acc[i][3] ? acc[i][3] : 0.f;
                *reinterpret_cast<float4*>(out + (p0 + i + 1) * COUT + co) = v;
            }
        }
    }
}

struct ConvParams {
    const float* w1; const float* b1;   // [3][54][64]
    const float* w2; const float* b2;   // [3][64][64]
    const float* w3; const float* b3;   // [3][64][128]
    const float* w4; const float* b4;   // [3][128][128]
};

constexpr int kRows150 = 162;           // 1 + 160 + 1 rows (tile covers 160 positions)
constexpr int kRows75  = 82;            // 1 + 80 + 1
constexpr int kBuf0Floats = kRows150 * 54 > kRows75 * 64 ? kRows150 * 54 : kRows75 * 64;   // X0 / X2
constexpr int kBuf1Floats = kRows150 * 64 > kRows75 * 128 ? kRows150 * 64 : kRows75 * 128; // X1 / X3
constexpr int kConvSmemBytes = (kBuf0Floats + kBuf1Floats) * 4;

// One CTA per window.  NORMALIZE: x is the raw [T][54] log and window w starts
// at row first + w (utils/data_handler.py:55-56 fused into the load);
// otherwise x is [B][150][54] already-normalised windows.
// act4: [B][37*128] fp32, index t*128 + c.
template <bool NORMALIZE>
__global__ void __launch_bounds__(256, 2)
conv_stack_kernel(const float* __restrict__ x, int64_t first, int64_t n_windows, ConvParams p, float* __restrict__ act4) {
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = smem + kBuf0Floats;
    __shared__ float s_mean[64], s_rstd[64];

    for (int64_t w = blockIdx.x; w < n_windows; w += gridDim.x) {
        const float* src = NORMALIZE ? x + (first + w) * 54 : x + w * (150 * 54);
        __syncthreads();                                     // previous window's readers are done
        for (int i = threadIdx.x; i < kBuf0Floats; i += 256) buf0[i] = 0.f;
        for (int i = threadIdx.x; i < kBuf1Floats; i += 256) buf1[i] = 0.f;
        __syncthreads();
        for (int i = threadIdx.x; i < 150 * 54; i += 256) buf0[54 + i] = __ldg(src + i);
        __syncthreads();
        if (NORMALIZE) {
            // two-pass mean / unbiased std per channel (utils/data_handler.py:55-56)
            // 4 threads per channel; threads past channel 53 redo channel 53 so the
            // full-warp shuffles stay convergent, and simply do not publish.
            const int cr = threadIdx.x >> 2, q = threadIdx.x & 3;
            const int c = cr < 54 ? cr : 53;
            float s = 0.f;
            for (int t = q; t < 150; t += 4) s += buf0[54 + t * 54 + c];
            s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
            const float mean = s / 150.f;
            float v = 0.f;
            for (int t = q; t < 150; t += 4) { float d = buf0[54 + t * 54 + c] - mean; v = fmaf(d, d, v); }
            v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (q == 0 && cr < 54) { s_mean[c] = mean; s_rstd[c] = sqrtf(v / 149.f); }
            __syncthreads();
            for (int i = threadIdx.x; i < 150 * 54; i += 256) {
                const int c = i % 54;
                buf0[54 + i] = (buf0[54 + i] - s_mean[c]) / s_rstd[c];   // division, as the reference does
            }
            __syncthreads();
        }
        conv3_relu_layer<54, 54, 64, 150, false, false>(buf0, buf1, p.w1, p.b1);   // src/contact_cnn.py:11-16
        __syncthreads();
        for (int i = threadIdx.x; i < kBuf0Floats; i += 256) buf0[i] = 0.f;
        __syncthreads();
        conv3_relu_layer<64, 64, 64, 150, true, false>(buf1, buf0, p.w2, p.b2);    // :17-25
        __syncthreads();
        for (int i = threadIdx.x; i < kBuf1Floats; i += 256) buf1[i] = 0.f;
        __syncthreads();
        conv3_relu_layer<64, 64, 128, 75, false, false>(buf0, buf1, p.w3, p.b3);   // :29-34
        __syncthreads();
        conv3_relu_layer<128, 128, 128, 75, true, true>(buf1, act4 + w * 4736, p.w4, p.b4);   // :35-43
    }
}

// ---------------------------------------------------------------------------
// C[M][N] = relu?(A[M][K] * Bp[K][N] + bias[N]); 128x128x16 tiles, 8x8 per thread
// ---------------------------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(256)
sgemm_bias_kernel(const float* __restrict__ A, const float* __restrict__ Bp, const float* __restrict__ bias,
                  float* __restrict__ C, int M, int N, int K) {
    constexpr int BM = 128, BN = 128, BK = 16;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % 16, ty = tid / 16;                  // 16 x 16 threads, each 8x8 (split 4+4)

    // A tile loader: 128 rows x 16 k = 512 float4; 2 per thread
    const int a_row = tid / 4, a_k4 = (tid % 4) * 4;         // rows a_row, a_row + 64
    // B tile loader: 16 k x 128 n = 512 float4; 2 per thread
    const int b_k = tid / 32, b_n4 = (tid % 32) * 4;         // k rows b_k, b_k + 8

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + a_row + h * 64;
            ra[h] = (m < M) ? __ldg(reinterpret_cast<const float4*>(A + (size_t)m * K + k0 + a_k4)) : make_float4(0, 0, 0, 0);
            const int n = n0 + b_n4;
            rb[h] = (n < N) ? __ldg(reinterpret_cast<const float4*>(Bp + (size_t)(k0 + b_k + h * 8) * N + n)) : make_float4(0, 0, 0, 0);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = a_row + h * 64;
            As[buf][a_k4 + 0][r] = ra[h].x; As[buf][a_k4 + 1][r] = ra[h].y;
            As[buf][a_k4 + 2][r] = ra[h].z; As[buf][a_k4 + 3][r] = ra[h].w;
            *reinterpret_cast<float4*>(&Bs[buf][b_k + h * 8][b_n4]) = rb[h];
        }
    };

    gload(0); sstore(0); __syncthreads();
    const int nk = K / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) { sstore(buf ^ 1); }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n >= N) continue;
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
            float4 v = make_float4(acc[i][jh * 4 + 0] + b4.x, acc[i][jh * 4 + 1] + b4.y,
                                   acc[i][jh * 4 + 2] + b4.z, acc[i][jh * 4 + 3] + b4.w);
            if (RELU) {
                v.x = v.x > 0.f || v.x != v.x ? v.x : 0.f; v.y = v.y > 0.f || v.y != v.y ? v.y : 0.f;
                v.z = v.z > 0.f || v.z != v.z ? v.z : 0.f; v.w = v.w > 0.f || v.w != v.w ? v.w : 0.f;
            }
            *reinterpret_cast<float4*>(C + (size_t)m * N + n) = v;
        }
    }
}

// ---------------------------------------------------------------------------
// fc.6 (512 -> 16) + argmax + decimal2binary; one warp per window
//   src/contact_cnn.py:56-57, src/inference_one_seq.py:26-27,59-62
// ---------------------------------------------------------------------------
__device__ __forceinline__ void argmax_bits_store(const float (&y)[16], int64_t w, float* logits, int32_t* cls, uint8_t* bits) {
    // torch.max semantics: first maximal index; a NaN is "greater" than everything and the first NaN wins.
    int best = 0; float bv = y[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) {
        const bool better = (y[j] > bv) || (y[j] != y[j] && bv == bv);
        if (better) { bv = y[j]; best = j; }
    }
    if (logits) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(logits + w * 16 + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
    }
    if (cls) cls[w] = best;
    if (bits) {
        uchar4 b4 = make_uchar4((best >> 3) & 1, (best >> 2) & 1, (best >> 1) & 1, best & 1);   // MSB first = leg 0
        *reinterpret_cast<uchar4*>(bits + w * 4) = b4;
    }
}

// block = 16 windows x 16 outputs (one thread per logit); weights k-major in smem so the 16 output
// threads of a window read 64 contiguous bytes and different windows broadcast
__global__ void __launch_bounds__(256)
fc3_argmax_kernel(const float* __restrict__ h2, const float* __restrict__ w3t, const float* __restrict__ b3,
                  int64_t n_windows, float* __restrict__ logits, int32_t* __restrict__ cls, uint8_t* __restrict__ bits) {
    extern __shared__ __align__(16) float fc3_smem[];
    float* ws = fc3_smem;                  // [512][16]
    float* hs = fc3_smem + 512 * 16;       // [16][516]
    const int tid = threadIdx.x, o = tid & 15, wl = tid >> 4;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    {   // w3t is already k-major [512][16]; issue all 8 loads, then all 8 stores (one memory latency, not eight)
        float4 wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = __ldg(reinterpret_cast<const float4*>(w3t) + tid + 256 * j);
#pragma unroll
        for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(ws)[tid + 256 * j] = wv[j];
    }
    const float bias = __ldg(b3 + o);
    asm volatile("griddepcontrol.wait;" ::: "memory");        // h2 comes from the previous kernel (no-op without PDL)
    for (int64_t w0 = (int64_t)blockIdx.x * 16; w0 < n_windows; w0 += (int64_t)gridDim.x * 16) {
        __syncthreads();
        {   // 16 rows x 128 float4: all loads first, then the stores
            float4 hv8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = tid + 256 * j, r = i >> 7, c4 = i & 127;
                hv8[j] = (w0 + r < n_windows) ? __ldg(reinterpret_cast<const float4*>(h2 + (w0 + r) * 512) + c4) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = tid + 256 * j, r = i >> 7, c4 = i & 127;
                *reinterpret_cast<float4*>(hs + r * 516 + c4 * 4) = hv8[j];
            }
        }
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};                                // 4 partial sums: shorter dependency chains
        const float* hrow = hs + wl * 516;
#pragma unroll 8
        for (int k = 0; k < 512; k += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hrow + k);
            acc[0] = fmaf(hv.x, ws[(k + 0) * 16 + o], acc[0]);
            acc[1] = fmaf(hv.y, ws[(k + 1) * 16 + o], acc[1]);
            acc[2] = fmaf(hv.z, ws[(k + 2) * 16 + o], acc[2]);
            acc[3] = fmaf(hv.w, ws[(k + 3) * 16 + o], acc[3]);
        }
        const float y_mine = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + bias;
        // gather the window's 16 logits into its first lane (lanes 0 and 16 of each warp)
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) y[j] = __shfl_sync(0xffffffffu, y_mine, (threadIdx.x & 16) + j);
        const int64_t w = w0 + wl;
        if (o == 0 && w < n_windows) argmax_bits_store(y, w, logits, cls, bits);
    }
}
constexpr int kFc3SmemBytes = (512 * 16 + 16 * 516) * 4;

// fc.6 folded into fc.3's epilogue (tc::EPI_FC_LOGITS): every window's 16 logits arrive as `nparts` shares
// part[s][w][16]; add them in a fixed order (+ bias), argmax, contact bits.  One thread per window.
__global__ void __launch_bounds__(128)
logit_shares_argmax_kernel(const float* __restrict__ part, const float* __restrict__ b3, int64_t n_windows, int nparts,
                           float* __restrict__ logits, int32_t* __restrict__ cls, uint8_t* __restrict__ bits) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = __ldg(b3 + j);
    asm volatile("griddepcontrol.wait;" ::: "memory");        // the shares come from the previous kernel
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    float s[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = 0.f;
    for (int sidx = 0; sidx < nparts; ++sidx) {
        const float4* src = reinterpret_cast<const float4*>(part + ((size_t)sidx * n_windows + w) * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 v = __ldcg(src + j);
            s[4 * j] += v.x; s[4 * j + 1] += v.y; s[4 * j + 2] += v.z; s[4 * j + 3] += v.w;
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] += s[j];
    argmax_bits_store(y, w, logits, cls, bits);
}

// ---------------------------------------------------------------------------
// decimal2binary on labels; accuracy counters
// ---------------------------------------------------------------------------
__global__ void decimal2binary_kernel(const int64_t* __restrict__ cls, int64_t n, uint8_t* __restrict__ bits) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t v = cls[i];
    *reinterpret_cast<uchar4*>(bits + i * 4) = make_uchar4((v & 8) != 0, (v & 4) != 0, (v & 2) != 0, (v & 1) != 0);
}

// float64 log on disk (utils/mat2numpy.py:73,80 saves float64 .npy) -> the fp32 device stream, round-to-nearest
// exactly as `torch.FloatTensor(np.load(...))` (utils/data_handler.py:21-26) does on the host
__global__ void ingest_f64_kernel(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
        if (i + 1 < n) {
            const double2 v = *reinterpret_cast<const double2*>(src + i);
            *reinterpret_cast<float2*>(dst + i) = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));
        } else {
            dst[i] = __double2float_rn(src[i]);
        }
    }
}

// counts[0] class agreement, [1..4] per-leg bit agreement, [5 + 4 leg + 2 gt + pred] the four 2x2 per-leg confusion
// matrices (rows = ground truth, columns = prediction, as sklearn's confusion_matrix(gt, pred): src/test.py:19-27),
// [21 + 16 gt + pred] the 16x16 class confusion matrix (what precision_score / jaccard_score(average='weighted') of
// src/test.py:50-70 are functions of).  Labels outside [0, 16) only count as class disagreements.
constexpr int kNumCounts = 5 + 16 + 256;
__global__ void __launch_bounds__(256)
accuracy_counts_kernel(const int32_t* __restrict__ cls, const int64_t* __restrict__ labels, int64_t n,
                       unsigned long long* __restrict__ counts) {
    __shared__ unsigned int h[kNumCounts];
    for (int j = threadIdx.x; j < kNumCounts; j += blockDim.x) h[j] = 0;
    __syncthreads();
    unsigned int c[5] = {0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = cls[i], g = labels[i];
        c[0] += (p == g);
        if ((unsigned long long)g < 16ull && (unsigned long long)p < 16ull) {
            unsigned int legs = 0;                       // 4 x 4-bit fields: which cell of each leg's 2x2 matrix
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const unsigned int pb = (unsigned int)(p >> (3 - l)) & 1u, gb = (unsigned int)(g >> (3 - l)) & 1u;
                c[1 + l] += (pb == gb);
                legs |= (2u * gb + pb) << (4 * l);
            }
#pragma unroll
            for (int l = 0; l < 4; ++l) atomicAdd(&h[5 + 4 * l + ((legs >> (4 * l)) & 3u)], 1u);
            atomicAdd(&h[21 + 16 * (int)g + (int)p], 1u);
        }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        unsigned int v = c[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&h[j], v);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kNumCounts; j += blockDim.x)
        if (h[j]) atomicAdd(counts + j, (unsigned long long)h[j]);
}

}  // namespace fp32
}  // namespace dce
