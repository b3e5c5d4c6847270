// Launch context shared by the fp32 and tensor-core paths: the stream, a launch
// counter, and an optional per-kernel CUDA-event profiler (dce_forward_profile).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "../../include/dce.h"

namespace dce {

constexpr int kMaxProfiled = 64;

struct Profiler {
    int n = 0;
    const char* names[kMaxProfiled];
    cudaEvent_t start[kMaxProfiled], stop[kMaxProfiled];
};

struct Ctx {
    cudaStream_t stream = nullptr;
    int launches = 0;
    cudaError_t err = cudaSuccess;
    Profiler* prof = nullptr;

    // bracket one kernel launch: KL_BEGIN(ctx, "name"); kernel<<<...>>>(...); KL_END(ctx);
    inline void begin(const char* name) {
        if (prof && prof->n < kMaxProfiled) {
            prof->names[prof->n] = name;
            cudaEventCreate(&prof->start[prof->n]);
            cudaEventCreate(&prof->stop[prof->n]);
            cudaEventRecord(prof->start[prof->n], stream);
        }
    }
    inline bool end() {
        ++launches;
        cudaError_t e = cudaGetLastError();
        if (prof && prof->n < kMaxProfiled) { cudaEventRecord(prof->stop[prof->n], stream); ++prof->n; }
        if (e != cudaSuccess) { err = e; return false; }
        return true;
    }
};

// cudaFuncSetAttribute is per device: remember which devices a kernel's attributes were set on.  Handles are shared
// across host threads (include/dce.h), so the first call per device runs its set-up under a lock the others wait on:
//     static DeviceOnce once;
//     if (auto first = once.need()) { ... cudaFuncSetAttribute(...) ... }      // lock held until `first` goes out of scope
struct DeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    struct Guard {
        DeviceOnce* o; int d; bool first;
        Guard(DeviceOnce* o_, int d_, bool first_) : o(o_), d(d_), first(first_) {}
        Guard(const Guard&) = delete;
        Guard& operator=(const Guard&) = delete;
        ~Guard() { if (first && ok) o->done[d] = true; o->mu.unlock(); }
        explicit operator bool() const { return first; }
        void fail() { ok = false; }              // the set-up did not complete: the next call on this device retries it
        bool ok = true;
    };
    Guard need() {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        mu.lock();
        return Guard(this, d, !done[d]);
    }
};

// Programmatic dependent launch: the kernel may start (prologue: barrier init, TMEM alloc, weight loads) while
// the previous kernel in the stream is draining; it calls pdl_wait() before touching anything that kernel wrote.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// the same for a kernel of CTA pairs: clusters of two CTAs (the two SMs of a TPC), grid.x even
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_pairs(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace dce

#define DCE_KL(ctx, name, ...) do { (ctx).begin(name); __VA_ARGS__; if (!(ctx).end()) return DCE_ECUDA; } while (0)
