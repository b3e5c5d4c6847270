// tcgen05 bf16x3 path (DCE_PREC_BF16X3) — placeholder until the kernels land.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dce_common.cuh"

namespace dce {
namespace tc {

struct PackedLayout { size_t begin, end; };
inline PackedLayout make_packed_layout(size_t base) { return PackedLayout{base, base}; }
inline int pack(char*, const PackedLayout&, const float* const*, Ctx&) { return DCE_OK; }
inline size_t workspace_bytes(int64_t) { return 256; }
inline int run(const char*, const PackedLayout&, int, const float*, bool, int64_t, int64_t, float*, int32_t*, uint8_t*,
               char*, Ctx&) { return DCE_EUNSUPPORTED; }

}  // namespace tc
}  // namespace dce
