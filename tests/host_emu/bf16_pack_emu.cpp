// CPU emulation of csrc/bf16_pack.cuh (thth_pack_bf16_kernel), see scale_dyn_emu.cpp.
#define SB_HOST_EMU 1
#include <cstddef>
#include <cstdint>
#include <cstring>

struct emu_uint3 { unsigned x, y, z; };
static emu_uint3 blockIdx, threadIdx, blockDim, gridDim;
struct float2 { float x, y; };
#define __global__
#define __device__
#define __restrict__
#define __forceinline__ inline
static inline unsigned __float_as_uint(float x) { unsigned u; std::memcpy(&u, &x, 4); return u; }

#include "../../scintools_b200/csrc/bf16_pack.cuh"

extern "C" void emu_pack_bf16(const float* M, unsigned* Mb, long count) {
    blockDim = {256, 1, 1};
    gridDim = {7, 1, 1};
    for (unsigned b = 0; b < gridDim.x; ++b)
        for (unsigned t = 0; t < blockDim.x; ++t) {
            blockIdx = {b, 0, 0};
            threadIdx = {t, 0, 0};
            sb::thth_pack_bf16_kernel(reinterpret_cast<const float2*>(M), Mb, (size_t)count);
        }
}
