// fp32 -> bf16 packing of the theta-theta triangle for the mixed-precision
// solver (eig_mixed.cu).  Kept barrier-free and intrinsic-light so that
// tests/host_emu can compile it for the CPU (SB_HOST_EMU).
#pragma once

namespace sb {

__device__ __forceinline__ unsigned bf16_bits(float x) {
    const unsigned u = __float_as_uint(x);
    if ((u & 0x7f800000u) == 0x7f800000u) return u >> 16;          // inf / nan: as is
    unsigned r = u + 0x7fffu + ((u >> 16) & 1u);                    // round to nearest even
    if ((r & 0x7f800000u) == 0x7f800000u) r = (u & 0x80000000u) | 0x7f7f0000u;   // no overflow to inf
    return r >> 16;
}

// Mb[i] = bf16(re) | bf16(im) << 16   (stand-alone converter; the sweep packs inside
// thth_build_kernel<true>; kept for tests/host_emu and ad-hoc use)
static __global__ void thth_pack_bf16_kernel(const float2* __restrict__ M, unsigned* __restrict__ Mb,
                                      size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count;
         i += (size_t)gridDim.x * blockDim.x) {
        const float2 v = M[i];
        Mb[i] = bf16_bits(v.x) | (bf16_bits(v.y) << 16);
    }
}

}  // namespace sb
