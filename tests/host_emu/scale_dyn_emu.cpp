// CPU emulation of csrc/scale_dyn.cu: the kernels have no barriers, shuffles
// or atomics, so running every (block, thread) sequentially is exact.  Built
// and called by tests/test_host_emulation.py (g++, ctypes).  TEST
// INFRASTRUCTURE: verifies the device code's arithmetic and indexing without
// a GPU; it is not a CPU fallback (nothing in scintools_b200 loads it).
#define SB_HOST_EMU 1
#include <cmath>
#include <cstddef>
#include <cstdint>

struct emu_uint3 { unsigned x, y, z; };
static emu_uint3 blockIdx, threadIdx, blockDim, gridDim;
struct float4 { float x, y, z, w; };
#define __global__
#define __device__
#define __restrict__
#define __forceinline__ inline

#include "../../scintools_b200/csrc/scale_dyn.cu"

extern "C" void emu_scale_dyn(const float* dyn, int nf, int nt, int flip, const float* a,
                              const float* cp, const float* inv, const float* g, float p0,
                              float pn, const int* idx, const float* w4, int nlam, float* M,
                              float* out) {
    blockDim = {128, 1, 1};
    gridDim = {(unsigned)((nt + 127) / 128), 1, 1};
    for (unsigned b = 0; b < gridDim.x; ++b)
        for (unsigned t = 0; t < blockDim.x; ++t) {
            blockIdx = {b, 0, 0};
            threadIdx = {t, 0, 0};
            sb::spline_moments_kernel(dyn, nf, nt, flip, a, cp, inv, g, p0, pn, M);
        }
    blockDim = {256, 1, 1};
    gridDim = {(unsigned)((nt + 255) / 256), (unsigned)nlam, 1};
    for (unsigned by = 0; by < gridDim.y; ++by)
        for (unsigned b = 0; b < gridDim.x; ++b)
            for (unsigned t = 0; t < blockDim.x; ++t) {
                blockIdx = {b, by, 0};
                threadIdx = {t, 0, 0};
                sb::spline_eval_kernel(dyn, M, nf, nt, flip, idx,
                                       reinterpret_cast<const float4*>(w4), nlam, out);
            }
}
