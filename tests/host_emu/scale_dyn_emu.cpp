// CPU run of csrc/scale_dyn.cu under the SIMT emulator (simt.h): the kernels
// have no barriers, shuffles or atomics, every fiber simply runs to completion.
// Built and called by tests/test_host_emulation.py (g++, ctypes).  TEST
// INFRASTRUCTURE: verifies the device code's arithmetic and indexing without
// a GPU; it is not a CPU fallback (nothing in scintools_b200 loads it).
#define SB_HOST_EMU 1
#include "simt.h"

#include "../../scintools_b200/csrc/scale_dyn.cu"

extern "C" void emu_scale_dyn(const float* dyn, int nf, int nt, int flip, const float* a,
                              const float* cp, const float* inv, const float* g, float p0,
                              float pn, const int* idx, const float* w4, int nlam, float* M,
                              float* out) {
    const unsigned gx = (unsigned)((nt + 127) / 128);
    for (unsigned b = 0; b < gx; ++b)
        emu::run_block(emu::Dim3{128, 1, 1}, emu::Dim3{b, 0, 0}, emu::Dim3{gx, 1, 1}, [&]() {
            sb::spline_moments_kernel(dyn, nf, nt, flip, a, cp, inv, g, p0, pn, M);
        });
    const unsigned ex = (unsigned)((nt + 255) / 256);
    for (unsigned by = 0; by < (unsigned)nlam; ++by)
        for (unsigned b = 0; b < ex; ++b)
            emu::run_block(emu::Dim3{256, 1, 1}, emu::Dim3{b, by, 0}, emu::Dim3{ex, (unsigned)nlam, 1},
                           [&]() {
                               sb::spline_eval_kernel(dyn, M, nf, nt, flip, idx,
                                                      reinterpret_cast<const float4*>(w4), nlam, out);
                           });
}
