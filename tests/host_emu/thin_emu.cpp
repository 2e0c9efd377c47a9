// CPU run of the "thin" theta-theta device code (csrc/thin.cu: thin_prep_kernel,
// thin_indexerr_kernel, thin_build_kernel, thin_sv_kernel<256>) under the SIMT
// emulator, launch geometry as in sb::thin_sweep.  TEST INFRASTRUCTURE.
#define SB_HOST_EMU 1
#include "simt.h"

#include <type_traits>

namespace sb {
alignas(128) unsigned char smem_raw[256 * 1024];
}
#include "../../scintools_b200/csrc/thin.cu"

extern "C" int emu_thin_sweep(const float* cs, long long ntau, long long nfd, double tau1,
                              double dtau, double tau_max, double fd1, double dfd,
                              const double* th1, int n1, const double* th2, int n2,
                              double center_cut, int power, const double* eta1, const double* eta2,
                              int neta, double tol, int max_iter, double* sv, int* status, int* n1r,
                              int* n2r, int* iters) {
    using namespace sb;
    ThinGeom t;
    ThthGeom& g = t.g;
    g.cs = reinterpret_cast<const float2*>(cs);
    g.ntau = ntau; g.nfd = nfd; g.cs_pitch = nfd; g.cs_half = 0;
    g.tau0 = tau1; g.dtau = dtau; g.half_dtau = dtau / 2; g.tau_absmax = tau_max;
    g.fd0 = fd1; g.dfd = dfd; g.half_dfd = dfd / 2; g.fd_half = 0.0;
    g.inv_dtau = 1.0 / dtau; g.inv_dfd = 1.0 / dfd;
    g.th = th1; g.n = n1; g.coherent = 1;
    t.th2 = th2; t.n2 = n2; t.tau_max = tau_max; t.center_cut = center_cut; t.power = power;
    if (max_iter <= 0 || max_iter > SB_LANCZOS_MAXIT) max_iter = SB_LANCZOS_MAXIT;
    const int ld1 = (n1 + 31) / 32 * 32, ld2 = (n2 + 31) / 32 * 32;
    std::vector<int> idx1((size_t)neta * ld1, 0), idx2((size_t)neta * ld2, 0);
    std::vector<float2> M((size_t)neta * ld1 * ld2);
    std::memset(M.data(), 0xff, M.size() * sizeof(float2));
    for (int e = 0; e < neta; ++e) status[e] = 0;
    for (int e = 0; e < neta; ++e)
        emu::run_block(emu::Dim3{32, 1, 1}, emu::Dim3{(unsigned)e, 0, 0}, emu::Dim3{(unsigned)neta, 1, 1},
                       [&]() {
                           thin_prep_kernel(t, eta1, eta2, neta, ld1, ld2, idx1.data(), idx2.data(),
                                            n1r, n2r);
                       });
    for (int e = 0; e < neta; ++e)
        for (unsigned bx = 0; bx < 2; ++bx)
            emu::run_block(emu::Dim3{256, 1, 1}, emu::Dim3{bx, (unsigned)e, 0},
                           emu::Dim3{2, (unsigned)neta, 1},
                           [&]() { thin_indexerr_kernel(t, eta1, eta2, status); });
    for (int e = 0; e < neta; ++e)
        for (unsigned ta = 0; ta < (unsigned)(ld2 / 32); ++ta)
            for (unsigned tb = 0; tb < (unsigned)(ld1 / 32); ++tb)
                emu::run_block(emu::Dim3{32, 8, 1}, emu::Dim3{(unsigned)e, ta, tb},
                               emu::Dim3{(unsigned)neta, (unsigned)(ld2 / 32), (unsigned)(ld1 / 32)},
                               [&]() {
                                   thin_build_kernel(t, eta1, eta2, 0, ld1, ld2, idx1.data(),
                                                     idx2.data(), n1r, n2r, M.data(), status);
                               });
    for (int e = 0; e < neta; ++e)
        emu::run_block(emu::Dim3{256, 1, 1}, emu::Dim3{(unsigned)e, 0, 0}, emu::Dim3{(unsigned)neta, 1, 1},
                       [&]() {
                           thin_sv_kernel<256>(M.data(), ld1, ld2, n1r, n2r, 0, sv, status, iters,
                                               tol, 2e-7, max_iter);
                       });
    return 0;
}
