// CPU run of the phase-retrieval device code of csrc/retrieval.cu under the
// SIMT emulator: rev_scatter_kernel + rev_finalise_kernel (np.histogram2d
// binning) and herm_eigvec_kernel (Lanczos with re-orthogonalisation).  The
// FFT-based parts (ifft2, Gerchberg-Saxton) are not emulated.  TEST INFRASTRUCTURE.
#define SB_HOST_EMU 1
#include "simt.h"

#include <type_traits>

namespace sb {
alignas(128) unsigned char smem_raw[256 * 1024];
}
#include "../../scintools_b200/csrc/retrieval.cu"

extern "C" int emu_rev_map(const float* thth, int n, const double* th, double eta, double tau0,
                           double dtau, int ntau, double fd0, double dfd, int nfd, int hermitian,
                           float* recov) {
    using namespace sb;
    const size_t bins = (size_t)ntau * nfd;
    std::vector<int> cnt(bins, 0);
    std::memset(recov, 0, bins * sizeof(float2));
    RevGeom g{th, n, eta, tau0, dtau, fd0, dfd, ntau, nfd};
    for (unsigned b = 0; b < 8; ++b)
        emu::run_block(emu::Dim3{256, 1, 1}, emu::Dim3{b, 0, 0}, emu::Dim3{8, 1, 1}, [&]() {
            rev_scatter_kernel(g, reinterpret_cast<const float2*>(thth), hermitian,
                               reinterpret_cast<float2*>(recov), cnt.data());
        });
    for (unsigned b = 0; b < 8; ++b)
        emu::run_block(emu::Dim3{256, 1, 1}, emu::Dim3{b, 0, 0}, emu::Dim3{8, 1, 1}, [&]() {
            rev_finalise_kernel(g, reinterpret_cast<float2*>(recov), cnt.data());
        });
    return 0;
}

extern "C" int emu_herm_eigvec(const float* A, int n, int ld, double tol, int max_iter, double* w,
                               float* V, int* info) {
    using namespace sb;
    std::vector<float2> Q((size_t)(max_iter + 1) * n);
    emu::run_block(emu::Dim3{(unsigned)EV_THREADS, 1, 1}, emu::Dim3{0, 0, 0}, emu::Dim3{1, 1, 1}, [&]() {
        herm_eigvec_kernel(reinterpret_cast<const float2*>(A), n, ld, Q.data(), max_iter, tol, w,
                           reinterpret_cast<float2*>(V), info);
    });
    return 0;
}
