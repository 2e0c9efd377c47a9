// CPU run of csrc/eig_mixed.cu (thth_pack_bf16_kernel + thth_eig_mixed_kernel)
// under the SIMT emulator of simt.h: 256 fibers per thread block, barriers,
// shuffles and the mbarrier ring emulated, the kernel source unchanged.
// TEST INFRASTRUCTURE (tests/test_host_emulation.py).
#define SB_HOST_EMU 1
#include "simt.h"

#include <float.h>
#include <math.h>

#include <type_traits>

namespace sb {
alignas(128) unsigned char smem_raw[256 * 1024];
}
#include "../../scintools_b200/csrc/eig_mixed.cu"

extern "C" int emu_eig_mixed(const float* M, int ld, const int* nred, int nb, double* eigs,
                             int* status, int* iters, double tol, double etol, int max_iter,
                             int variant) {
    const size_t count = (size_t)nb * ld * ld;
    std::vector<unsigned> Mb(count);
    // pack kernel: no rendezvous, one "thread" covers everything through its grid stride
    emu::run_block(32, 0, [&]() {
        if (threadIdx.x == 0) {
            for (size_t i = 0; i < count; ++i) {
                const float2 v = reinterpret_cast<const float2*>(M)[i];
                Mb[i] = sb::bf16_bits(v.x) | (sb::bf16_bits(v.y) << 16);
            }
        }
    });
    std::vector<__half2> gbasis((size_t)nb * sb::EM_NBG * ld);
    for (int e = 0; e < nb; ++e) {
        std::memset(sb::smem_raw, 0xa5, sizeof(sb::smem_raw));     // garbage, like real shared memory
        emu::run_block(sb::EM_THREADS, (unsigned)e, [&]() {
            if (variant >= 2)
                sb::thth_eig_mixed_kernel<4, true>(reinterpret_cast<const float2*>(M), Mb.data(), ld,
                                                   nred, 0, eigs, status, iters, tol, etol, max_iter,
                                                   gbasis.data());
            else
                sb::thth_eig_mixed_kernel<2, false>(reinterpret_cast<const float2*>(M), Mb.data(), ld,
                                                    nred, 0, eigs, status, iters, tol, etol, max_iter,
                                                    nullptr);
        });
    }
    return 0;
}
