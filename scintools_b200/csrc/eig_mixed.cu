// Mixed-precision Lanczos for the theta-theta eigenvalue (ththmod.Eval_calc,
// scintools/ththmod.py:371-401) -- ROUND-2 CANDIDATE, OFF BY DEFAULT.
//
//   STATUS: written after the round-1 GPU budget was spent.  It has NOT run on a
//   GPU yet; it does run, unchanged, under the CPU SIMT emulator of
//   tests/host_emu (test_eig_mixed_kernel_on_host: reference eigenvalues to
//   1.5e-7, fp32 restart path included).  Enabled only by SB_EIG_MIXED=1;
//   nothing in the default path calls it; its GPU test is opt-in
//   (SB_TEST_UNVERIFIED=1).
//
// Idea (numerics verified on the CPU, profiles/probe_mixed_precision.py and
// r1_probe_mixed_precision.json): the streaming solver is bound by re-reading
// the fp32 triangle on every Lanczos step.  Run the ITERATION on a bf16 copy of
// the triangle (half the bytes), keep the Lanczos basis in shared memory
// (fp16), form the Ritz vector y, and report the Rayleigh quotient
// <y, A y> / <y, y> with the fp32 triangle in ONE extra pass.  The Rayleigh
// quotient is second order in the vector error: 5e-7 relative error at
// n = 511 (bf16 Ritz value alone: 1.7e-4), same step counts.  Traffic per
// curvature: m * 0.52 MB + 1.04 MB instead of m * 1.04 MB (m ~ 18).
//
// Two variants (template), to be A/B-timed in round 2:
//   <2, false>  SB_EIG_MIXED=1: 2 bf16 stages per warp, basis (24 x fp16) in shared memory
//   <4, true>   SB_EIG_MIXED=2: 4 bf16 stages per warp (as many bytes in flight as the
//               fp32 kernel, twice the elements), basis (64 slots) in global memory (L2)
//
// The bf16 copy is written by thth_build_kernel<true> (thth.cu) next to the fp32
// triangle (+0.5 GB of writes per 1024-eta sweep, no extra pass).
//
// Kernel = the generic TMA row loop of thth_eig_kernel with
//   * 2 x 2 KB bf16 stages per warp (bulk copies from column (a+1) & ~3),
//     the same 4 KB re-used as ONE fp32 stage for the final pass;
//   * up to 24 basis vectors (half2) in shared memory; if the solve needs more
//     steps it restarts in plain fp32 mode and reports the Ritz value, i.e.
//     exactly what thth_eig_kernel does.
#ifndef SB_HOST_EMU
#include <cuda_fp16.h>
#endif
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "bf16_pack.cuh"
#include "lanczos.cuh"
#include "tma.cuh"

namespace sb {

enum { EM_ST_INDEX_ERROR = 1, EM_ST_ZERO_START = 2, EM_ST_TOO_SMALL = 4,
       EM_ST_NOT_CONVERGED = 8 };

constexpr int EM_THREADS = 256;
constexpr int EM_NW = EM_THREADS / 32;
#ifdef SB_EM_NB
constexpr int EM_NB = SB_EM_NB;      // tests/host_emu: few slots to exercise the fp32 restart
constexpr int EM_NBG = SB_EM_NB;
#else
constexpr int EM_NB = 24;            // basis slots in shared memory
constexpr int EM_NBG = 64;           // basis slots in global memory
#endif

template <int NSB, bool BG>
__global__ void __launch_bounds__(EM_THREADS)
thth_eig_mixed_kernel(const float2* __restrict__ Mbase, const unsigned* __restrict__ Mbbase,
                      int ld, const int* __restrict__ nred, int eta0,
                      double* __restrict__ eigs, int* __restrict__ status,
                      int* __restrict__ iters, double tol, double etol, int max_iter,
                      __half2* __restrict__ gbasis) {
    constexpr int NSF = NSB / 2;               // fp32 stages of 4 KB in the same ring
    constexpr int NB = BG ? EM_NBG : EM_NB;    // basis slots
    constexpr int RING = NSB * 2048;           // ring bytes per warp
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LanczosShared& S = *reinterpret_cast<LanczosShared*>(smem_raw);
    float2* v = reinterpret_cast<float2*>(smem_raw + sizeof(LanczosShared));
    float2* vp = v + ld;
    float2* w = vp + ld;          // row sums, then the new Lanczos vector
    float2* u = w + ld;           // column sums
    unsigned char* ring = reinterpret_cast<unsigned char*>(u + ld);     // [NW][RING]
    float2* part = reinterpret_cast<float2*>(ring);                      // [NW][512] scratch (RING >= 4096)
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(ring + (size_t)EM_NW * RING);
    __half2* sbasis = reinterpret_cast<__half2*>(mbar + NSB * EM_NW + 2);   // [EM_NB][ld] if !BG
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = blockIdx.x;
    const int n = nred[eta0 + e];
    const float2* M = Mbase + (size_t)e * ld * ld;
    const unsigned* Mb = Mbbase + (size_t)e * ld * ld;
    __half2* basis = BG ? gbasis + (size_t)e * NB * ld : sbasis;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    if (status[eta0 + e] & EM_ST_INDEX_ERROR) {
        if (tid == 0) { eigs[eta0 + e] = qnan; iters[eta0 + e] = 0; }
        return;
    }
    if (n < 3) {
        if (tid == 0) {
            eigs[eta0 + e] = qnan; iters[eta0 + e] = 0;
            status[eta0 + e] |= EM_ST_TOO_SMALL;
        }
        return;
    }
    if (tid == 0) {
        for (int i = 0; i < NSB * EM_NW; ++i) mbar_init(mbar + i, 1);
        fence_mbarrier_init();
    }
    __syncthreads();
    const int ncol4 = (n + 1) >> 1;            // float4 / uint2 groups = two complex columns
    const int ncolq = ((n + 3) >> 2) << 2;     // bf16 rows are fetched in multiples of 4 columns
    unsigned char* mystage = ring + (size_t)warp * RING;
    unsigned long long* mybar = mbar + NSB * warp;
    unsigned phbits = 0;                       // bit s: phase parity of this warp's barrier s

    // ---- y = A x for the strict upper triangle (row sums into w, column sums
    // into u); BF = true: bf16 rows, NSB stages of 2 KB; false: fp32 rows, NSB/2 stages of 4 KB
    auto matvec = [&](auto bfc) {
        constexpr bool BF = decltype(bfc)::value;
        constexpr int NSTG = BF ? NSB : NSF;
        for (int c = tid; c < ld; c += EM_THREADS) w[c] = make_float2(0.f, 0.f);
        __syncthreads();
        float4 yc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) yc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int K = (n - 2 >= warp) ? (n - 2 - warp) / EM_NW + 1 : 0;
        auto issue = [&](int k) {
            const int a2 = warp + EM_NW * k;
            const int st = k % NSTG;
            if (BF) {
                const int c_lo = (a2 + 1) & ~3;
                const unsigned bytes = (unsigned)(ncolq - c_lo) * 4u;
                mbar_expect_tx(mybar + st, bytes);
                bulk_g2s(mystage + st * 2048 + c_lo * 4, Mb + (size_t)a2 * ld + c_lo, bytes,
                         mybar + st);
            } else {
                const int c_lo = (a2 + 1) & ~1;
                const unsigned bytes = (unsigned)(2 * ncol4 - c_lo) * 8u;
                mbar_expect_tx(mybar + st, bytes);
                bulk_g2s(mystage + st * 4096 + c_lo * 8, M + (size_t)a2 * ld + c_lo, bytes,
                         mybar + st);
            }
        };
        if (lane == 0)
            for (int k = 0; k < NSTG && k < K; ++k) issue(k);
        for (int k = 0; k < K; ++k) {
            const int a = warp + EM_NW * k;
            const int first4 = (a + 1) >> 1;
            const float2 xa = v[a];
            const int st = k % NSTG;
            while (!mbar_try_wait(mybar + st, (phbits >> st) & 1u)) {}
            phbits ^= 1u << st;
            float4 mm[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c4 = lane + 32 * j;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c4 >= first4 && c4 < ncol4) {
                    if (BF) {
                        const uint2 r = reinterpret_cast<const uint2*>(mystage + st * 2048)[c4];
                        q.x = __uint_as_float(r.x << 16);
                        q.y = __uint_as_float(r.x & 0xffff0000u);
                        q.z = __uint_as_float(r.y << 16);
                        q.w = __uint_as_float(r.y & 0xffff0000u);
                    } else {
                        q = reinterpret_cast<const float4*>(mystage + st * 4096)[c4];
                    }
                }
                mm[j] = q;
            }
            float rx = 0.f, ry = 0.f;
            const int jskip = first4 >> 5;      // groups entirely left of the diagonal
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < jskip) continue;
                const int c4 = lane + 32 * j;
                const float4 q = mm[j];
                const float4 x = (2 * c4 < ld) ? *reinterpret_cast<const float4*>(v + 2 * c4)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                rx = fmaf(q.x, x.x, rx); rx = fmaf(-q.y, x.y, rx);
                rx = fmaf(q.z, x.z, rx); rx = fmaf(-q.w, x.w, rx);
                ry = fmaf(q.x, x.y, ry); ry = fmaf(q.y, x.x, ry);
                ry = fmaf(q.z, x.w, ry); ry = fmaf(q.w, x.z, ry);
                // conj(A) * v[a]
                yc[j].x = fmaf(q.x, xa.x, yc[j].x); yc[j].x = fmaf(q.y, xa.y, yc[j].x);
                yc[j].y = fmaf(q.x, xa.y, yc[j].y); yc[j].y = fmaf(-q.y, xa.x, yc[j].y);
                yc[j].z = fmaf(q.z, xa.x, yc[j].z); yc[j].z = fmaf(q.w, xa.y, yc[j].z);
                yc[j].w = fmaf(q.z, xa.y, yc[j].w); yc[j].w = fmaf(-q.w, xa.x, yc[j].w);
            }
            __syncwarp();                       // every lane is done reading the stage
            if (lane == 0 && k + NSTG < K) issue(k + NSTG);
            rx = warp_sum(rx);
            ry = warp_sum(ry);
            if (lane == 0) { w[a].x += rx; w[a].y += ry; }
        }
        __syncthreads();                        // every warp is done with its stages
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(part + warp * 512 + 2 * (lane + 32 * j)) = yc[j];
        __syncthreads();
        for (int c = tid; c < 512; c += EM_THREADS) {
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int kk = 0; kk < EM_NW; ++kk) { sx += part[kk * 512 + c].x; sy += part[kk * 512 + c].y; }
            if (c < ld) u[c] = make_float2(sx, sy);
        }
        // the scratch aliases the stages: order these generic writes before the
        // next bulk copies (async proxy)
        fence_proxy_async();
        __syncthreads();
    };

    int mode = 0;                               // 0: bf16 iteration + basis, 1: plain fp32
    int m = 0;
    for (;;) {
        // v0 = row n//2 of the Hermitian matrix (ththmod.py:398-399), fp32 in both modes
        const int h = n / 2;
        double part0 = 0.0;
        for (int c = tid; c < ld; c += EM_THREADS) {
            float2 x = make_float2(0.f, 0.f);
            if (c < n && c > h) x = M[(size_t)h * ld + c];
            else if (c < h) { x = M[(size_t)c * ld + h]; x.y = -x.y; }
            v[c] = x;
            vp[c] = make_float2(0.f, 0.f);
            part0 += (double)x.x * x.x + (double)x.y * x.y;
        }
        part0 = warp_sum(part0);
        __syncthreads();
        if (lane == 0) S.red[0][warp] = part0;
        if (tid == 0) {
            S.done = 0; S.lo = 0.0; S.theta = 0.0; S.res = 0.0; S.m_lo2 = 0; S.lo2 = 0.0;
            S.next_check = 1; S.beta2[0] = 0.0;
        }
        __syncthreads();
        double nrm2 = 0.0;
        for (int k = 0; k < EM_NW; ++k) nrm2 += S.red[0][k];
        if (!(nrm2 > 0.0) || !isfinite(nrm2)) {
            if (tid == 0) {
                eigs[eta0 + e] = qnan; iters[eta0 + e] = 0;
                status[eta0 + e] |= EM_ST_ZERO_START;
            }
            return;
        }
        {
            const float s = (float)(1.0 / sqrt(nrm2));
            for (int c = tid; c < ld; c += EM_THREADS) { v[c].x *= s; v[c].y *= s; }
        }
        __syncthreads();

        float beta_prev = 0.f;
        bool overflow = false;
        m = 0;
        for (int it = 0; it < max_iter; ++it) {
            if (mode == 0) {
                if (it >= NB) { overflow = true; break; }
                for (int c = tid; c < ld; c += EM_THREADS)
                    basis[(size_t)it * ld + c] = __floats2half2_rn(v[c].x, v[c].y);
                matvec(std::true_type{});
            } else {
                matvec(std::false_type{});
            }
            // ---- alpha = Re <v, A v>
            double apart = 0.0;
            for (int c = tid; c < n; c += EM_THREADS) {
                float2 x = w[c];
                x.x += u[c].x;
                x.y += u[c].y;
                w[c] = x;
                apart += (double)(v[c].x * x.x + v[c].y * x.y);
            }
            apart = warp_sum(apart);
            if (lane == 0) S.red[0][warp] = apart;
            __syncthreads();
            double alpha = 0.0;
            for (int k = 0; k < EM_NW; ++k) alpha += S.red[0][k];
            // ---- w -= alpha v + beta_prev vp ; beta = ||w||
            const float af = (float)alpha;
            double bpart = 0.0;
            for (int c = tid; c < n; c += EM_THREADS) {
                float2 x = w[c];
                x.x -= af * v[c].x + beta_prev * vp[c].x;
                x.y -= af * v[c].y + beta_prev * vp[c].y;
                w[c] = x;
                bpart += (double)x.x * x.x + (double)x.y * x.y;
            }
            bpart = warp_sum(bpart);
            if (lane == 0) S.red[1][warp] = bpart;
            __syncthreads();
            double b2 = 0.0;
            for (int k = 0; k < EM_NW; ++k) b2 += S.red[1][k];
            const double beta = sqrt(b2);
            m = it + 1;
            if (tid == 0) { S.alpha[it] = alpha; S.beta[m] = beta; S.beta2[m] = b2; }
            __syncthreads();
            const bool last = (it + 1 == max_iter);
            if (warp == 0 && (m >= S.next_check || last || !(beta > 0.0)))
                lanczos_check(S, m, tol, etol);
            __syncthreads();
            if (S.done || !isfinite(alpha)) break;
            // ---- rotate: vp = v, v = w / beta
            const float ib = (float)(1.0 / beta);
            for (int c = tid; c < n; c += EM_THREADS) {
                const float2 x = w[c];
                vp[c] = v[c];
                v[c] = make_float2(x.x * ib, x.y * ib);
            }
            beta_prev = (float)beta;
            __syncthreads();
        }
        if (mode == 0 && overflow) {            // more steps than basis slots: redo in fp32
            mode = 1;
            __syncthreads();
            continue;
        }
        break;
    }
    if (mode == 1 || !S.done) {
        // plain result (what thth_eig_kernel reports)
        if (tid == 0) {
            eigs[eta0 + e] = fabs(S.theta);
            iters[eta0 + e] = m;
            if (!S.done) status[eta0 + e] |= EM_ST_NOT_CONVERGED;
        }
        return;
    }
    // ---- Ritz vector of T_m at theta (backward recurrence, grows towards s_0),
    // y = sum_j s_j q_j, eigenvalue = Rayleigh quotient with the fp32 triangle
    if (tid == 0) {
        const double theta = S.theta;
        double* s = S.piv;
        s[m - 1] = 1.0;
        if (m >= 2) s[m - 2] = (S.beta[m - 1] != 0.0) ? (theta - S.alpha[m - 1]) / S.beta[m - 1] : 0.0;
        for (int i = m - 2; i >= 1; --i) {
            const double t = (theta - S.alpha[i]) * s[i] - S.beta[i + 1] * s[i + 1];
            s[i - 1] = (S.beta[i] != 0.0) ? t / S.beta[i] : 0.0;
            if (fabs(s[i - 1]) > 1e150)
                for (int k = i - 1; k < m; ++k) s[k] *= 1e-150;
        }
        double nn = 0.0;
        for (int i = 0; i < m; ++i) nn += s[i] * s[i];
        nn = 1.0 / sqrt(nn);
        for (int i = 0; i < m; ++i) s[i] *= nn;
    }
    __syncthreads();
    for (int c = tid; c < ld; c += EM_THREADS) {
        float sx = 0.f, sy = 0.f;
        if (c < n) {
            for (int j = 0; j < m; ++j) {
                const float2 q = __half22float2(basis[(size_t)j * ld + c]);
                const float sj = (float)S.piv[j];
                sx = fmaf(sj, q.x, sx);
                sy = fmaf(sj, q.y, sy);
            }
        }
        v[c] = make_float2(sx, sy);
    }
    __syncthreads();
    matvec(std::false_type{});
    double num = 0.0, den = 0.0;
    for (int c = tid; c < n; c += EM_THREADS) {
        const float2 y = v[c];
        const float ax = w[c].x + u[c].x, ay = w[c].y + u[c].y;
        num += (double)y.x * ax + (double)y.y * ay;
        den += (double)y.x * y.x + (double)y.y * y.y;
    }
    num = warp_sum(num);
    den = warp_sum(den);
    if (lane == 0) { S.red[0][warp] = num; S.red[1][warp] = den; }
    __syncthreads();
    if (tid == 0) {
        double sn = 0.0, sd = 0.0;
        for (int k = 0; k < EM_NW; ++k) { sn += S.red[0][k]; sd += S.red[1][k]; }
        eigs[eta0 + e] = (sd > 0.0) ? fabs(sn / sd) : fabs(S.theta);
        iters[eta0 + e] = m;
    }
}

#ifndef SB_HOST_EMU
// 0: disabled (default); 1 / 2: the variant selected by SB_EIG_MIXED
int eig_mixed_variant(int ld) {
    const char* ev = getenv("SB_EIG_MIXED");
    const int variant = ev ? atoi(ev) : 0;
    return (variant > 0 && ld <= 512) ? (variant >= 2 ? 2 : 1) : 0;
}

// d_Mb: the bf16 copy of d_M written by thth_build_kernel<true>.  Returns 1.
int eig_mixed_launch(const float2* d_M, const unsigned* d_Mb, int variant, int ld,
                     const int* d_nred, int e0, int nb, double* d_eigs, int* d_status,
                     int* d_iters, double tol, double etol, int max_iter, cudaStream_t st) {
    const size_t base = sizeof(LanczosShared) + 4 * (size_t)ld * sizeof(float2);
    if (variant >= 2) {
        // 4 stages, basis in global memory (L2-resident: 128 KB per curvature)
        __half2* d_basis = (__half2*)workspace(7, (size_t)nb * EM_NBG * ld * sizeof(__half2));
        if (!d_basis) return SB_ERR_NOMEM;
        const size_t smem = base + (size_t)EM_NW * 4 * 2048 + (4 * EM_NW + 2) * sizeof(unsigned long long);
        SB_CUDA(cudaFuncSetAttribute(thth_eig_mixed_kernel<4, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        thth_eig_mixed_kernel<4, true><<<nb, EM_THREADS, smem, st>>>(
            d_M, d_Mb, ld, d_nred, e0, d_eigs, d_status, d_iters, tol, etol, max_iter, d_basis);
    } else {
        const size_t smem = base + (size_t)EM_NW * 2 * 2048 + (2 * EM_NW + 2) * sizeof(unsigned long long) +
                            (size_t)EM_NB * ld * sizeof(__half2);
        SB_CUDA(cudaFuncSetAttribute(thth_eig_mixed_kernel<2, false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        thth_eig_mixed_kernel<2, false><<<nb, EM_THREADS, smem, st>>>(
            d_M, d_Mb, ld, d_nred, e0, d_eigs, d_status, d_iters, tol, etol, max_iter, nullptr);
    }
    SB_LAUNCH_CHECK();
    return 1;
}

#endif  // SB_HOST_EMU

}  // namespace sb
