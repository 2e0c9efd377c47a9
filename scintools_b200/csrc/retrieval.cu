// Phase retrieval building blocks:
//   rev_map      ththmod.py:176-258   theta-theta -> conjugate spectrum (scatter)
//   herm_eigvec  ththmod.py:300-307   top eigenpair of the reduced theta-theta
//                                     matrix (eigsh(.., 1, which='LA') in modeler)
//   ifft2        ththmod.py:321, 1462 ifft2(ifftshift(recov)), cropped
// used by the Python mirrors of modeler / single_chunk_retrieval.
#include <float.h>
#include <math.h>
#include <stdlib.h>

#ifndef SB_HOST_EMU            // tests/host_emu runs the scatter / eigenpair kernels on the CPU
#include "fft_generic.cuh"
#endif
#include "lanczos.cuh"

namespace sb {

// --------------------------------------------------------------------------
// rev_map.  np.histogram2d with explicit edges e_k = (k - 0.5) * d + x0
// (k = 0..N, the same two roundings as numpy): bin = #{e_k <= x} - 1, x == e_N
// belongs to the last bin, anything outside is dropped.  Bit-exact bins; the
// weighted sums are fp32 atomics (order-dependent in the last bits).
// --------------------------------------------------------------------------
__device__ __forceinline__ double hist_edge(int k, double x0, double d) {
    return __dadd_rn(__dmul_rn((double)k - 0.5, d), x0);
}
__device__ __forceinline__ int hist_bin(double x, double x0, double d, int N) {
    if (!(x == x)) return -1;
    const double g = (x - x0) / d + 0.5;
    if (!(g > -2.0) || !(g < (double)N + 2.0)) return -1;
    int k = (int)floor(g);
    k = max(0, min(k, N));
    while (k < N && hist_edge(k + 1, x0, d) <= x) ++k;
    while (k >= 0 && hist_edge(k, x0, d) > x) --k;
    if (k < 0) return -1;
    if (k == N) return (x == hist_edge(N, x0, d)) ? N - 1 : -1;
    return k;
}

struct RevGeom {
    const double* th;
    int n;
    double eta, tau0, dtau, fd0, dfd;
    int ntau, nfd;
};

__global__ void rev_scatter_kernel(RevGeom g, const float2* __restrict__ thth, int hermitian,
                                   float2* __restrict__ acc, int* __restrict__ cnt) {
    const long total = (long)g.n * g.n;
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < total;
         p += (long)gridDim.x * blockDim.x) {
        const int i = (int)(p / g.n), j = (int)(p - (long)i * g.n);
        if (i == j) continue;   // zero Jacobian: the DC bin is NaN -> 0 (finalise kernel)
        const double ti = g.th[i], tj = g.th[j];
        // fd_map[i][j] = th[j] - th[i];  tau_map = eta * (th[j]^2 - th[i]^2)
        const double x = __dsub_rn(tj, ti);
        const double y = __dmul_rn(g.eta, __dsub_rn(__dmul_rn(tj, tj), __dmul_rn(ti, ti)));
        const double jac = sqrt(fabs(__dmul_rn(__dmul_rn(2.0, g.eta), __dsub_rn(ti, tj))));
        const float2 v = thth[p];
        const float wre = (float)((double)v.x / jac), wim = (float)((double)v.y / jac);
        int bx = hist_bin(x, g.fd0, g.dfd, g.nfd), by = hist_bin(y, g.tau0, g.dtau, g.ntau);
        if (bx >= 0 && by >= 0) {
            const size_t o = (size_t)by * g.nfd + bx;
            atomicAdd(&acc[o].x, wre);
            atomicAdd(&acc[o].y, wim);
            atomicAdd(&cnt[o], 1);
        }
        if (hermitian) {
            bx = hist_bin(-x, g.fd0, g.dfd, g.nfd);
            by = hist_bin(-y, g.tau0, g.dtau, g.ntau);
            if (bx >= 0 && by >= 0) {
                const size_t o = (size_t)by * g.nfd + bx;
                atomicAdd(&acc[o].x, wre);
                atomicAdd(&acc[o].y, -wim);
                atomicAdd(&cnt[o], 1);
            }
        }
    }
}

__global__ void rev_finalise_kernel(RevGeom g, float2* __restrict__ acc,
                                    const int* __restrict__ cnt) {
    const long total = (long)g.ntau * g.nfd;
    // the bin of (fd, tau) = (0, 0) receives the n diagonal points with an
    // infinite / NaN weight: NaN after the division, 0 after nan_to_num
    const int bx0 = hist_bin(0.0, g.fd0, g.dfd, g.nfd), by0 = hist_bin(0.0, g.tau0, g.dtau, g.ntau);
    const long dc = (bx0 >= 0 && by0 >= 0) ? (long)by0 * g.nfd + bx0 : -1;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total;
         o += (long)gridDim.x * blockDim.x) {
        const int c = cnt[o];
        float2 v = acc[o];
        if (c > 0 && o != dc) {
            const float s = 1.0f / (float)c;
            v.x *= s;
            v.y *= s;
            // np.nan_to_num(recov): NaN -> 0, +-inf -> +-largest float
            v.x = (v.x != v.x) ? 0.f : fminf(fmaxf(v.x, -FLT_MAX), FLT_MAX);
            v.y = (v.y != v.y) ? 0.f : fminf(fmaxf(v.y, -FLT_MAX), FLT_MAX);
        } else {
            v = make_float2(0.f, 0.f);
        }
        acc[o] = v;
    }
}

#ifndef SB_HOST_EMU
int rev_map(const float2* thth, int n, const double* th_dev, double eta, double tau0,
            double dtau, int ntau, double fd0, double dfd, int nfd, int hermitian,
            float2* recov, cudaStream_t st) {
    if (!(dtau > 0.0) || !(dfd > 0.0)) {
        set_error("rev_map needs ascending tau / fd axes (bins must increase monotonically)");
        return SB_ERR_ARG;
    }
    const size_t bins = (size_t)ntau * nfd;
    int* cnt = (int*)workspace(1, bins * sizeof(int));
    if (!cnt) return SB_ERR_NOMEM;
    SB_CUDA(cudaMemsetAsync(recov, 0, bins * sizeof(float2), st));
    SB_CUDA(cudaMemsetAsync(cnt, 0, bins * sizeof(int), st));
    RevGeom g{th_dev, n, eta, tau0, dtau, fd0, dfd, ntau, nfd};
    const long total = (long)n * n;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    rev_scatter_kernel<<<blocks, 256, 0, st>>>(g, thth, hermitian, recov, cnt);
    SB_LAUNCH_CHECK();
    int fb = (int)((bins + 255) / 256);
    if (fb > 148 * 16) fb = 148 * 16;
    rev_finalise_kernel<<<fb, 256, 0, st>>>(g, recov, cnt);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

#endif  // SB_HOST_EMU

// --------------------------------------------------------------------------
// Top eigenpair (largest algebraic) of one full Hermitian complex matrix:
// Lanczos with the basis kept in global memory and classical Gram-Schmidt
// re-orthogonalisation (twice), Ritz vector from the backward three-term
// recurrence of T_m at theta.  One CTA; used once per chunk.
// --------------------------------------------------------------------------
constexpr int EV_THREADS = 512;
constexpr int EV_NW = EV_THREADS / 32;

__device__ __forceinline__ double ev_block_sum(double x, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    double s = 0.0;
    for (int k = 0; k < EV_NW; ++k) s += red[k];
    return s;
}

__global__ void __launch_bounds__(EV_THREADS)
herm_eigvec_kernel(const float2* __restrict__ A, int n, int ld, float2* __restrict__ Q,
                   int max_iter, double tol, double* __restrict__ w_out,
                   float2* __restrict__ V_out, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LanczosShared& S = *reinterpret_cast<LanczosShared*>(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw + sizeof(LanczosShared));   // [EV_NW]
    double2* coef = reinterpret_cast<double2*>(red + 32);                        // [max_iter + 1]
    float2* v = reinterpret_cast<float2*>(coef + SB_LANCZOS_MAXIT + 1);
    float2* w = v + n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    // start vector: row n//2, like Eval_calc (any vector with a component
    // along the top eigenvector would do; eigsh uses a random one)
    const int h = n / 2;
    double p0 = 0.0;
    for (int c = tid; c < n; c += EV_THREADS) {
        const float2 x = A[(size_t)h * ld + c];
        v[c] = x;
        p0 += (double)x.x * x.x + (double)x.y * x.y;
    }
    if (tid == 0) {
        S.done = 0; S.lo = 0.0; S.theta = 0.0; S.res = 0.0; S.next_check = 1; S.m_last = 0; S.beta2[0] = 0.0;
        S.m_lo2 = 0; S.lo2 = 0.0;
    }
    const double nrm2 = ev_block_sum(p0, red);
    if (!(nrm2 > 0.0) || !isfinite(nrm2) || n < 2) {
        if (tid == 0) { *w_out = qnan; info[0] = 0; info[1] = 2; }
        for (int c = tid; c < n; c += EV_THREADS) V_out[c] = make_float2(0.f, 0.f);
        return;
    }
    {
        const float s = (float)(1.0 / sqrt(nrm2));
        for (int c = tid; c < n; c += EV_THREADS) { v[c].x *= s; v[c].y *= s; }
    }
    __syncthreads();
    int m = 0;
    float beta_prev = 0.f;
    for (int it = 0; it < max_iter; ++it) {
        for (int c = tid; c < n; c += EV_THREADS) Q[(size_t)it * n + c] = v[c];
        // w = A v : one warp per row, lanes across the columns
        for (int a = warp; a < n; a += EV_NW) {
            const float2* row = A + (size_t)a * ld;
            float rx = 0.f, ry = 0.f;
            for (int c = lane; c < n; c += 32) {
                const float2 q = row[c], x = v[c];
                rx = fmaf(q.x, x.x, rx); rx = fmaf(-q.y, x.y, rx);
                ry = fmaf(q.x, x.y, ry); ry = fmaf(q.y, x.x, ry);
            }
            rx = warp_sum(rx);
            ry = warp_sum(ry);
            if (lane == 0) w[a] = make_float2(rx, ry);
        }
        __syncthreads();
        double ap = 0.0;
        for (int c = tid; c < n; c += EV_THREADS)
            ap += (double)v[c].x * w[c].x + (double)v[c].y * w[c].y;
        const double alpha = ev_block_sum(ap, red);
        {
            const float af = (float)alpha;
            const float2* qp = Q + (size_t)(it > 0 ? it - 1 : 0) * n;
            for (int c = tid; c < n; c += EV_THREADS) {
                float2 x = w[c];
                const float2 pv = it > 0 ? qp[c] : make_float2(0.f, 0.f);
                x.x -= af * v[c].x + beta_prev * pv.x;
                x.y -= af * v[c].y + beta_prev * pv.y;
                w[c] = x;
            }
        }
        __syncthreads();
        // classical Gram-Schmidt against Q[0..it], twice: warp j computes <Q_j, w>
        for (int pass = 0; pass < 2; ++pass) {
            for (int j0 = 0; j0 <= it; j0 += EV_NW) {
                const int j = j0 + warp;
                if (j <= it) {
                    const float2* q = Q + (size_t)j * n;
                    double cx_ = 0.0, cy_ = 0.0;
                    for (int c = lane; c < n; c += 32) {
                        const float2 a = q[c], b = w[c];   // conj(a) * b
                        cx_ += (double)a.x * b.x + (double)a.y * b.y;
                        cy_ += (double)a.x * b.y - (double)a.y * b.x;
                    }
                    cx_ = warp_sum(cx_);
                    cy_ = warp_sum(cy_);
                    if (lane == 0) coef[j] = make_double2(cx_, cy_);
                }
            }
            __syncthreads();
            for (int c = tid; c < n; c += EV_THREADS) {
                double sx = 0.0, sy = 0.0;
                for (int j = 0; j <= it; ++j) {
                    const float2 q = Q[(size_t)j * n + c];
                    const double2 cf = coef[j];
                    sx += cf.x * q.x - cf.y * q.y;
                    sy += cf.x * q.y + cf.y * q.x;
                }
                w[c].x -= (float)sx;
                w[c].y -= (float)sy;
            }
            __syncthreads();
        }
        double bp = 0.0;
        for (int c = tid; c < n; c += EV_THREADS) bp += (double)w[c].x * w[c].x + (double)w[c].y * w[c].y;
        const double b2 = ev_block_sum(bp, red);
        const double beta = sqrt(b2);
        m = it + 1;
        if (tid == 0) { S.alpha[it] = alpha; S.beta[m] = beta; S.beta2[m] = b2; }
        __syncthreads();
        if (warp == 0) lanczos_check(S, m, tol, 0.0);
        __syncthreads();
        if (S.done || !isfinite(alpha) || m == max_iter || m == n) break;
        const float ib = (float)(1.0 / beta);
        for (int c = tid; c < n; c += EV_THREADS) v[c] = make_float2(w[c].x * ib, w[c].y * ib);
        beta_prev = (float)beta;
        __syncthreads();
    }
    // Ritz vector of T_m at theta, backward recurrence (grows towards s_0)
    if (tid == 0) {
        const double theta = S.theta;
        double* s = S.piv;
        s[m - 1] = 1.0;
        if (m >= 2) s[m - 2] = (S.beta[m - 1] != 0.0) ? (theta - S.alpha[m - 1]) / S.beta[m - 1] : 0.0;
        for (int i = m - 2; i >= 1; --i) {
            double t = (theta - S.alpha[i]) * s[i] - S.beta[i + 1] * s[i + 1];
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
    double yp = 0.0;
    for (int c = tid; c < n; c += EV_THREADS) {
        double sx = 0.0, sy = 0.0;
        for (int j = 0; j < m; ++j) {
            const float2 q = Q[(size_t)j * n + c];
            sx += S.piv[j] * q.x;
            sy += S.piv[j] * q.y;
        }
        w[c] = make_float2((float)sx, (float)sy);
        yp += sx * sx + sy * sy;
    }
    const double yn = ev_block_sum(yp, red);
    const float ys = (float)(1.0 / sqrt(yn));
    for (int c = tid; c < n; c += EV_THREADS) V_out[c] = make_float2(w[c].x * ys, w[c].y * ys);
    if (tid == 0) {
        *w_out = S.theta;
        info[0] = m;
        // the requested residual (1e-7 by default) is close to the fp32 rounding floor; a
        // residual <= 2e-6 |theta| at the iteration cap is still a converged pair for every
        // consumer (eigenvalue error ~ res^2 / gap), anything worse is flagged
        info[1] = (S.done || S.res <= 2e-6 * fabs(S.theta)) ? 0 : 8;
    }
}

#ifndef SB_HOST_EMU
int herm_eigvec(const float2* A, int n, int ld, double tol, int max_iter, double* w_dev,
                float2* V_dev, int* info_dev, cudaStream_t st) {
    if (!(tol > 0.0)) tol = 1e-7;
    if (max_iter <= 0 || max_iter > SB_LANCZOS_MAXIT) max_iter = 96;
    if (max_iter > n) max_iter = n;
    if (n < 1 || n > 8192) {
        set_error("herm_eigvec: n = %d outside 1..8192", n);
        return SB_ERR_UNSUPPORTED;
    }
    float2* Q = (float2*)workspace(2, (size_t)(max_iter + 1) * n * sizeof(float2));
    if (!Q) return SB_ERR_NOMEM;
    const size_t smem = sizeof(LanczosShared) + 32 * sizeof(double) +
                        (SB_LANCZOS_MAXIT + 1) * sizeof(double2) + 2 * (size_t)n * sizeof(float2);
    SB_CUDA(cudaFuncSetAttribute(herm_eigvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    herm_eigvec_kernel<<<1, EV_THREADS, smem, st>>>(A, n, ld, Q, max_iter, tol, w_dev, V_dev,
                                                    info_dev);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

// --------------------------------------------------------------------------
// out[:crop0, :crop1] = scale * ifft2(ifftshift(in)) for power-of-two sizes.
// Rows (contiguous axis) in shared memory with the ifftshift folded into the
// load, then the strided axis in four-step tiles over the kept columns only.
// --------------------------------------------------------------------------
struct ShiftedRowLoad {   // row r, sample n of ifftshift(in)
    const float2* in;
    int n0, n1, centred;
    __device__ __forceinline__ float2 operator()(long row, int n) const {
        const int r = centred ? (int)((row + n0 / 2) & (n0 - 1)) : (int)row;
        const int c = centred ? ((n + n1 / 2) & (n1 - 1)) : n;
        return in[(size_t)r * n1 + c];
    }
};
struct CropStore {        // out[k][c], k = k1 + R1 k2
    float2* outc;
    float* outr;
    int R1, crop0, crop1;
    float scale;
    __device__ __forceinline__ void operator()(int y, int k, int c, float2 v) const {
        const int row = y + R1 * k;
        if (row >= crop0 || c >= crop1) return;
        const size_t o = (size_t)row * crop1 + c;
        if (outr) outr[o] = v.x * scale;
        else outc[o] = make_float2(v.x * scale, v.y * scale);
    }
};

// dynspec.cu: chirp-z variant for arbitrary sizes
int ifft2_c2c_any(const float2* in, int n0, int n1, int centred, int crop0, int crop1,
                  double scale, int real_only, void* out, cudaStream_t st, int conj_in);

int ifft2_c2c(const float2* in, int n0, int n1, int centred, int crop0, int crop1,
              double scale, int real_only, void* out, cudaStream_t st) {
    if ((n0 & (n0 - 1)) || (n1 & (n1 - 1)))
        return ifft2_c2c_any(in, n0, n1, centred, crop0, crop1, scale, real_only, out, st, 0);
    if (n0 < 8 || n1 < 8 || (n0 & (n0 - 1)) || (n1 & (n1 - 1)) || n0 > 65536 || n1 > 32768) {
        set_error("ifft2: sizes %d x %d must be powers of two (8..65536 x 8..32768)", n0, n1);
        return SB_ERR_UNSUPPORTED;
    }
    if (crop0 <= 0 || crop0 > n0) crop0 = n0;
    if (crop1 <= 0 || crop1 > n1) crop1 = n1;
    float2* B1 = (float2*)workspace(3, (size_t)n0 * n1 * sizeof(float2));
    float2* B2 = (float2*)workspace(4, (size_t)n0 * n1 * sizeof(float2));
    if (!B1 || !B2) return SB_ERR_NOMEM;
    ShiftedRowLoad ld{in, n0, n1, centred};
    PlainRowStore<float2> rs{B1, n1};
    int rc = SB_OK;
    SB_ROW_DISPATCH(n1, rc = (launch_row_c2c<float, N1, N2, +1>(ld, rs, n0, st)));
    if (rc) return rc;
    int R1, R2;
    split_len(n0, &R1, &R2);
    StrideALoad<float2> la{B1, n1, R2};
    CropStore cs{real_only ? nullptr : (float2*)out, real_only ? (float*)out : nullptr, R1,
                 crop0, crop1, (float)(scale / ((double)n0 * (double)n1))};
    return cols_generic<float, +1>(la, B2, n1, n0, crop1, cs, st);
}

// --------------------------------------------------------------------------
// Gerchberg-Saxton iterations on a wavefield (Dynspec.gerchberg_saxton,
// dynspec.py:1883-1896): fft2 -> zero the tau < 0 rows -> ifft2 -> put the
// measured amplitude back where it is known.  fftshift / ifftshift cancel, so
// the causality mask is applied to the unshifted rows; mask and amplitude are
// fused into the final stores of the two column passes.
// --------------------------------------------------------------------------
struct PitchRowLoadC {
    const float2* in;
    long pitch;
    __device__ __forceinline__ float2 operator()(long row, int n) const {
        return in[(size_t)row * pitch + n];
    }
};
struct RowMaskStore {     // out[k][c] = rowmask[k] ? 0 : v,  k = k1 + R1 k2
    float2* out;
    long pitch;
    int R1;
    const unsigned char* rowmask;
    __device__ __forceinline__ void operator()(int y, int k, int c, float2 v) const {
        const int row = y + R1 * k;
        out[(size_t)row * pitch + c] = rowmask[row] ? make_float2(0.f, 0.f) : v;
    }
};
struct AmplitudeStore {   // w = v / (n0 n1); where amp is not NaN: amp * exp(i angle(w))
    float2* out;
    long pitch;
    int R1;
    const float* amp;
    float scale;
    __device__ __forceinline__ void operator()(int y, int k, int c, float2 v) const {
        const int row = y + R1 * k;
        const size_t o = (size_t)row * pitch + c;
        float2 w = make_float2(v.x * scale, v.y * scale);
        const float a = amp[o];
        if (a == a) {
            const float m = hypotf(w.x, w.y);
            w = (m > 0.f) ? make_float2(a * (w.x / m), a * (w.y / m)) : make_float2(a, 0.f);
        }
        out[o] = w;
    }
};

// ---- any-size variant on the chirp-z inverse (round-2 candidate, unverified):
//   T = ifft2(conj W) = conj(fft2 W) / N;  zero the masked rows of T;
//   W = N * ifft2(conj T) = ifft2(masked fft2 W);  amplitude step.
__global__ void gs_rowmask_kernel(float2* T, const unsigned char* __restrict__ rowmask,
                                  long n0, long n1) {
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < n0 * n1;
         o += (long)gridDim.x * blockDim.x)
        if (rowmask[o / n1]) T[o] = make_float2(0.f, 0.f);
}
__global__ void gs_amplitude_kernel(float2* W, const float* __restrict__ amp, long count) {
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < count;
         o += (long)gridDim.x * blockDim.x) {
        const float a = amp[o];
        if (a == a) {
            const float2 w = W[o];
            const float m = hypotf(w.x, w.y);
            W[o] = (m > 0.f) ? make_float2(a * (w.x / m), a * (w.y / m)) : make_float2(a, 0.f);
        }
    }
}
static int gerchberg_saxton_any(float2* W, const float* amp, const unsigned char* rowmask,
                                int n0, int n1, int niter, cudaStream_t st) {
    const long count = (long)n0 * n1;
    float2* T = (float2*)workspace(2, (size_t)count * sizeof(float2));
    if (!T) return SB_ERR_NOMEM;
    int blocks = (int)((count + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    for (int it = 0; it < niter; ++it) {
        int rc = ifft2_c2c_any(W, n0, n1, 0, 0, 0, 1.0, 0, T, st, 1);
        if (rc) return rc;
        gs_rowmask_kernel<<<blocks, 256, 0, st>>>(T, rowmask, n0, n1);
        SB_LAUNCH_CHECK();
        rc = ifft2_c2c_any(T, n0, n1, 0, 0, 0, (double)count, 0, W, st, 1);
        if (rc) return rc;
        gs_amplitude_kernel<<<blocks, 256, 0, st>>>(W, amp, count);
        SB_LAUNCH_CHECK();
    }
    return SB_OK;
}

int gerchberg_saxton(float2* W, const float* amp, const unsigned char* rowmask, int n0, int n1,
                     int niter, cudaStream_t st) {
    if ((n0 & (n0 - 1)) || (n1 & (n1 - 1)))
        return gerchberg_saxton_any(W, amp, rowmask, n0, n1, niter, st);
    if (n0 < 8 || n1 < 8 || (n0 & (n0 - 1)) || (n1 & (n1 - 1)) || n0 > 65536 || n1 > 32768) {
        set_error("gerchberg_saxton: wavefield %d x %d must have power-of-two sizes", n0, n1);
        return SB_ERR_UNSUPPORTED;
    }
    const size_t bytes = (size_t)n0 * n1 * sizeof(float2);
    float2* B1 = (float2*)workspace(3, bytes);
    float2* B2 = (float2*)workspace(4, bytes);
    float2* B3 = (float2*)workspace(5, bytes);
    if (!B1 || !B2 || !B3) return SB_ERR_NOMEM;
    int R1, R2;
    split_len(n0, &R1, &R2);
    for (int it = 0; it < niter; ++it) {
        int rc = SB_OK;
        PitchRowLoadC l0{W, n1};
        PlainRowStore<float2> s1{B1, n1};
        SB_ROW_DISPATCH(n1, rc = (launch_row_c2c<float, N1, N2, -1>(l0, s1, n0, st)));
        if (rc) return rc;
        StrideALoad<float2> la{B1, n1, R2};
        RowMaskStore ms{B3, n1, R1, rowmask};
        rc = cols_generic<float, -1>(la, B2, n1, n0, n1, ms, st);
        if (rc) return rc;
        PitchRowLoadC l3{B3, n1};
        SB_ROW_DISPATCH(n1, rc = (launch_row_c2c<float, N1, N2, +1>(l3, s1, n0, st)));
        if (rc) return rc;
        AmplitudeStore as{W, n1, R1, amp, (float)(1.0 / ((double)n0 * (double)n1))};
        rc = cols_generic<float, +1>(la, B2, n1, n0, n1, as, st);
        if (rc) return rc;
    }
    return SB_OK;
}

#endif  // SB_HOST_EMU

}  // namespace sb
