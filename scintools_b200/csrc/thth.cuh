// theta-theta geometry shared by the gather / eigen kernels.
#pragma once
#include "common.cuh"

namespace sb {

// Everything the per-point index math of thth_map needs
// (reference: scintools/ththmod.py:83-107).  Scalars are computed on the
// host with the reference's own numpy expressions so they are bit-identical.
struct ThthGeom {
    const float2* cs;      // conjugate spectrum, fftshifted rows
    long long ntau, nfd;   // logical size
    long long cs_pitch;    // elements per stored row
    int cs_valid_cols;     // half layout: stored columns that hold data (0 = all nfd/2+1)
    const float* cs_bound; // device: upper bound of max |CS| or null (the sweep scans)
    int cs_half;           // 0: full [ntau][nfd]; 1: Hermitian half [ntau][nfd/2+1]
                           //    holding the UNSHIFTED columns k = 0..nfd/2 (fd >= 0)
    double tau0, dtau, half_dtau, tau_absmax;  // tau[0], mean diff, /2, |tau.max()|
    double fd0, dfd, half_dfd, fd_half;        // fd[0], mean diff, /2, |fd.max()|/2
    double inv_dtau, inv_dfd;                  // reciprocals for the fast exact floor
    const double* th;      // theta bin centres (recentred), length n
    int n;
    int coherent;          // 1: complex CS, 0: |CS| (incoherent theta-theta)
};

struct ThthPoint {
    long long tq, fq;  // tau_inv, fd_inv (ththmod.py:94-97)
    bool pnt;          // pnts mask (ththmod.py:100)
    bool index_error;  // numpy would raise IndexError (fd_inv < -nfd)
};

// th1 = theta of the COLUMN, th2 = theta of the ROW (ththmod.py:86-87).
__device__ __forceinline__ ThthPoint thth_point(const ThthGeom& g, double eta,
                                                double th1, double th2) {
    ThthPoint p;
    double d = __dsub_rn(__dmul_rn(th1, th1), __dmul_rn(th2, th2));
    double a = __dadd_rn(__dsub_rn(__dmul_rn(eta, d), g.tau0), g.half_dtau);
    double b = __dadd_rn(__dsub_rn(__dsub_rn(th1, th2), g.fd0), g.half_dfd);
    double tqd = floor_div_fast(a, g.dtau, g.inv_dtau);
    double fqd = floor_div_fast(b, g.dfd, g.inv_dfd);
    // .astype(int): NaN / out-of-range -> INT64_MIN like numpy on x86
    p.tq = (tqd == tqd && fabs(tqd) < 9.0e18) ? (long long)tqd : LLONG_MIN;
    p.fq = (fqd == fqd && fabs(fqd) < 9.0e18) ? (long long)fqd : LLONG_MIN;
    p.pnt = (p.tq > 0) && (p.tq < g.ntau) && (p.fq < g.nfd);
    p.index_error = p.pnt && (p.fq < -g.nfd);
    return p;
}

// Gathered, Jacobian-weighted value before the Hermitian fill
// (ththmod.py:104,107).  Negative fd_inv wraps like python indexing.
__device__ __forceinline__ float2 thth_value(const ThthGeom& g, double eta,
                                             double th1, double th2,
                                             const ThthPoint& p) {
    float2 v = make_float2(0.f, 0.f);
    if (p.pnt && !p.index_error) {
        long long fi = p.fq < 0 ? p.fq + g.nfd : p.fq;
        if (!g.cs_half) {
            v = __ldg(g.cs + (size_t)p.tq * (size_t)g.cs_pitch + (size_t)fi);
        } else {
            // CS of a real dynamic spectrum: CS[-tau, -fd] = conj(CS[tau, fd])
            const long long h = g.nfd / 2;
            long long r = p.tq, c;
            bool cj = false;
            if (fi >= h) c = fi - h;
            else if (fi == 0) c = h;
            else { c = h - fi; r = (g.ntau - p.tq) % g.ntau; cj = true; }
            v = __ldg(g.cs + (size_t)r * (size_t)g.cs_pitch + (size_t)c);
            if (cj) v.y = -v.y;
        }
        if (!g.coherent) v = make_float2(hypotf(v.x, v.y), 0.f);
    }
    // Jacobian sqrt|2 eta (th2 - th1)| (ththmod.py:107); fp32 sqrt is ample
    float wf = sqrtf((float)fabs(2.0 * eta * (th2 - th1)));
    v.x *= wf;
    v.y *= wf;
    return v;
}

// np.nan_to_num on one float component (NaN -> 0, +-inf -> +-FLT_MAX).
__device__ __forceinline__ float nan_to_num(float x) {
    if (x != x) return 0.f;
    if (isinf(x)) return x > 0 ? 3.402823466e+38f : -3.402823466e+38f;
    return x;
}

}  // namespace sb
