// Chirp-z (Bluestein) building blocks shared by the conjugate-spectrum path and
// the any-size inverse 2-D FFT of dynspec.cu: the chirp / kernel tables and the
// load / store functors of the row and column passes.  Barrier-free, so
// tests/host_emu can compile this header for the CPU (SB_HOST_EMU) and check
// the index / conjugation / scaling logic around a reference DFT.
#pragma once
#ifndef SB_HOST_EMU
#include "fft_core.cuh"
#endif

namespace sb {

__global__ void chirp_fill_kernel(float2* w, float2* b, int N, int M) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= M) return;
    // b[m] = conj(w[|m|]) for -N < m < N (wrapped mod M), else 0
    const int m = n < N ? n : (M - n < N ? M - n : -1);
    float2 bv = make_float2(0.f, 0.f);
    if (m >= 0) {
        const long long q = ((long long)m * m) % (2LL * N);
        double s, c;
        sincospi((double)q / (double)N, &s, &c);
        bv = make_float2((float)c, (float)s);            // conj(w) = exp(+i pi m^2/N)
        if (n < N) w[n] = make_float2((float)c, (float)-s);
    }
    b[n] = bv;
}

struct VecLoad {     // single row / column loader
    const float2* v;
    __device__ __forceinline__ float2 operator()(long, int n) const { return v[n]; }
};
struct VecStore {
    float2* v;
    __device__ __forceinline__ void operator()(long, int k, float2 x) const { v[k] = x; }
};
struct ColVecLoad {  // y = r2, i = r1 over a [M][1] array
    const float2* v;
    int R2;
    __device__ __forceinline__ float2 operator()(int y, int i, int) const { return v[i * R2 + y]; }
};
struct ColVecStore {
    float2* v;
    int R1;
    __device__ __forceinline__ void operator()(int y, int k, int, float2 x) const { v[y + R1 * k] = x; }
};

struct ChirpRowLoad {    // a[n] = (x[f][n] - sub) * w[n], zero beyond the live samples
    const float* dyn;
    int nt;
    const float2* w;
    const double* stats;  // non-null: subtract stats[4] (device mean)
    float sub;
    __device__ __forceinline__ float2 operator()(long row, int n) const {
        if (n >= nt) return make_float2(0.f, 0.f);
        const float x = dyn[(size_t)row * nt + n] - (stats ? (float)stats[4] : sub);
        const float2 c = w[n];
        return make_float2(x * c.x, x * c.y);
    }
};
struct MulVecRowStore {  // out[row][k] = v * B[k]
    float2* out;
    long pitch;
    const float2* B;
    __device__ __forceinline__ void operator()(long row, int k, float2 v) const {
        out[row * pitch + k] = cmul(v, B[k]);
    }
};
struct PitchRowLoad {
    const float2* in;
    long pitch;
    __device__ __forceinline__ float2 operator()(long row, int n) const { return in[row * pitch + n]; }
};
struct ChirpOutRowStore {  // Y[row][k] = v * w[k] / M, k < N
    float2* out;
    long pitch;
    const float2* w;
    int N;
    float scale;
    __device__ __forceinline__ void operator()(long row, int k, float2 v) const {
        if (k < N) {
            const float2 r = cmul(v, w[k]);
            out[row * pitch + k] = make_float2(r.x * scale, r.y * scale);
        }
    }
};
struct ChirpColALoad {   // y = r2, i = r1 : Y[row][c] * wF[row], zero beyond live rows
    const float2* Y;
    long pitch;
    int R2, live;
    const float2* w;
    __device__ __forceinline__ float2 operator()(int y, int i, int c) const {
        const int row = i * R2 + y;
        return row < live ? cmul(Y[(size_t)row * pitch + c], w[row]) : make_float2(0.f, 0.f);
    }
};
struct MulVecColStore {  // out[k][c] = v * B[k], k = k1 + R1 k2
    float2* out;
    long pitch;
    int R1;
    const float2* B;
    __device__ __forceinline__ void operator()(int y, int k, int c, float2 v) const {
        const int kk = y + R1 * k;
        out[(size_t)kk * pitch + c] = cmul(v, B[kk]);
    }
};
struct PlainColALoad {
    const float2* in;
    long pitch;
    int R2;
    __device__ __forceinline__ float2 operator()(int y, int i, int c) const {
        return in[(size_t)(i * R2 + y) * pitch + c];
    }
};
struct ChirpCsStore {    // CS[(k + N/2) % N][(c + NT/2) % NT] = v * wF[k] / M (+dc), masks
    float2* CS;
    int NF, NT, R1;
    const float2* w;
    float scale;
    const unsigned char* rowmask;
    float dc;
    const double* dc_stats;
    __device__ __forceinline__ void operator()(int y, int k, int c, float2 v) const {
        const int kf = y + R1 * k;
        if (kf >= NF) return;
        float2 r = cmul(v, w[kf]);
        r.x *= scale;
        r.y *= scale;
        if (kf == 0 && c == 0)
            r.x += dc_stats ? (float)(dc_stats[4] * (double)NF * (double)NT) : dc;
        const int rs = (kf + NF / 2) % NF, cs = (c + NT / 2) % NT;
        if (rowmask && rowmask[rs]) r = make_float2(0.f, 0.f);
        CS[(size_t)rs * NT + cs] = r;
    }
};

struct ChirpRowLoadC {   // a[n] = conj(x[r'][n']) * w[n] with the ifftshift folded in
    const float2* in;
    int n0, n1, centred, keep;     // keep != 0: transform conj(in), i.e. no negation here
    const float2* w;
    __device__ __forceinline__ float2 operator()(long row, int n) const {
        if (n >= n1) return make_float2(0.f, 0.f);
        const int r = centred ? (int)((row + n0 / 2) % n0) : (int)row;
        const int c = centred ? (n + n1 / 2) % n1 : n;
        float2 v = in[(size_t)r * n1 + c];
        if (!keep) v.y = -v.y;
        return cmul(v, w[n]);
    }
};
struct ChirpCropStore {  // out[k][c] = conj(v * wF[k]) * scale, k < crop0, c < crop1
    float2* outc;
    float* outr;
    int R1, crop0, crop1;
    const float2* w;
    float scale;
    __device__ __forceinline__ void operator()(int y, int k, int c, float2 v) const {
        const int kf = y + R1 * k;
        if (kf >= crop0 || c >= crop1) return;
        const float2 r = cmul(v, w[kf]);
        const size_t o = (size_t)kf * crop1 + c;
        if (outr) outr[o] = r.x * scale;
        else outc[o] = make_float2(r.x * scale, -r.y * scale);
    }
};

}  // namespace sb
