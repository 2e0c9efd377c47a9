// Functors and the strided-axis driver shared by the complex 2-D transforms
// (sim.cu: phase screen / propagation, retrieval.cu: inverse of a centred
// conjugate spectrum).  Row kernels: load(row, n), store(row, k, v).  Tile
// kernels (four-step column pass): load(y, i, c), store(y, k, c, v).
#pragma once
#include "fft_kernels.cuh"

namespace sb {

template <typename C> struct PlainRowStore {
    C* out;
    long pitch;
    __device__ __forceinline__ void operator()(long row, int k, C v) const {
        out[row * pitch + k] = v;
    }
};
template <typename C> struct StrideALoad {   // y = r2, i = r1
    const C* in;
    long pitch;
    int R2;
    __device__ __forceinline__ C operator()(int y, int i, int c) const {
        return in[(size_t)(i * R2 + y) * pitch + c];
    }
};
template <typename C> struct TwiddleAStore { // times W_R^(dir y k), row k*R2 + y
    C* out;
    long pitch;
    int R2, R;
    const C* wR;
    __device__ __forceinline__ void operator()(int y, int k, int c, C v) const {
        out[(size_t)(k * R2 + y) * pitch + c] = cmul(v, wR[(y * k) & (R - 1)]);
    }
};
template <typename C> struct BlockBLoad {    // y = k1, i = r2
    const C* in;
    long pitch;
    int R2;
    __device__ __forceinline__ C operator()(int y, int i, int c) const {
        return in[(size_t)(y * R2 + i) * pitch + c];
    }
};
// generic strided-axis transform of a [R][pitch] complex array, functor on
// the final store: storeB(y=k1, k=k2, c, v)
template <typename T, int DIR, class LoadA, class StoreB>
static inline int cols_generic(LoadA la, cx<T>* tmp, long pitch, int R, int ncols,
                        StoreB sb, cudaStream_t st) {
    using C = cx<T>;
    constexpr int W = 256 / sizeof(C);
    int R1, R2;
    split_len(R, &R1, &R2);
    const C* wR = twiddle_table<T>(R, DIR, st);
    if (!wR) return SB_ERR_NOMEM;
    TwiddleAStore<C> sa{tmp, pitch, R2, R, wR};
    int rc = SB_OK;
    SB_TILE_DISPATCH(R1, rc = (launch_tile_fft<T, LL, W, DIR>(la, sa, ncols, R2, st)));
    if (rc) return rc;
    BlockBLoad<C> lb{tmp, pitch, R2};
    SB_TILE_DISPATCH(R2, rc = (launch_tile_fft<T, LL, W, DIR>(lb, sb, ncols, R1, st)));
    return rc;
}

}  // namespace sb
