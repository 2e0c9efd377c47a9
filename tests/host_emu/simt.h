// Minimal single-CTA SIMT emulator for tests (TEST INFRASTRUCTURE, not a CPU
// fallback): every CUDA thread of one thread block is a ucontext fiber;
// __syncthreads / __syncwarp / warp shuffles / ballots are rendezvous points
// at which a fiber yields to a round-robin scheduler; bulk copies
// (cp.async.bulk + mbarrier) complete synchronously.  Enough to run kernels
// that use shared memory, barriers, shuffles and the mbarrier ring on the CPU
// and check their arithmetic, indexing and control flow.
#pragma once
#include <ucontext.h>

#include <math.h>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- CUDA vocabulary ------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __restrict__
#define __forceinline__ inline
#define __shared__
#define __align__(x)
#define __launch_bounds__(...)
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline unsigned __float_as_uint(float x) { unsigned u; std::memcpy(&u, &x, 4); return u; }
static inline float __uint_as_float(unsigned u) { float x; std::memcpy(&x, &u, 4); return x; }
static inline double __longlong_as_double(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline int __ffs(unsigned v) { return v ? __builtin_ffs((int)v) : 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
// compile the emulation with -ffp-contract=off: these are single IEEE operations
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
// one fiber runs at a time: plain read-modify-write is atomic here
static inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
using std::isfinite;

namespace emu {
struct Dim3 { unsigned x, y, z; };
struct Fiber { ucontext_t ctx; char* stack = nullptr; bool done = false; };
constexpr size_t STACK_BYTES = 192 * 1024;
// stacks are allocated once and reused by every emulated block (no zero-fill)
inline char* stack_of(int i) {
    static std::vector<char*> pool;
    while ((int)pool.size() <= i) pool.push_back((char*)std::malloc(STACK_BYTES));
    return pool[i];
}
struct Warp {
    unsigned long long slot[32];
    int arrive = 0, arrive2 = 0;
    unsigned gen = 0, gen2 = 0;
};
struct Block {
    int nthreads = 0, cur = 0;
    Dim3 bdim{1, 1, 1}, gdim{1, 1, 1};
    std::vector<Fiber> th;
    ucontext_t sched;
    int bar_count = 0;
    unsigned bar_gen = 0;
    std::vector<Warp> warps;
    Dim3 bidx{0, 0, 0};
    std::function<void()> body;
};
inline Block*& B() { static Block* b = nullptr; return b; }
inline void yield() { Block* b = B(); swapcontext(&b->th[b->cur].ctx, &b->sched); }
inline Dim3 thread_idx() {
    const Block* b = B();
    return Dim3{(unsigned)b->cur % b->bdim.x, ((unsigned)b->cur / b->bdim.x) % b->bdim.y,
                (unsigned)b->cur / (b->bdim.x * b->bdim.y)};
}
inline Dim3 block_idx() { return B()->bidx; }
inline void trampoline() {
    Block* b = B();
    b->body();
    b->th[b->cur].done = true;
    swapcontext(&b->th[b->cur].ctx, &b->sched);
}
// run one thread block of `nthreads` threads executing `body`
inline void run_block(Dim3 bdim, Dim3 bidx, Dim3 gdim, std::function<void()> body);
inline void run_block(int nthreads, unsigned block_x, std::function<void()> body) {
    run_block(Dim3{(unsigned)nthreads, 1, 1}, Dim3{block_x, 0, 0}, Dim3{block_x + 1, 1, 1}, body);
}
inline void run_block(Dim3 bdim, Dim3 bidx, Dim3 gdim, std::function<void()> body) {
    const int nthreads = (int)(bdim.x * bdim.y * bdim.z);
    Block blk;
    blk.bdim = bdim;
    blk.gdim = gdim;
    blk.nthreads = nthreads;
    blk.th.resize(nthreads);
    blk.warps.resize((nthreads + 31) / 32);
    blk.bidx = bidx;
    blk.body = body;
    B() = &blk;
    for (int i = 0; i < nthreads; ++i) {
        Fiber& f = blk.th[i];
        f.stack = stack_of(i);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    for (;;) {
        bool any = false;
        for (int i = 0; i < nthreads; ++i) {
            if (blk.th[i].done) continue;
            any = true;
            blk.cur = i;
            swapcontext(&blk.sched, &blk.th[i].ctx);
        }
        if (!any) break;
    }
    B() = nullptr;
}
inline void block_barrier() {
    Block* b = B();
    const unsigned g = b->bar_gen;
    if (++b->bar_count == b->nthreads) { b->bar_count = 0; b->bar_gen++; }
    else while (b->bar_gen == g) yield();
}
// all 32 lanes of the calling warp post a 64-bit value; `out` receives all of them
inline void warp_gather(unsigned long long mine, unsigned long long* out) {
    Block* b = B();
    Warp& w = b->warps[b->cur >> 5];
    const int lane = b->cur & 31;
    w.slot[lane] = mine;
    const unsigned g = w.gen;
    if (++w.arrive == 32) { w.arrive = 0; w.gen++; }
    else while (w.gen == g) yield();
    for (int i = 0; i < 32; ++i) out[i] = w.slot[i];
    const unsigned g2 = w.gen2;
    if (++w.arrive2 == 32) { w.arrive2 = 0; w.gen2++; }
    else while (w.gen2 == g2) yield();
}
template <typename T> inline unsigned long long to_bits(T v) {
    unsigned long long u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <typename T> inline T from_bits(unsigned long long u) {
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}
}  // namespace emu

#define threadIdx (emu::thread_idx())
#define blockIdx (emu::block_idx())
#define blockDim (emu::B()->bdim)
#define gridDim (emu::B()->gdim)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    unsigned long long all[32];
    emu::warp_gather(0, all);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
    unsigned long long all[32];
    emu::warp_gather(emu::to_bits(v), all);
    return emu::from_bits<T>(all[src & 31]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
    unsigned long long all[32];
    emu::warp_gather(emu::to_bits(v), all);
    return emu::from_bits<T>(all[((int)(emu::B()->cur & 31)) ^ m]);
}
static inline bool __any_sync(unsigned, bool p) {
    unsigned long long all[32];
    emu::warp_gather(p ? 1ull : 0ull, all);
    for (int i = 0; i < 32; ++i) if (all[i]) return true;
    return false;
}
static inline unsigned __ballot_sync(unsigned, bool p) {
    unsigned long long all[32];
    emu::warp_gather(p ? 1ull : 0ull, all);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(all[i] & 1ull) << i;
    return r;
}

// ---- cuda_fp16.h subset: IEEE half, round to nearest even ------------------
struct __half2 { unsigned short x, y; };
static inline unsigned short emu_f2h(float f) {
    const unsigned u = __float_as_uint(f);
    const unsigned sign = (u >> 16) & 0x8000u;
    const int e = (int)((u >> 23) & 0xff) - 127 + 15;
    unsigned m = u & 0x7fffffu;
    if (((u >> 23) & 0xff) == 0xff) return (unsigned short)(sign | 0x7c00u | (m ? 0x200u : 0));
    if (e >= 31) return (unsigned short)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (unsigned short)sign;
        m |= 0x800000u;
        const int shift = 14 - e;
        unsigned h = m >> shift;
        const unsigned rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
        return (unsigned short)(sign | h);
    }
    unsigned h = (unsigned)(e << 10) | (m >> 13);
    const unsigned rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
    return (unsigned short)(sign | h);
}
static inline float emu_h2f(unsigned short h) {
    const unsigned sign = (unsigned)(h & 0x8000u) << 16;
    int e = (h >> 10) & 0x1f;
    unsigned m = h & 0x3ffu;
    if (e == 0) {
        if (!m) return __uint_as_float(sign);
        while (!(m & 0x400u)) { m <<= 1; --e; }
        ++e;
        m &= 0x3ffu;
    } else if (e == 31) {
        return __uint_as_float(sign | 0x7f800000u | (m << 13));
    }
    return __uint_as_float(sign | (unsigned)((e - 15 + 127) << 23) | (m << 13));
}
static inline __half2 __floats2half2_rn(float a, float b) { return __half2{emu_f2h(a), emu_f2h(b)}; }
static inline float2 __half22float2(__half2 h) { return make_float2(emu_h2f(h.x), emu_h2f(h.y)); }
