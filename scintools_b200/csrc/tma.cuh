// cp.async.bulk (TMA, non-tensor) + mbarrier helpers shared by the eigen solvers.
#pragma once

#ifdef SB_HOST_EMU
// tests/host_emu/simt.h: bulk copies complete synchronously; the 8-byte barrier
// word holds the phase parity (bit 0) and the pending transaction bytes (bits 8+),
// so that several copies can complete one phase
namespace sb {
inline void mbar_init(unsigned long long* bar, int) { *bar = 0; }
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    *bar += (unsigned long long)bytes << 8;
}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    std::memcpy(dst, src, bytes);
    *bar -= (unsigned long long)bytes << 8;
    if ((*bar >> 8) == 0) *bar ^= 1ull;
}
inline bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    if ((unsigned)(*bar & 1ull) != parity) return true;
    emu::yield();
    return false;
}
inline void fence_mbarrier_init() {}
inline void fence_proxy_async() {}
// cp.async (LDGSTS): completes synchronously
inline void cp_async16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
// shared-memory byte address: a plain pointer on the host
typedef unsigned char* smem_addr;
inline smem_addr smem_addr_of(void* p) { return (smem_addr)p; }
inline void cp_async16_s(smem_addr dst, const void* src) { std::memcpy(dst, src, 16); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
}  // namespace sb
#else

namespace sb {

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void fence_mbarrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// cp.async (LDGSTS): 16-byte global -> shared copy of the issuing thread, L2 only
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// same with the destination given as a 32-bit shared address
typedef unsigned smem_addr;
__device__ __forceinline__ smem_addr smem_addr_of(void* p) { return smem_u32(p); }
__device__ __forceinline__ void cp_async16_s(smem_addr dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// order generic-proxy shared-memory writes before later bulk copies (async proxy)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace sb
#endif  // SB_HOST_EMU
