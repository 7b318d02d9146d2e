// common.cuh — shared device helpers.  The kernels are written against a small CUDA subset
// (thread indices, warp shuffles, __syncwarp, dynamic shared memory) so that tests/emu can compile the very
// same sources for the host and step the warp-level algorithms lane by lane (QMPC_EMU; test-only, never
// part of the product build).
#pragma once

#ifdef QMPC_EMU
#include "emu_cuda.h"
#else
#include <cuda_runtime.h>
#define QMPC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define QMPC_STATIC_SMEM(type, name, count) __shared__ type name[count]
#endif

namespace qmpc {

constexpr unsigned FULL = 0xffffffffu;

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(FULL, v, o); v = w < v ? w : v; }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(FULL, v, o); v = w > v ? w : v; }
    return v;
}
// sum over the 16 lanes of the calling half-warp (hmask = that half's lanes)
template <typename T>
__device__ __forceinline__ T half_sum(unsigned hmask, T v)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(hmask, v, o);
    return v;
}

// bring one 128-byte line into L1 ahead of use (no-op in the host emulation)
__device__ __forceinline__ void prefetch_l1(const void* p)
{
#ifndef QMPC_EMU
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ---- TMA bulk copy global -> shared memory with an mbarrier (cp.async.bulk, SASS UBLKCP): one thread issues the copy,
// every consumer waits on the barrier's phase.  In the host emulation the issuing thread copies synchronously and the
// wait is a block / warp barrier (bulk_wait_block / bulk_wait_warp).
#ifndef QMPC_EMU
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// order this thread's earlier generic-proxy accesses to shared memory before async-proxy (TMA) writes to it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one copy of `bytes` (multiple of 16, both addresses 16-byte aligned) that completes on `bar` (byte count announced
// with mbar_expect beforehand)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global counterpart (bulk async-group completion): the issuing thread commits and, before the buffer is reused or
// the kernel ends, waits until the source has been read.  Writers of the buffer run fence_proxy_async() first.
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QMPC_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra QMPC_MBAR_DONE;\n"
        "bra QMPC_MBAR_WAIT;\n"
        "QMPC_MBAR_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#else
__device__ __forceinline__ void mbar_init(unsigned long long*, int) {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_expect(unsigned long long*, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { std::memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) { std::memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_s2g_wait_read() {}
__device__ __forceinline__ void mbar_wait(unsigned long long*, unsigned) {}
#endif
// consumers' side: every thread of the CTA (resp. lane of the warp) calls this after the issuing thread has issued
__device__ __forceinline__ void bulk_wait_block(unsigned long long* bar, unsigned parity)
{
#ifdef QMPC_EMU
    (void)bar; (void)parity;
    __syncthreads();
#else
    mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void bulk_wait_warp(unsigned long long* bar, unsigned parity)
{
#ifdef QMPC_EMU
    (void)bar; (void)parity;
    __syncwarp();
#else
    mbar_wait(bar, parity);
#endif
}

// ---- completion flags between the warps of one CTA (a producer warp publishes shared-memory data, consumer warps wait
// for it without a CTA barrier): an mbarrier that every lane of the producer warp arrives on once per generation
// (arrive = release, try_wait = acquire at CTA scope; the waiters suspend in hardware).  Generation g completes phase g,
// whose parity is g & 1.  Host emulation: a 64-bit arrival counter, 32 arrivals per generation.  (A generation counter
// polled with ld.acquire.cta instead of the mbarrier was measured 5 % slower: the pollers take issue slots.)
__device__ __forceinline__ void flag_init(unsigned long long* bar)
{
#ifdef QMPC_EMU
    __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST);
#else
    mbar_init(bar, 32);
#endif
}
__device__ __forceinline__ void flag_arrive(unsigned long long* bar)
{
#ifdef QMPC_EMU
    std::atomic_ref<unsigned long long> a(*bar);
    a.fetch_add(1ull);
    a.notify_all();
#else
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
#endif
}
__device__ __forceinline__ void flag_wait(unsigned long long* bar, unsigned generation)
{
#ifdef QMPC_EMU
    std::atomic_ref<unsigned long long> a(*bar);        // blocks in the kernel (futex) instead of spinning: 256 OS threads per CTA
    for (unsigned long long v = a.load(); v < 32ull * (generation + 1ull); v = a.load()) a.wait(v);
#else
    mbar_wait(bar, generation & 1u);
#endif
}

template <typename real> __device__ __forceinline__ real rrsqrt(real x);
#ifdef QMPC_EMU
template <> __device__ __forceinline__ double rrsqrt<double>(double x) { return 1.0 / sqrt(x); }
template <> __device__ __forceinline__ float rrsqrt<float>(float x) { return 1.0f / sqrtf(x); }
#else
template <> __device__ __forceinline__ double rrsqrt<double>(double x) { return rsqrt(x); }
template <> __device__ __forceinline__ float rrsqrt<float>(float x) { return 1.0f / sqrtf(x); }
#endif

#ifndef QMPC_RSQRT_NOBRANCH
#define QMPC_RSQRT_NOBRANCH 1       // the 4x4 Cholesky of both solver kernels uses rsqrt_nobranch (0: the library rsqrt, A/B)
#endif
// Branch-free 1/sqrt(x) in double: hardware approximation (2^-22) + two Newton steps -> ~1 ulp.  The library rsqrt() wraps
// the same approximation in a test-and-call for denormal / huge arguments, and that branch ends the basic block: the
// scheduler can then not move independent work into the latency of a chain of rsqrts (the dense kernel's 4x4 Cholesky).
// x <= 0, NaN -> NaN / inf as the library; denormal x is flushed to 0 -> inf (the callers treat non-finite as breakdown).
__device__ __forceinline__ double rsqrt_nobranch(double x)
{
#ifdef QMPC_EMU
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
#endif
}

template <typename real> __device__ __forceinline__ bool rfinite(real x) { return x - x == real(0); }

}  // namespace qmpc
