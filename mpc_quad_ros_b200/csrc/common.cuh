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

template <typename real> __device__ __forceinline__ real rrsqrt(real x);
#ifdef QMPC_EMU
template <> __device__ __forceinline__ double rrsqrt<double>(double x) { return 1.0 / sqrt(x); }
template <> __device__ __forceinline__ float rrsqrt<float>(float x) { return 1.0f / sqrtf(x); }
#else
template <> __device__ __forceinline__ double rrsqrt<double>(double x) { return rsqrt(x); }
template <> __device__ __forceinline__ float rrsqrt<float>(float x) { return 1.0f / sqrtf(x); }
#endif

template <typename real> __device__ __forceinline__ bool rfinite(real x) { return x - x == real(0); }

}  // namespace qmpc
