// tests/emu/emu_cuda.h — TEST-ONLY host emulation of the small CUDA subset the kernels use.
// Lets tests compile mpc_quad_ros_b200/csrc/*.cuh with g++ (-DQMPC_EMU) and run the warp-level algorithms
// with one OS thread per lane (shuffles / __syncwarp = barriers), so that the kernel LOGIC can be checked
// against the oracle without a GPU.  Never linked into libqmpc.so; the product path has no CPU fallback.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

struct double2 { double x, y; };
struct float2 { float x, y; };

namespace emu {
struct uint3e { unsigned x, y, z; };
struct WarpState {
    std::barrier<> full{32}, lo{16}, hi{16};
    uint64_t slot[32];
};
inline thread_local uint3e t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local WarpState* t_warp;
inline thread_local std::barrier<>* t_block;
inline thread_local unsigned char* t_smem;
inline thread_local int t_lane;

inline void sync(unsigned mask)
{
    if (mask == 0xffffffffu) t_warp->full.arrive_and_wait();
    else if (mask == 0x0000ffffu) t_warp->lo.arrive_and_wait();
    else if (mask == 0xffff0000u) t_warp->hi.arrive_and_wait();
    else { std::fprintf(stderr, "emu: unsupported mask %08x\n", mask); std::abort(); }
}
template <typename T>
inline T shfl(unsigned mask, T v, int src)
{
    static_assert(sizeof(T) <= 8, "");
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    t_warp->slot[t_lane] = bits;
    sync(mask);
    uint64_t r = t_warp->slot[src & 31];
    sync(mask);
    T out;
    std::memcpy(&out, &r, sizeof(T));
    return out;
}
// run `body` as a grid of blocks x threads (blockDim multiple of 32); blocks run one after the other
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body)
{
    for (unsigned b = 0; b < grid; ++b) {
        std::vector<unsigned char> smem(smem_bytes + 64);
        unsigned char* sm = smem.data() + ((64 - (reinterpret_cast<uintptr_t>(smem.data()) & 63)) & 63);
        std::vector<std::unique_ptr<WarpState>> warps;
        for (unsigned w = 0; w < block / 32; ++w) warps.emplace_back(new WarpState());
        std::barrier<> blockbar((std::ptrdiff_t)block);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t, b]() {
                t_threadIdx = {t, 0, 0}; t_blockIdx = {b, 0, 0}; t_blockDim = {block, 1, 1}; t_gridDim = {grid, 1, 1};
                t_warp = warps[t / 32].get(); t_lane = t & 31; t_smem = sm; t_block = &blockbar;
                body();
                // a lane that returned early must still let its warp-mates pass later barriers: kernels in this
                // repo only exit whole warps / whole 16-lane groups, so nothing to do here.
            });
        for (auto& x : th) x.join();
    }
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
#define QMPC_DYN_SMEM(name) unsigned char* name = emu::t_smem
// static shared array: one per block; the emulation runs blocks one after the other, so one process-wide array per call site
#define QMPC_STATIC_SMEM(type, name, count) static type name[count]

template <typename T> inline T __shfl_sync(unsigned m, T v, int src) { return emu::shfl(m, v, src); }
template <typename T> inline T __shfl_xor_sync(unsigned m, T v, int x) { return emu::shfl(m, v, emu::t_lane ^ x); }
inline void __syncwarp(unsigned m = 0xffffffffu) { emu::sync(m); }
inline void __syncthreads() { emu::t_block->arrive_and_wait(); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T __ldg(const T* p) { return *p; }
