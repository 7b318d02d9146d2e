// dependent-issue latencies of the operations the solver kernels chain (one warp, one CTA): cycles per op
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double* out, long long* cyc, double seed, int n)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = seed + threadIdx.x; sm[threadIdx.x + 32] = seed;
    __syncwarp();
    double x = seed + threadIdx.x * 1e-3, y = 1.0000001, z = 1e-9;
    float xf = float(x), yf = 1.0000001f, zf = 1e-9f;
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = fma(x, y, z);
            if (OP == 1) x = x * y;
            if (OP == 2) x = x + z;
            if (OP == 3) x = rsqrt(x) + 1.0;
            if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1);
            if (OP == 5) { idx = (int)sm[idx & 63] & 31; }
            if (OP == 6) xf = fmaf(xf, yf, zf);
            if (OP == 7) x = 1.0 / x + 1.0;
            if (OP == 8) { x = fma(x, y, z); __syncwarp(); }
            if (OP == 9) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31]; __syncwarp(); }
            if (OP == 10) xf = rsqrtf(xf) + 1.0f;
            if (OP == 11) { x = fma(x, y, z); __syncthreads(); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
    out[threadIdx.x] = x + xf + idx;
}
// throughput with 4 independent chains per thread (ILP) : DFMA
__global__ void k_ilp(double* out, long long* cyc, double seed, int n)
{
    double x0 = seed, x1 = seed + 1, x2 = seed + 2, x3 = seed + 3, x4 = seed + 4, x5 = seed + 5, x6 = seed + 6, x7 = seed + 7, y = 1.0000001, z = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            x0 = fma(x0, y, z); x1 = fma(x1, y, z); x2 = fma(x2, y, z); x3 = fma(x3, y, z);
            x4 = fma(x4, y, z); x5 = fma(x5, y, z); x6 = fma(x6, y, z); x7 = fma(x7, y, z);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 16 * 8);
    const int n = 2000;
    const char* names[] = {"DFMA", "DMUL", "DADD", "rsqrt(double)+1", "SHFL(double)", "LDS dependent", "FFMA", "1/x (double)+1", "DFMA+syncwarp", "STS+sync+LDS+sync", "rsqrtf+1", "DFMA+syncthreads(256thr)"};
#define RUN(OP, TH) k<OP><<<1, TH>>>(out, cyc, 1.5, n); cudaDeviceSynchronize(); k<OP><<<1, TH>>>(out, cyc, 1.5, n); cudaDeviceSynchronize(); printf("%-28s %7.1f cycles/op\n", names[OP], double(cyc[OP]) / (n * 16.0));
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32) RUN(6, 32) RUN(7, 32) RUN(8, 32) RUN(9, 32) RUN(10, 32) RUN(11, 256)
    for (int th : {32, 128, 256, 512, 1024}) {
        k_ilp<<<1, th>>>(out, cyc, 1.5, n); cudaDeviceSynchronize();
        k_ilp<<<1, th>>>(out, cyc, 1.5, n); cudaDeviceSynchronize();
        printf("DFMA 8 chains/thread, %4d threads on one SM: %6.2f cycles per warp-DFMA per SMSP-warp => %.1f DFMA lanes/clk/SM\n", th,
               double(cyc[0]) / (n * 32.0), (double)th * n * 32.0 / double(cyc[0]));
    }
    cudaError_t e = cudaGetLastError();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
