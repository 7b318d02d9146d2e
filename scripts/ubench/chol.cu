// latency of the pieces of the dense kernel's block-column chain (one warp, one CTA): cycles per call
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mpc_quad_ros_b200/csrc/mpc_kernels_dense.cuh"
using namespace qmpc;
template <int OP> __global__ void k(double* out, long long* cyc, double seed, int n)
{
    __shared__ __align__(16) double sm[512];
    for (int i = threadIdx.x; i < 512; i += 32) sm[i] = 0.01 * (i % 7) + seed * 1e-3;
    __syncwarp();
    double M[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) M[t] = 0.1 + 0.01 * t + seed * 1e-6;
    M[0] = 4 + seed * 1e-6; M[5] = 5; M[10] = 6; M[15] = 7;
    double acc[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) acc[t] = 0.5 + t;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        if (OP == 0) { Chol4<double> L; L.factor(M); M[0] = 4 + L.i3 * 1e-9; }
        if (OP == 1) { Chol4<double> L; L.factor(M); Inv4<double> Ni; Ni.from(L); M[0] = 4 + Ni.n30 * 1e-9; }
        if (OP == 2) { Chol4<double> L; L.factor(M); double z[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { L.fsolve(acc + q * 4, z); acc[q * 4] = 0.5 + z[3] * 1e-9; acc[q * 4 + 1] = z[2]; }
            M[0] = 4 + acc[12] * 1e-9; }
        if (OP == 3) {      // tile update from shared memory: 16 LDS.128 + 64 DFMA, address depends on the previous result
            const int off = ((int)acc[15]) & 1;
            double li[16], lj[16];
#pragma unroll
            for (int t = 0; t < 16; t += 2) { ld2(sm + off * 2 + t, li[t], li[t + 1]); ld2(sm + 64 + off * 2 + t, lj[t], lj[t + 1]); }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    double s = acc[q * 4 + r];
#pragma unroll
                    for (int c = 0; c < 4; ++c) s = fma(-li[q * 4 + c], lj[r * 4 + c], s);
                    acc[q * 4 + r] = s;
                }
        }
        if (OP == 4) { M[0] = rsqrt(M[0]) + 3.5; }
        if (OP == 5) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(M[0])); M[0] = y + 3.5; }
        if (OP == 6) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(M[0]));      // + 2 Newton steps
            double h = 0.5 * M[0]; double e = fma(-h * y, y, 0.5); y = fma(y, e, y); e = fma(-h * y, y, 0.5); y = fma(y, e, y); M[0] = y + 3.5; }
        if (OP == 7) { float yf = rsqrtf((float)M[0]); double y = yf; double h = 0.5 * M[0];
            double e = fma(-h * y, y, 0.5); y = fma(y, e, y); e = fma(-h * y, y, 0.5); y = fma(y, e, y); M[0] = y + 3.5; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) s += M[t] + acc[t];
    out[threadIdx.x] = s;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 16 * 8);
    const int n = 2000;
    const char* names[] = {"Chol4.factor", "Chol4.factor + Inv4.from", "Chol4.factor + 4 fsolve", "tile update (16 LDS.128 + 64 DFMA)", "rsqrt(double)", "rsqrt.approx.ftz.f64", "rsqrt.approx.f64 + 2 Newton", "rsqrtf + 2 Newton (fp64)"};
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1.5, n); cudaDeviceSynchronize(); k<OP><<<1, 32>>>(out, cyc, 1.5, n); cudaDeviceSynchronize(); printf("%-40s %7.1f cycles\n", names[OP], double(cyc[OP]) / n);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    // accuracy of the Newton variants
    return 0;
}
