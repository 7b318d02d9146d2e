// cycles per 80x80 factorisation of the dense kernel's three Cholesky variants, one CTA alone on an SM (or CTAS per SM)
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../mpc_quad_ros_b200/csrc/mpc_kernels_dense.cuh"
using namespace qmpc;
template <int V> __global__ void __launch_bounds__(256, 2) k(const double* Hin, double* out, long long* cyc, int N, int reps, int fixed)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const DenseLayout lay = dense_layout(N);
    IpmArgs<double> a{};
    DenseCtx<double> c{a};
    const int tid = threadIdx.x;
    c.tid = tid; c.lane = tid & 31; c.N = N; c.E = 4 * N; c.T = lay.T; c.GS = lay.GS;
    c.Ht = sm + lay.Ht; c.Lt = sm + lay.Lt;
    double* v = sm + lay.vec;
    c.rt = v + 11 * c.E; c.dR = v + 12 * c.E; c.cl = v + 7 * c.E; c.fx = c.cl;
    c.cbar = reinterpret_cast<unsigned long long*>(sm + lay.cbar); c.fgen = 0; c.smbase = sm;
    if (tid == 0) for (int kk = 0; kk < N; ++kk) flag_init(c.cbar + kk);
    { int i = 0; while ((i + 1) * (i + 2) / 2 <= tid) ++i; c.ti = i; c.tj = tid - i * (i + 1) / 2; }
    { int cc = 0, start = 0; while (cc < N && start + (N - cc + 1) <= tid) { start += N - cc + 1; ++cc; } c.fj = cc < N ? cc : -1; c.fi = cc + (tid - start); }
    for (int t = tid; t < lay.T * TS; t += 256) c.Ht[t] = Hin[t];
    long long tot = 0;
    for (int r = 0; r < reps; ++r) {
        if (tid < c.E) { c.rt[tid] = 1.0 + 0.01 * tid; c.dR[tid] = 0.5; c.fx[tid] = (tid % 7 == 3) ? 1.0 : 0.0; }
        __syncthreads();
        const long long t0 = clock64();
        if (V == 0) c.factor(fixed != 0);
        if (V == 1) c.factor_cols(fixed != 0);
        if (V == 2) c.factor_rl1(fixed != 0);
        tot += clock64() - t0;
        __syncthreads();
    }
    if (tid == 0 && blockIdx.x == 0) cyc[V] = tot / reps;
    if (blockIdx.x == 0) {
        for (int t = tid; t < lay.T * TS; t += 256) out[V * 8192 + t] = c.Lt[t];
        if (tid < c.E) out[V * 8192 + 7000 + tid] = c.rt[tid];
    }
}
int main(int argc, char** argv)
{
    const int N = 20, E = 80, T = N * (N + 1) / 2;
    const int ctas = argc > 1 ? atoi(argv[1]) : 1;
    std::vector<double> A(E * E), H(T * TS, 0.0);
    for (int i = 0; i < E; ++i) for (int j = 0; j < E; ++j) A[i * E + j] = (i == j ? 3.0 : 0.0) + 0.5 / (1.0 + abs(i - j)) + 0.01 * ((i * 7 + j * 7) % 5);
    for (int i = 0; i < E; ++i) for (int j = 0; j < i; ++j) A[i * E + j] = A[j * E + i];
    for (int bi = 0; bi < N; ++bi) for (int bj = 0; bj <= bi; ++bj) for (int q = 0; q < 4; ++q) for (int r = 0; r < 4; ++r)
        H[(bi * (bi + 1) / 2 + bj) * TS + q * 4 + r] = A[(4 * bi + q) * E + 4 * bj + r];
    double *dH, *out; long long* cyc;
    cudaMalloc(&dH, H.size() * 8); cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
    cudaMallocManaged(&out, 3 * 8192 * 8); cudaMallocManaged(&cyc, 8 * 8);
    const size_t smem = dense_layout(N).total * 8;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int fixed = 0; fixed < 2; ++fixed) {
        k<0><<<ctas, 256, smem>>>(dH, out, cyc, N, 50, fixed); cudaDeviceSynchronize();
        k<1><<<ctas, 256, smem>>>(dH, out, cyc, N, 50, fixed); cudaDeviceSynchronize();
        k<2><<<ctas, 256, smem>>>(dH, out, cyc, N, 50, fixed); cudaDeviceSynchronize();
        double e1 = 0, e2 = 0;
        for (int t = 0; t < 8192; ++t) { e1 = fmax(e1, fabs(out[8192 + t] - out[t])); e2 = fmax(e2, fabs(out[2 * 8192 + t] - out[t])); }
        printf("fixed %d ctas %d: cycles per factorisation: two-barrier %lld | column warps %lld | one-barrier %lld   (max diff vs two-barrier: %.1e %.1e)  %s\n",
               fixed, ctas, cyc[0], cyc[1], cyc[2], e1, e2, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
