// tests/emu/emu_kernels.cpp — TEST-ONLY.  Compiles the product kernels for the host (see emu_cuda.h) and exposes
// them with host pointers for tests/test_kernel_logic_emu.py.
#define QMPC_EMU 1
#include "emu_cuda.h"
#include "../../mpc_quad_ros_b200/csrc/mpc_kernels.cuh"
#include "../../mpc_quad_ros_b200/csrc/host_params.h"

using namespace qmpc;

template <typename real>
static int run_solve(const HostOcp* o, const double* x0, const double* yref, const double* yref_e,
                     const double* alpha, double* xit, double* uit, double* u0, double* cost, int* status,
                     int* iters, int* rounds, unsigned char* act, real* Wout)
{
    const int B = o->batch, N = o->n_nodes;
    std::vector<real> W((size_t)B * N * WT), fac((size_t)B * N * FAC);
    LinArgs<real> la;
    fill_lin_args(*o, la);
    la.xit = xit; la.uit = uit; la.yref = yref; la.alpha = alpha; la.gpX = o->gp_X; la.W = W.data();
    const unsigned threads = 128, total = (unsigned)B * N * 16;
    emu::launch((total + threads - 1) / threads, threads, 0, [&]() { qmpc_linearize_kernel<real>(la); });
    IpmArgs<real> ia;
    fill_ipm_args(*o, ia);
    ia.x0 = x0; ia.yref = yref; ia.yref_e = yref_e; ia.xit = xit; ia.uit = uit; ia.W = W.data(); ia.fac = fac.data();
    ia.u0 = u0; ia.cost = cost; ia.status = status; ia.iters = iters; ia.rounds = rounds; ia.act = act;
    constexpr int WARPS = 4;
    emu::launch((B + WARPS - 1) / WARPS, WARPS * 32, (size_t)WARPS * ia.smem_per_warp * sizeof(real),
                [&]() { qmpc_ipm_kernel<real, WARPS>(ia); });
    if (Wout) std::memcpy(Wout, W.data(), W.size() * sizeof(real));
    return 0;
}

extern "C" int emu_solve_f64(const HostOcp* o, const double* x0, const double* yref, const double* yref_e,
                             const double* alpha, double* xit, double* uit, double* u0, double* cost,
                             int* status, int* iters, int* rounds, unsigned char* act, double* Wout)
{
    return run_solve<double>(o, x0, yref, yref_e, alpha, xit, uit, u0, cost, status, iters, rounds, act, Wout);
}
extern "C" int emu_solve_f32(const HostOcp* o, const double* x0, const double* yref, const double* yref_e,
                             const double* alpha, double* xit, double* uit, double* u0, double* cost,
                             int* status, int* iters, int* rounds, unsigned char* act, float* Wout)
{
    return run_solve<float>(o, x0, yref, yref_e, alpha, xit, uit, u0, cost, status, iters, rounds, act, Wout);
}
