// tests/emu/emu_kernels.cpp — TEST-ONLY.  Compiles the product kernels for the host (see emu_cuda.h) and exposes
// them with host pointers for tests/test_kernel_logic_emu.py.
#define QMPC_EMU 1
#include "emu_cuda.h"
#include "../../mpc_quad_ros_b200/csrc/mpc_kernels.cuh"
#include "../../mpc_quad_ros_b200/csrc/mpc_kernels_dense.cuh"
#include "../../mpc_quad_ros_b200/csrc/host_params.h"

using namespace qmpc;

template <typename real>
static int run_solve(const HostOcp* o, const double* x0, const double* yref, const double* yref_e,
                     const double* alpha, double* xit, double* uit, double* u0, double* cost, int* status,
                     int* iters, int* rounds, unsigned char* act, real* Wout, int* hard_out = nullptr)
{
    const int B = o->batch, N = o->n_nodes;
    std::vector<real> W(((size_t)B * N + 1) * WT), fac((size_t)B * N * FAC);
    LinArgs<double, real> la;
    fill_lin_args(*o, la);
    la.xit = xit; la.uit = uit; la.yref = yref; la.alpha = alpha; la.gpX = o->gp_X; la.W = W.data();
    {
        double grid[6];
        gp_grid_detect(o->n_basis ? o->gp_X : nullptr, o->n_basis, grid);      // as qmpc_create does
        if (getenv("EMU_GP_DIRECT")) for (double& g : grid) g = 0;              // test hook: force the general (one exp per value) path
        fill_gp_grid(grid, la.mp);
    }
    emu::launch(((unsigned)B * N + LIN_NB - 1) / LIN_NB, LIN_THREADS, 0, [&]() { qmpc_linearize_kernel<double, real>(la); });
    if (getenv("EMU_ROUND_TILES_FP32"))      // experiment: how much of the fp32 build's error is the rounding of the tile data alone
        for (auto& v : W) v = real(float(v));
    IpmArgs<real> ia;
    fill_ipm_args(*o, ia);
    ia.x0 = x0; ia.yref = yref; ia.yref_e = yref_e; ia.xit = xit; ia.uit = uit; ia.W = W.data(); ia.fac = fac.data();
    ia.u0 = u0; ia.cost = cost; ia.status = status; ia.iters = iters; ia.rounds = rounds; ia.act = act;
    constexpr int WARPS = 4;
    const int variant = solver_variant(*o, DN_MAX_N);      // the same dispatch as capi.cu solve_impl
    if (hard_out) *hard_out = 0;
    if (variant == 1) {
        emu::launch((B + WARPS - 1) / WARPS, WARPS * 32, (size_t)WARPS * ia.smem_per_warp * sizeof(real),
                    [&]() { qmpc_ipm_kernel<real, WARPS>(ia); });
    } else {      // screening kernel (warm-started rounds) + dense kernel for the rest
        std::vector<int> list(B), cnt(1, 0);
        fill_screen_args(*o, ia);
        const bool screen = ia.warm_rounds > 0;
        if (screen) {
            ia.hard_list = list.data(); ia.hard_count = cnt.data();
            emu::launch((B + WARPS - 1) / WARPS, WARPS * 32, (size_t)WARPS * ia.smem_per_warp * sizeof(real),
                        [&]() { qmpc_ipm_kernel<real, WARPS>(ia); });
            if (hard_out) *hard_out = cnt[0];
        } else if (hard_out) *hard_out = B;
        DenseArgs<real> dn;
        dn.b = ia; dn.hard_list = screen ? list.data() : nullptr; dn.hard_count = screen ? cnt.data() : nullptr;
        int next = 0;
        dn.next_item = &next;
        if constexpr (sizeof(real) == 8)
            emu::launch(2, DN_THREADS, (size_t)dense_layout(N).total * sizeof(real), [&]() { qmpc_dense_kernel<real>(dn); });
    }
    if (Wout) std::memcpy(Wout, W.data(), (size_t)B * N * WT * sizeof(real));
    return 0;
}

extern "C" int emu_solve_f64(const HostOcp* o, const double* x0, const double* yref, const double* yref_e,
                             const double* alpha, double* xit, double* uit, double* u0, double* cost,
                             int* status, int* iters, int* rounds, unsigned char* act, double* Wout, int* hard_out)
{
    return run_solve<double>(o, x0, yref, yref_e, alpha, xit, uit, u0, cost, status, iters, rounds, act, Wout, hard_out);
}
extern "C" int emu_solve_f32(const HostOcp* o, const double* x0, const double* yref, const double* yref_e,
                             const double* alpha, double* xit, double* uit, double* u0, double* cost,
                             int* status, int* iters, int* rounds, unsigned char* act, float* Wout, int* hard_out)
{
    return run_solve<float>(o, x0, yref, yref_e, alpha, xit, uit, u0, cost, status, iters, rounds, act, Wout, hard_out);
}
