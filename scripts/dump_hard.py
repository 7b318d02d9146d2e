"""Dump the inputs of the slowest OCPs of a closed loop (IPM iterations >= IT_MIN or rounds >= RD_MIN) so that they can be
replayed through the emulated kernels on the CPU (tests/emu): x0, reference chunk, alpha, SQP iterate, remembered active set."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import random_smooth_trajectories
B, N, M = 4096, 20, 20
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
it_min, rd_min = int(os.environ.get("IT_MIN", 14)), int(os.environ.get("RD_MIN", 20))
sample_step, sample_n = int(os.environ.get("SAMPLE_STEP", -1)), int(os.environ.get("SAMPLE_N", 64))   # also dump a random sample of one step
quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B)
opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe)
traj = random_smooth_trajectories(B, steps + N + 2, 1.0 / N, seed=1234)
loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()))
cases, hist_it, hist_rd = [], np.zeros(64, dtype=np.int64), np.zeros(64, dtype=np.int64)
for s in range(steps):
    xit, uit = opt.get_iterate(); alpha = gpe.alpha_tensor().clone(); x_now = loop.x.clone(); act = opt.get_active_set()
    loop.step()
    st, it = opt.solver_status(); rd = opt.solver_rounds()
    if s >= 1:
        hist_it += np.bincount(it.cpu().numpy().clip(0, 63), minlength=64); hist_rd += np.bincount(rd.cpu().numpy().clip(0, 63), minlength=64)
        sel = torch.nonzero((it >= it_min) | (rd >= rd_min)).flatten().tolist()
        if s == sample_step:
            sel = np.random.default_rng(0).choice(B, sample_n, replace=False).tolist()
        for b in (sel if s == sample_step else sel[:6]):
            if len(cases) < 60 or s == sample_step:
                cases.append(dict(step=s, b=b, x0=x_now[b].cpu().numpy(), chunk=loop.chunk[b].cpu().numpy(), alpha=alpha[b].cpu().numpy(),
                                  xit=xit[b].cpu().numpy(), uit=uit[b].cpu().numpy(), act=act[b].cpu().numpy(),
                                  status=int(st[b]), iters=int(it[b]), rounds=int(rd[b])))
np.save(os.path.join(ROOT, "gpurun_out", "hard_cases.npy"), np.array(cases, dtype=object), allow_pickle=True)
print("saved", len(cases), "cases")
print("ipm iteration histogram (steps >= 1):", hist_it[:45].tolist())
print("rounds histogram:", hist_rd[:50].tolist())
for c in cases[:60]:
    print(c["step"], c["b"], "status", c["status"], "iters", c["iters"], "rounds", c["rounds"])
