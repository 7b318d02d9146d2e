"""In-process A/B of solver policies on the bench workload (value leg only: 8 stream groups, closed loop incl. plant).
For every policy string: `reps` fresh closed loops; each reports steps/s in the driver's window (steps 5..25 from the zero
iterate) and in a steady window (steps 40..100).  usage: python scripts/tune_policy.py [reps] "opts1" "opts2" ...
(opts like screen_rounds=5,bail_round=3; "" = library defaults)"""
import gc, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop, GroupedClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import random_smooth_trajectories

B, N, M, G = int(os.environ.get("BATCH", 4096)), 20, 20, int(os.environ.get("GROUPS", 8))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
configs = sys.argv[2:] or [""]
dev = torch.device("cuda:0")
traj = random_smooth_trajectories(B, 100 + N + 2, 1.0 / N, seed=1234)
x0 = traj[:, 0, :].copy()


def run(opts):
    pol = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in opts.split(",") if kv}

    def make(first=0, count=B):
        quad = Quadrotor3D(drag=True, batch=count, device=dev).set_hummingbird_params()
        gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=count, device=dev)
        opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe, **pol)
        return ClosedLoop(quad, opt, torch.as_tensor(traj[first:first + count]), torch.as_tensor(x0[first:first + count]))

    loop = GroupedClosedLoop(make, B, G) if G > 1 else make()
    gc.collect(); gc.disable()          # the previous run's handles are freed (cudaFree synchronises) before anything is timed
    out = []
    done = 0
    for (upto, timed) in ((5, False), (25, True), (40, False), (100, True)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if G > 1: loop.fork()
        for _ in range(upto - done): loop.step()
        if G > 1: loop.join()
        e1.record()
        torch.cuda.synchronize()
        if timed: out.append(B * (upto - done) / (e0.elapsed_time(e1) * 1e-3))
        done = upto
    gc.enable()
    return out


for opts in configs:
    r = np.array([run(opts) for _ in range(reps)])
    print(f"[{opts or 'defaults'}] driver-window M steps/s: {np.round(r[:, 0] / 1e6, 3).tolist()} median {np.median(r[:, 0]) / 1e6:.3f} | "
          f"steady: {np.round(r[:, 1] / 1e6, 3).tolist()} median {np.median(r[:, 1]) / 1e6:.3f}", flush=True)
