"""BASELINE configs 4 and 5 on one GPU: horizon/basis sweep with fp64 vs fp32 accuracy next to throughput, and the
high-speed lemniscate stress case with p99 step latency.  Writes profiles/<tag>_sweep.json and a markdown table.
    python scripts/sweep.py [tag]
Accuracy per cell = per-step parity against the CPU oracle (test infrastructure) on 48 random single-step problems of
that shape; throughput = closed loop incl. plant, CUDA events, 10 warm-up + 40 timed steps."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mpc_quad_ros_b200 import _capi
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import lemniscate_trajectories, random_smooth_trajectories
from oracle import oracle as orc
from helpers import make_gp, oracle_solve_batch, random_ocp_batch, u_rel, x_rel

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
W, K = 10, 40


def closed_loop(B, N, M, prec, workload="random_smooth", v_peak=15.0):
    dt = 1.0 / N
    quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
    gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B) if M else None
    opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe, precision=prec)
    Kt = W + K + N + 2
    traj = lemniscate_trajectories(B, Kt, dt, v_peak=v_peak) if workload == "lemniscate" else random_smooth_trajectories(B, Kt, dt)
    loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()))
    for _ in range(W):
        loop.step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    sat, its, rds = [], [], []
    ev[0].record()
    for s in range(K):
        loop.step()
        ev[s + 1].record()
    torch.cuda.synchronize()
    lat = np.array([ev[s].elapsed_time(ev[s + 1]) for s in range(K)])
    for s in range(5):                                    # a few more steps for solver statistics (host reads, untimed)
        loop.step()
        u = loop.u0
        sat.append(float(((u <= 0) | (u >= 1)).double().mean().item()))
        its.append(float(opt.solver_status()[1].double().mean().item()))
        rds.append(float(opt.solver_rounds().double().mean().item()))
    st = opt.solver_status()[0]
    return dict(steps_per_s=B * K / (lat.sum() * 1e-3), ms_per_step=float(lat.mean()), p50_ms=float(np.percentile(lat, 50)),
                p99_ms=float(np.percentile(lat, 99)), u0_saturated_frac=float(np.mean(sat)), ipm_iters=float(np.mean(its)),
                refine_rounds=float(np.mean(rds)), not_ok=int((st != 0).sum().item()))


def accuracy(N, M, prec):
    B, dt = 48, 1.0 / N
    quadv = orc.quad_hummingbird()
    gp = make_gp(M) if M else None
    sc = random_ocp_batch(B, N, dt, quadv, gp, seed=1000 + N + M)
    quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
    gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B) if M else None
    opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe, precision=prec)
    opt.set_iterate(torch.as_tensor(sc["xit"]), torch.as_tensor(sc["uit"]))
    yref, yref_e = torch.as_tensor(sc["yref"]).cuda().contiguous(), torch.as_tensor(sc["yref_e"]).cuda().contiguous()
    _capi.check(_capi.lib().qmpc_set_yref(opt._h, _capi.ptr(yref), _capi.ptr(yref_e), _capi.stream_ptr()))
    if gp is not None:
        opt.set_rgp_params(torch.as_tensor(sc["mu"]))
    x, u, _, _ = opt.run_optimization(torch.as_tensor(sc["x0"]).cuda())
    xo, uo, _, _ = oracle_solve_batch(sc, quadv, dt, N, gp)
    return u_rel(u.cpu().numpy(), uo), x_rel(x.cpu().numpy(), xo)


out = {"config4": [], "config5": []}
for N in (10, 20, 50):
    for M in (20, 50, 100):
        for prec in (64, 32):
            B = 4096 if N <= 20 else 2048
            r = closed_loop(B, N, M, prec)
            eu, ex = accuracy(N, M, prec)
            r.update(N=N, M=M, precision=prec, vehicles=B, u_rel_err_vs_oracle=eu, x_rel_err_vs_oracle=ex)
            out["config4"].append(r)
            print(json.dumps(r), flush=True)
for v_peak in (15.0, 20.0):
    r = closed_loop(16384, 20, 20, 64, "lemniscate", v_peak)
    r.update(N=20, M=20, precision=64, vehicles=16384, v_peak=v_peak)
    out["config5"].append(r)
    print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_sweep.json"), "w"), indent=1)
with open(os.path.join(ROOT, "gpurun_out", f"{tag}_sweep.md"), "w") as f:
    f.write("| N | M | precision | vehicles | control steps/s | ms/step | p99 ms | IPM it | rounds | u rel err | x rel err |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in out["config4"]:
        f.write(f"| {r['N']} | {r['M']} | fp{r['precision']} | {r['vehicles']} | {r['steps_per_s']:.3e} | {r['ms_per_step']:.2f} | {r['p99_ms']:.2f} | "
                f"{r['ipm_iters']:.2f} | {r['refine_rounds']:.2f} | {r['u_rel_err_vs_oracle']:.1e} | {r['x_rel_err_vs_oracle']:.1e} |\n")
    f.write("\nlemniscate stress (config 5), 16384 vehicles, N=20, M=20, fp64\n\n| v_peak | control steps/s | ms/step | p50 ms | p99 ms | u0 saturated | IPM it | rounds | not ok |\n|---|---|---|---|---|---|---|---|---|\n")
    for r in out["config5"]:
        f.write(f"| {r['v_peak']} | {r['steps_per_s']:.3e} | {r['ms_per_step']:.2f} | {r['p50_ms']:.2f} | {r['p99_ms']:.2f} | {r['u0_saturated_frac']:.3f} | "
                f"{r['ipm_iters']:.2f} | {r['refine_rounds']:.2f} | {r['not_ok']} |\n")
