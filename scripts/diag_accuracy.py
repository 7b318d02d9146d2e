"""Per-OCP accuracy of one cold solve against the CPU oracle (test infrastructure) for a given shape: python scripts/diag_accuracy.py N M [seed]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mpc_quad_ros_b200 import _capi
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from oracle import oracle as orc
from helpers import make_gp, oracle_solve_batch, random_ocp_batch
N, M = int(sys.argv[1]), int(sys.argv[2]); seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1000 + N + M
B, dt = 48, 1.0 / N
quadv = orc.quad_hummingbird(); gp = make_gp(M) if M else None
sc = random_ocp_batch(B, N, dt, quadv, gp, seed=seed)
quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B) if M else None
opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe)
opt.set_iterate(torch.as_tensor(sc["xit"]), torch.as_tensor(sc["uit"]))
yref, yref_e = torch.as_tensor(sc["yref"]).cuda().contiguous(), torch.as_tensor(sc["yref_e"]).cuda().contiguous()
_capi.check(_capi.lib().qmpc_set_yref(opt._h, _capi.ptr(yref), _capi.ptr(yref_e), _capi.stream_ptr()))
if gp is not None: opt.set_rgp_params(torch.as_tensor(sc["mu"]))
x, u, _, _ = opt.run_optimization(torch.as_tensor(sc["x0"]).cuda())
st, it = opt.solver_status(); rd = opt.solver_rounds()
xo, uo, _, ito = oracle_solve_batch(sc, quadv, dt, N, gp)
eu = np.abs(u.cpu().numpy() - uo).reshape(B, -1).max(1); ex = np.abs(x.cpu().numpy() - xo).reshape(B, -1).max(1)
print("variant", os.environ.get("QMPC_IPM_VARIANT"), "rollout", os.environ.get("QMPC_FINAL_ROLLOUT"), "max u err %.2e max x err %.2e" % (eu.max(), ex.max()))
worst = np.argsort(-eu)[:6]
print("worst:", [(int(b), "%.1e" % eu[b], "%.1e" % ex[b], int(st[b]), int(it[b]), int(rd[b]), int(ito[b])) for b in worst], "(b, u err, x err, status, it, rounds, oracle it)")
