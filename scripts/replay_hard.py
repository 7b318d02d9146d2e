"""Replay dumped OCPs (scripts/dump_hard.py -> gpurun_out/hard_cases.npy) through the emulated kernels on the CPU.
usage: python scripts/replay_hard.py [flavour] [case indices ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc
from emu import emu
from helpers import make_gp
flavour = sys.argv[1] if len(sys.argv) > 1 else ""
cases = np.load(os.path.join(ROOT, "gpurun_out", "hard_cases.npy"), allow_pickle=True)
sel = [int(i) for i in sys.argv[2:]] or range(len(cases))
N, M = 20, 20
gp = make_gp(M)
quad = orc.quad_hummingbird()
policy = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in os.environ.get("OPTS", "").split(",") if kv}
cfg, keep = emu.make_config(1, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, gp.X, gp.theta, **policy)
for i in sel:
    c = cases[i]
    yref, yref_e = orc.make_yref(c["chunk"])
    x, u = c["xit"][None].copy(), c["uit"][None].copy()
    r = emu.solve(cfg, c["x0"][None], yref[None], yref_e[None], c["alpha"][None], x, u, act=c["act"][None].copy(), variant=2, flavour=flavour)
    xo, uo = c["xit"].copy(), c["uit"].copy()
    ro = orc.rti_step(quad, 1.0 / N, N, c["x0"], yref, yref_e, xo, uo, gp=gp, alpha=c["alpha"])
    print(f"case {i} (step {c['step']} vehicle {c['b']}): GPU it {c['iters']} rd {c['rounds']} st {c['status']} | emu it {r['iters'][0]} rd {r['rounds'][0]} st {r['status'][0]} "
          f"| oracle it {ro['iters']} st {ro['status']} | max|u-oracle| {np.abs(u[0] - uo).max():.1e}", flush=True)
