"""Closed loop of the reference's pure-Python simulation (reference src/execute_trajectory.py:172-279), batched.

`simulate_trajectory` keeps the reference's signature and per-step order, calling the controller API method by
method (set_reference_trajectory -> run_optimization -> discrete_dynamics -> quad.update -> compute_a_drag ->
regress_and_update_RGP_model).  `ClosedLoop` is the same loop through the fused C-ABI call qmpc_step, with the
reference chunk and the plant also on the GPU, so that a control step never touches the host."""
import ctypes as C

import numpy as np
import torch

from . import _capi
from .utils import utils


def plant_substeps(optimization_dt, simulation_dt):
    """`while control_time < dt: ...; control_time += simulation_dt` (execute_trajectory.py:232-243): the float
    accumulation decides the count (20 @ 0.1, 11 @ 0.05 — SURVEY App. C-10)"""
    t, n = 0.0, 0
    while t < optimization_dt:
        t += simulation_dt
        n += 1
    return n


def simulate_trajectory(quad, quad_opt, quad_nominal, x0, x_trajectory, simulation_length, Nopt, simulation_dt, logger=None):
    """reference execute_trajectory.py:172-279.  Returns a dict of per-step lists with the reference's log keys."""
    log = {k: [] for k in ("x_odom", "x_pred_odom", "x_ref", "t_odom", "w_odom", "t_cpu", "cost_solution",
                           "rgp_mu_g_t", "rgp_C_g_t", "v_body", "a_drag")}
    single = quad_opt.batch == 1 and not torch.is_tensor(x_trajectory)
    simulation_time = 0.0
    n_sub = plant_substeps(quad_opt.optimization_dt, simulation_dt)
    x_pred_prev = None
    for i in range(Nopt):
        x_ref = utils.get_reference_chunk(x_trajectory, i, quad_opt.n_nodes)
        yref, yref_N = quad_opt.set_reference_trajectory(x_ref)
        x = quad.get_state(quaternion=True, stacked=True)
        x = x.copy() if single else x.clone()
        x_opt, w_opt, t_cpu, cost = quad_opt.run_optimization(x)
        w = w_opt[0, :].ravel() if single else w_opt[:, 0, :].contiguous()
        x_pred = quad_nominal.discrete_dynamics(x, w, quad_opt.optimization_dt)
        for _ in range(n_sub):
            quad.update(w, simulation_dt)
        mu = Cm = v_body = a_drag = None
        if quad_opt.gpe:
            x_pm1 = x_pred_prev if x_pred_prev is not None else x
            v_body, a_drag = utils.compute_a_drag(x, x_pm1, quad_opt.optimization_dt)
            mu, Cm = quad_opt.regress_and_update_RGP_model(v_body, a_drag)
        x_pred_prev = x_pred
        for k, v in (("x_odom", x), ("x_pred_odom", x_pred), ("x_ref", x_ref[0] if single else x_ref[:, 0]),
                     ("t_odom", simulation_time), ("w_odom", w), ("t_cpu", t_cpu), ("cost_solution", cost),
                     ("rgp_mu_g_t", mu), ("rgp_C_g_t", Cm), ("v_body", v_body), ("a_drag", a_drag)):
            log[k].append(v)
        if logger is not None:
            logger.log({k: log[k][-1] for k in log})
        simulation_time += quad_opt.optimization_dt
    return log


class ClosedLoop:
    """B vehicles, everything resident on the GPU: per control step
         chunk = get_reference_chunk(traj, i)            (qmpc_reference_chunk)
         qmpc_step(x_now, chunk) -> u0, x_pred, RGP      (set_reference, linearize, ipm, post_solve, rgp_regress)
         plant period with u0                             (qmpc_plant_period)
    """

    def __init__(self, quad, quad_opt, traj, x_init, simulation_dt=5e-3, shared_swarm=None):
        """traj: sampled references [B,K,13] (chunked on the GPU), or a trajectory.DeviceReference (generated on the GPU)"""
        self.quad, self.opt = quad, quad_opt
        self.shared_swarm = shared_swarm          # swarm.SharedSwarmRGP: one RGP for all vehicles / ranks (config 3)
        self.B, self.N, self.dev = quad_opt.batch, quad_opt.n_nodes, quad_opt.device
        self.refgen = None
        if torch.is_tensor(traj):
            self.traj = traj.to(self.dev, torch.float64).contiguous()
            assert self.traj.shape[0] == self.B and self.traj.shape[2] == 13
            self.K = self.traj.shape[1]
        else:
            self.refgen, self.traj, self.K = traj, None, traj.K
            assert traj.B == self.B
        self.x = x_init.to(self.dev, torch.float64).contiguous().clone()
        self.x_pred_prev = torch.zeros_like(self.x)
        self.chunk = torch.empty((self.B, self.N, 13), dtype=torch.float64, device=self.dev)
        self.u0 = torch.zeros((self.B, 4), dtype=torch.float64, device=self.dev)
        self.sim_dt, self.n_sub = simulation_dt, plant_substeps(quad_opt.optimization_dt, simulation_dt)
        self._quadv, self._plantv = quad.quad_vector(), quad.plant_vector()
        self.i = 0
        self.launches_per_step = 5 + (2 if quad_opt.gpe is not None and quad_opt.gpe.batch == self.B else 0)

    def control(self, x_now, i):
        """controller half of the step for an externally supplied state (end-to-end path: host buffers in/out)"""
        lib, s = _capi.lib(), _capi.stream_ptr()
        if self.refgen is not None:
            self.refgen.chunk(i, self.N, self.chunk)
        else:
            _capi.check(lib.qmpc_reference_chunk(self.B, self.K, _capi.ptr(self.traj), int(i), self.N, 1, _capi.ptr(self.chunk), s))
        if self.shared_swarm is not None:
            # shared model: residual -> accumulate -> all-reduce -> apply on a side stream, overlapped with this step's solve
            # (identical posterior on every rank; the solve of step t uses the model pushed after step t-1)
            self.shared_swarm.begin(x_now, self.x_pred_prev, first_step=(i == 0))
            self.opt.step(x_now, self.chunk, self.x_pred_prev, first_step=(i == 0), u0_out=self.u0, rgp=False)
            self.shared_swarm.end()
        else:
            self.opt.step(x_now, self.chunk, self.x_pred_prev, first_step=(i == 0), u0_out=self.u0)
        return self.u0

    def step(self):
        lib, s = _capi.lib(), _capi.stream_ptr()
        if self.shared_swarm is None and self.refgen is None:   # one C call: chunk -> solve -> u0 -> prediction -> residual -> RGP -> plant
            g = self.opt.gpe._h if self.opt.gpe is not None else C.c_void_p(0)
            _capi.check(lib.qmpc_closed_loop_step(self.opt._h, g, _capi.ptr(self.traj), self.K, int(self.i), _capi.ptr(self.x),
                                                  _capi.ptr(self.x_pred_prev), _capi.ptr(self.chunk), _capi.ptr(self.u0),
                                                  self._plantv.ctypes.data_as(C.c_void_p), C.c_double(self.sim_dt), self.n_sub, s))
        else:
            self.control(self.x, self.i)
            _capi.check(lib.qmpc_plant_period(self._quadv.ctypes.data_as(C.c_void_p), self._plantv.ctypes.data_as(C.c_void_p),
                                              self.B, _capi.ptr(self.x), _capi.ptr(self.u0), C.c_double(self.sim_dt), self.n_sub, s))
        self.i += 1

    def run(self, steps, record=False):
        xs, us = [], []
        for _ in range(steps):
            if record:
                xs.append(self.x.clone())
            self.step()
            if record:
                us.append(self.u0.clone())
        if record:
            return torch.stack(xs), torch.stack(us)


class GroupedClosedLoop:
    """The same closed loop with the vehicles split into G independent groups, each on its own CUDA stream.
    Vehicles never interact, so the groups advance independently: the kernels of one group overlap the tail of another
    (a step's duration is set by its slowest vehicle - IPM iteration counts vary - while most SMs are already idle)."""

    def __init__(self, make_loop, batch, groups):
        assert batch % groups == 0, "vehicles must split evenly over the groups"
        self.groups, self.per = groups, batch // groups
        self.streams = [torch.cuda.Stream() for _ in range(groups)]
        self.loops = []
        for g in range(groups):
            with torch.cuda.stream(self.streams[g]):
                self.loops.append(make_loop(g * self.per, self.per))
        torch.cuda.synchronize()

    def step(self):
        for s, lp in zip(self.streams, self.loops):
            with torch.cuda.stream(s):
                lp.step()

    def fork(self):
        """group streams wait for the work already queued on the current stream"""
        ev = torch.cuda.Event()
        ev.record()
        for s in self.streams:
            s.wait_event(ev)

    def join(self):
        """the current stream waits for every group"""
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)
