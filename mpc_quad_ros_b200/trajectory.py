"""Reference trajectories [K,13] (q = (1,0,0,0), r = 0 as in TrajectoryGenerator.load_trajectory,
reference src/trajectory_generation/TrajectoryGenerator.py:223-244).

Two ways to feed a closed loop:
  * sampled on the host once (`*_trajectories`, numpy set-up code) and chunked on the GPU every step
    (utils.get_reference_chunk -> qmpc_reference_chunk), or
  * generated on the GPU every step from per-vehicle parameters (`*_params` -> `DeviceReference` ->
    qmpc_reference_generate): nothing but the vehicle state ever crosses the host/device boundary."""
import ctypes as C

import numpy as np

REFGEN_NPAR = 32          # include/qmpc.h QMPC_REFGEN_NPAR
KIND_SINUSOIDS, KIND_LEMNISCATE, KIND_CIRCLE = 0, 1, 2


def _to_state(p, v):
    x = np.zeros(p.shape[:-1] + (13,))
    x[..., 0:3], x[..., 3], x[..., 7:10] = p, 1.0, v
    return x


def _csv_round(a):
    """the reference writes the samples with fmt='%.6f' and reads them back (TrajectoryGenerator.py:74,231)"""
    return np.array([float("%.6f" % v) for v in np.ravel(a)]).reshape(np.shape(a))


def sample_circle_trajectory_accelerating(radius, v_max, t_max=10, dt=0.01, start_point=np.zeros(3), csv_rounding=True):
    """TrajectoryGenerator.sample_circle_trajectory_accelerating (TrajectoryGenerator.py:41-74) + load_trajectory"""
    ts = np.arange(0, t_max, dt)
    n = len(ts)
    p, v = np.empty((n, 3)), np.empty((n, 3))
    w_max, phi = v_max / radius, 0.0
    for i in range(n):
        k = ((i + 1) / float(n) * 2) - 1
        dw = (np.sin((k * 2 * np.pi + np.pi * 3 / 2) * 0.5) + 1) / 2
        w = dw * w_max
        phi = phi + w * dt
        p[i] = np.array([radius * np.cos(phi), radius * np.sin(phi), 0]) + np.array([-radius, 0.0, 0.0]) + start_point
        v[i] = np.array([-radius * w * np.sin(phi), radius * w * np.cos(phi), 0])
    if csv_rounding:
        ts, p, v = _csv_round(ts), _csv_round(p), _csv_round(v)
    return _to_state(p, v), ts


def random_smooth_trajectories(B, K, dt, seed=1234, v_max=10.0, z0=3.0, a_max=None):
    """BASELINE config 2: per axis a sum of 3 sinusoids, amplitudes U(1,5) m, frequencies
    U(0.05,0.3) Hz, random phases, z offset, analytic velocity, scaled so that |v| <= v_max (and, if given, |a| <= a_max).  Vehicle b uses the
    counter-based stream seed+b (identical on every rank / in the CPU baseline).  returns [B,K,13]"""
    t = np.arange(K) * dt
    out = np.empty((B, K, 13))
    for b in range(B):
        rng = np.random.Generator(np.random.Philox(key=seed + b))
        amp, f, ph = rng.uniform(1, 5, (3, 3)), rng.uniform(0.05, 0.3, (3, 3)), rng.uniform(0, 2 * np.pi, (3, 3))
        arg = 2 * np.pi * f[:, :, None] * t[None, None, :] + ph[:, :, None]
        p = (amp[:, :, None] * np.sin(arg)).sum(1)
        v = (amp[:, :, None] * 2 * np.pi * f[:, :, None] * np.cos(arg)).sum(1)
        s = min(1.0, v_max / max(np.linalg.norm(v, axis=0).max(), 1e-9))
        if a_max is not None:
            acc = -(amp[:, :, None] * (2 * np.pi * f[:, :, None]) ** 2 * np.sin(arg)).sum(1)
            s = min(s, a_max / max(np.linalg.norm(acc, axis=0).max(), 1e-9))
        p, v = p * s, v * s
        p = p - p[:, :1]                      # start at the origin of the pattern ...
        p[2] += z0                            # ... hovering at z0
        out[b] = _to_state(p.T, v.T)
    return out


def lemniscate_trajectories(B, K, dt, v_peak=15.0, a=10.0, z0=3.0, seed=1234, ramp=3.0):
    """BASELINE config 5: p = (a sin wt, a sin wt cos wt, z0), w chosen so that the peak speed is v_peak
    (thrust limits active); vehicles differ by a random phase and heading."""
    t = np.arange(K) * dt
    w = v_peak / (a * np.sqrt(2.0))           # |v|max = a w sqrt(2) at the crossing
    out = np.empty((B, K, 13))
    for b in range(B):
        rng = np.random.Generator(np.random.Philox(key=seed + b))
        ph, yaw = rng.uniform(0, 2 * np.pi), rng.uniform(0, 2 * np.pi)
        th = ph + w * np.where(t < ramp, t * t / (2.0 * ramp), t - 0.5 * ramp)      # angular rate ramps 0 -> w over `ramp` s (start from rest)
        dth = w * np.where(t < ramp, t / ramp, 1.0)
        px, py = a * np.sin(th), a * np.sin(th) * np.cos(th)
        vx, vy = a * np.cos(th) * dth, a * np.cos(2 * th) * dth
        c, s = np.cos(yaw), np.sin(yaw)
        p = np.stack([c * px - s * py, s * px + c * py, np.full_like(t, z0)], 1)
        v = np.stack([c * vx - s * vy, s * vx + c * vy, np.zeros_like(t)], 1)
        p[:, :2] -= p[0, :2]
        out[b] = _to_state(p, v)
    return out


# ---- per-vehicle generator parameters (layout: include/qmpc.h qmpc_reference_generate) ----------------------------------

def random_smooth_params(B, K, dt, seed=1234, v_max=10.0, z0=3.0, a_max=None):
    """parameters [B,32] of the `random_smooth_trajectories` of the same arguments (same Philox streams, same scaling)"""
    t = np.arange(K) * dt
    par = np.zeros((B, REFGEN_NPAR))
    for b in range(B):
        rng = np.random.Generator(np.random.Philox(key=seed + b))
        amp, f, ph = rng.uniform(1, 5, (3, 3)), rng.uniform(0.05, 0.3, (3, 3)), rng.uniform(0, 2 * np.pi, (3, 3))
        arg = 2 * np.pi * f[:, :, None] * t[None, None, :] + ph[:, :, None]
        v = (amp[:, :, None] * 2 * np.pi * f[:, :, None] * np.cos(arg)).sum(1)
        s = min(1.0, v_max / max(np.linalg.norm(v, axis=0).max(), 1e-9))
        if a_max is not None:
            acc = -(amp[:, :, None] * (2 * np.pi * f[:, :, None]) ** 2 * np.sin(arg)).sum(1)
            s = min(s, a_max / max(np.linalg.norm(acc, axis=0).max(), 1e-9))
        p0 = (amp * np.sin(ph)).sum(1) * s
        par[b, 0:9], par[b, 9:18], par[b, 18:27] = amp.ravel(), f.ravel(), ph.ravel()
        par[b, 27], par[b, 28:31], par[b, 31] = s, p0, z0
    return par


def lemniscate_params(B, v_peak=15.0, a=10.0, z0=3.0, seed=1234, ramp=3.0):
    """parameters [B,32] of `lemniscate_trajectories`"""
    w = v_peak / (a * np.sqrt(2.0))
    par = np.zeros((B, REFGEN_NPAR))
    for b in range(B):
        rng = np.random.Generator(np.random.Philox(key=seed + b))
        ph, yaw = rng.uniform(0, 2 * np.pi), rng.uniform(0, 2 * np.pi)
        px, py = a * np.sin(ph), a * np.sin(ph) * np.cos(ph)
        c, s_ = np.cos(yaw), np.sin(yaw)
        par[b, :8] = [ph, yaw, w, a, z0, c * px - s_ * py, s_ * px + c * py, ramp]
    return par


def circle_params(radius, v_max, t_max=10, dt=0.01, start_point=np.zeros(3), csv_rounding=True, B=1):
    """parameters [B,32] of `sample_circle_trajectory_accelerating` (the same circle for every vehicle)"""
    n = len(np.arange(0, t_max, dt))
    par = np.zeros((B, REFGEN_NPAR))
    par[:, 0], par[:, 1], par[:, 2], par[:, 3:6], par[:, 6] = radius, v_max, n, start_point, float(bool(csv_rounding))
    return par


class DeviceReference:
    """reference generator on the GPU: `chunk(idx, N, out)` fills out [B,N,13] with what
    utils.get_reference_chunk(trajectory, idx, N, skip) would return for the sampled trajectory of K rows"""

    def __init__(self, kind, params, K, dt, device="cuda:0"):
        import torch
        self.kind, self.K, self.dt = int(kind), int(K), float(dt)
        self.params = torch.as_tensor(np.ascontiguousarray(params, dtype=np.float64), device=device).contiguous()
        assert self.params.shape[1] == REFGEN_NPAR
        self.B = self.params.shape[0]

    def chunk(self, idx, N, out, skip=1):
        from . import _capi
        _capi.check(_capi.lib().qmpc_reference_generate(self.kind, self.B, _capi.ptr(self.params), self.K, int(idx), int(N), int(skip),
                                                        C.c_double(self.dt), _capi.ptr(out), _capi.stream_ptr()))
        return out
