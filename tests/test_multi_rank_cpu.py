"""N > 1 host logic on CPU: world_size-2 gloo processes.  Covers the vehicle sharding, the per-rank synthetic streams
and the shared-swarm exchange (information-form sums all-reduced, identical posterior on every rank) — with the oracle
standing in for the device kernels, since there is no GPU here."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _info_contribution(gp, mu, Cm, xt, yt):
    """sum_v j_v^T j_v / r_v and sum_v j_v^T y_v / r_v at the pre-update model (oracle-side restatement)"""
    M = gp.M
    out = np.zeros((3, M * M + M))
    for d in range(3):
        L, sf, sn = gp.theta[d]
        for v in range(xt.shape[0]):
            kv = sf ** 2 * np.exp(-0.5 * (xt[v, d] - gp.X[d]) ** 2 / L ** 2)
            jt = kv @ gp.Kx_inv[d]
            r = sf ** 2 - jt @ kv + sn ** 2
            out[d, :M * M] += np.outer(jt, jt).ravel() / r
            out[d, M * M:] += jt * yt[v, d] / r
    return out


def _apply_info(mu, Cm, info):
    M = mu.shape[1]
    mu2, C2 = np.empty_like(mu), np.empty_like(Cm)
    for d in range(3):
        Lam, eta = info[d, :M * M].reshape(M, M), info[d, M * M:]
        A = np.eye(M) + Cm[d] @ Lam
        C2[d] = np.linalg.solve(A, Cm[d])
        mu2[d] = np.linalg.solve(A, mu[d] + Cm[d] @ eta)
    return mu2, C2


def _worker(rank, world, port, total, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpc_quad_ros_b200.swarm import allreduce_info, shard_range
    from mpc_quad_ros_b200.trajectory import random_smooth_trajectories
    from oracle import oracle as orc
    from helpers import make_gp
    first, count = shard_range(total, rank, world)
    # every rank generates only its own vehicles; streams are keyed by the GLOBAL vehicle index
    traj = random_smooth_trajectories(count, 8, 0.05, seed=1234 + first)
    gp = make_gp(10)
    mu = np.zeros((3, gp.M)); Cm = np.stack([orc.rgp_prior(gp.X[d], gp.theta[d])[0] for d in range(3)])
    rng = np.random.default_rng(100)                       # same global sample table on every rank
    xt_all = rng.uniform(-8, 8, (total, 3)); yt_all = -0.3 * xt_all + 0.02 * rng.standard_normal((total, 3))
    info = torch.as_tensor(_info_contribution(gp, mu, Cm, xt_all[first:first + count], yt_all[first:first + count]))
    allreduce_info(info)
    mu2, C2 = _apply_info(mu, Cm, info.numpy())
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, count, traj[:, 0, :3].copy(), mu2, C2))
    if rank == 0:
        ret.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharding_and_shared_swarm_exchange_world2():
    total, world = 11, 2                                   # ragged: 6 + 5 vehicles
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, ret)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = ret.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mpc_quad_ros_b200.trajectory import random_smooth_trajectories
    from oracle import oracle as orc
    from helpers import make_gp
    # 1. the shards tile [0,total) exactly and reproduce the single-process streams
    ranges = sorted((g[0], g[1]) for g in gathered)
    assert ranges == [(0, 6), (6, 5)]
    full = random_smooth_trajectories(total, 8, 0.05, seed=1234)
    for first, count, p0, _, _ in gathered:
        assert np.array_equal(p0, full[first:first + count, 0, :3])
    # 2. every rank holds the same posterior, equal to sequential single-sample regress over ALL vehicles (order-free)
    assert np.array_equal(gathered[0][3], gathered[1][3]) and np.array_equal(gathered[0][4], gathered[1][4])
    gp = make_gp(10)
    rng = np.random.default_rng(100)
    xt_all = rng.uniform(-8, 8, (total, 3)); yt_all = -0.3 * xt_all + 0.02 * rng.standard_normal((total, 3))
    mu = np.zeros((3, gp.M)); Cm = np.stack([orc.rgp_prior(gp.X[d], gp.theta[d])[0] for d in range(3)])
    for v in range(total):
        for d in range(3):
            orc.rgp_regress(gp.X[d], gp.theta[d], gp.Kx_inv[d], mu[d], Cm[d], xt_all[v, d], yt_all[v, d])
    assert np.abs(gathered[0][3] - mu).max() < 1e-9 * max(1.0, np.abs(mu).max())
    assert np.abs(gathered[0][4] - Cm).max() < 1e-9 * np.abs(Cm).max()


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from mpc_quad_ros_b200.swarm import shard_range
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for first, count in spans:
                assert first == min(pos, total) and count >= 0
                pos += count
