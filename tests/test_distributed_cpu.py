"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: band / atmosphere plans and the all-gather assembly.
The compute kernels are not involved (they need a GPU); the shards here are synthetic functions of the global index."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radtxfr_b200 import distributed as D


def test_plan_bands_properties():
    for nx, world in ((800000, 8), (800000, 3), (5000, 4), (100, 2), (2048 * 7 + 5, 4)):
        bands = D.plan_bands(nx, world)
        assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == nx
        for (a, b), (c, d) in zip(bands, bands[1:]):
            assert b == c and a <= b
        assert all(a % D.TILE == 0 for a, _ in bands if a < nx)
    # cost-balanced plan: all the cost in the last tiles -> early bands are wide
    w = np.zeros(10); w[-2:] = 1.0
    bands = D.plan_bands(10 * D.TILE, 2, weights=w)
    assert bands[0] == (0, 9 * D.TILE) and bands[1] == (9 * D.TILE, 10 * D.TILE)


def test_estimate_tile_costs_balances_dense_regions():
    """Cost-balanced bands: a line list three times denser in the upper half of the axis moves the cut towards it."""
    import types
    rng = np.random.default_rng(0)
    nu = np.sort(np.concatenate([rng.uniform(700.0, 1100.0, 2000), rng.uniform(1100.0, 1500.0, 6000)]))
    db = types.SimpleNamespace(nu_sorted=nu, gamma_air_mean=0.06, n_air_mean=0.6, mass_mean=18.0)
    X = np.linspace(700.0, 1500.0, 800000)
    T = np.linspace(290.0, 220.0, 10); p = np.geomspace(1.0, 0.01, 10)
    cost = D.estimate_tile_costs(db, T, p, X)
    assert cost.size == (X.size + D.TILE - 1) // D.TILE and (cost > 0).all()
    half = cost.size // 2
    assert 2.8 < cost[half:].sum() / cost[:half].sum() < 4.5       # three times the lines, and wider Doppler cores
    by_cost = D.plan_bands(X.size, 2, weights=cost)
    by_size = D.plan_bands(X.size, 2)
    assert by_cost[0][1] > by_size[0][1] and by_cost[0][1] % D.TILE == 0
    loads = [cost[a // D.TILE:(b + D.TILE - 1) // D.TILE].sum() for a, b in by_cost]
    assert abs(loads[0] - loads[1]) / sum(loads) < 0.02


def test_plan_atmospheres():
    o = D.plan_atmospheres(10, 4)
    assert [list(x) for x in o] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert sorted(np.concatenate(o).tolist()) == list(range(10))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nx, n_atm, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bands = D.plan_bands(nx, world)
        i0, i1 = bands[rank]
        idx = torch.arange(i0, i1, dtype=torch.float64)
        local = torch.stack([idx * 1.5, -idx, idx * idx])                  # 3 "fields" that encode the global index
        full = D.all_gather_bands(local, bands)
        g = torch.arange(nx, dtype=torch.float64)
        ok1 = torch.equal(full, torch.stack([g * 1.5, -g, g * g]))
        owners = D.plan_atmospheres(n_atm, world)
        mine = torch.as_tensor(owners[rank], dtype=torch.float64)
        loc = mine[:, None] * torch.ones(1, 4, dtype=torch.float64) + torch.arange(4, dtype=torch.float64)[None, :] * 0.25
        gath = D.all_gather_atmospheres(loc, owners, n_atm)
        exp = torch.arange(n_atm, dtype=torch.float64)[:, None] + torch.arange(4, dtype=torch.float64)[None, :] * 0.25
        ok2 = torch.equal(gath, exp)
        ret[rank] = bool(ok1 and ok2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,n_atm", [(2048 * 5 + 17, 7), (100, 1)])
def test_gather_world2_gloo(nx, n_atm):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, n_atm, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}
