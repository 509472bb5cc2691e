"""world_size-2 (gloo, CPU) test of the initial-condition sharding and the single packed all-reduce: the sharded
evaluation equals the single-rank one.  The per-rank compute is the CPU oracle (test infrastructure)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quandary_b200 import desc_ndesign, sharding
from qb_systems import basis_states, cfg_system, random_alpha


def test_local_ranges_follow_reference_rules():
    assert sharding.local_ic_range(256, 3, 8) == (96, 128)
    assert [sharding.local_ic_range(16, r, 4) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 16)]
    with pytest.raises(ValueError):
        sharding.local_ic_range(10, 0, 4)       # src/main.cpp:150-153
    with pytest.raises(ValueError):
        sharding.local_ic_range(2, 0, 4)
    g = np.arange(5.0)
    buf = sharding.pack(g, {"obj_cost_re": 1.5, "fidelity_re": 0.25})
    g2, sc = sharding.unpack(buf, 5)
    assert np.array_equal(g, g2) and sc["obj_cost_re"] == 1.5 and sc["fidelity_re"] == 0.25 and sc["penalty"] == 0.0


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import qb_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc, ninit = cfg_system("C2", ntime=20)
    alpha = random_alpha(desc_ndesign(desc), amp=0.02)
    rho0 = basis_states(4, True)
    rbar = np.random.default_rng(5).standard_normal(rho0.shape) / ninit
    lo, hi = sharding.local_ic_range(ninit, rank, world)
    orc = qb_oracle.load()
    rT, g, _, _ = orc.batch_gradient(desc, alpha, rho0[lo:hi], rbar[lo:hi], nthreads=1)
    buf = torch.from_numpy(sharding.pack(g, {"fidelity_re": float(np.sum(rT * rbar[lo:hi]))}))
    sharding.allreduce_(buf)
    if rank == 0:
        q.put(buf.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_rank(oracle):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    buf = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    desc, ninit = cfg_system("C2", ntime=20)
    nd = desc_ndesign(desc)
    alpha = random_alpha(nd, amp=0.02)
    rho0 = basis_states(4, True)
    rbar = np.random.default_rng(5).standard_normal(rho0.shape) / ninit
    rT, g, _, _ = oracle.batch_gradient(desc, alpha, rho0, rbar, nthreads=1)
    g2, sc = sharding.unpack(buf, nd)
    assert np.max(np.abs(g2 - g)) <= 1e-13 * np.max(np.abs(g))
    assert abs(sc["fidelity_re"] - float(np.sum(rT * rbar))) < 1e-13
