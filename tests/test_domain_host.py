"""Host logic of the decomposed (multi-GPU) step on CPU: ownership ranges, the padded
gather layout and the per-step exchange, run with world_size 2 over gloo.  The CUDA
kernels are replaced by the oracle here (forces for the owned atoms, velocity Verlet on
the owned atoms); the GPU run of the same logic is tests/test_gpu_domain.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_ranges_cover_every_atom_once():
    from torchmd_b200.domain import SlabDecomposition

    for n, w in ((99999, 8), (99999, 4), (10, 3), (7, 8), (291, 2)):
        decs = [SlabDecomposition(n, w, r) for r in range(w)]
        owned = np.zeros(n, int)
        for d in decs:
            owned[d.lo : d.hi] += 1
            assert d.padded == d.chunk * w >= n and 0 <= d.lo <= d.hi <= n
        assert (owned == 1).all()
        assert decs[0].ranges() == [(d.lo, d.hi) for d in decs]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import refmd
    from torchmd_b200 import testsystems
    from torchmd_b200.domain import SlabDecomposition

    torch.set_num_threads(1)
    sysd = testsystems.water_box(30, seed=2)
    par = testsystems.water_parameters(sysd, precision=torch.float64)
    terms = ["lj", "electrostatics", "bonds", "angles"]
    of = refmd.OracleForces(par, terms, cutoff=4.0, rfa=True, switch_dist=3.0)
    n = len(sysd["coords"])
    box = torch.zeros(1, 3, 3, dtype=torch.float64)
    for k in range(3):
        box[0, k, k] = float(sysd["box"][k])
    torch.manual_seed(0)
    vel0 = refmd.maxwell_boltzmann(par.masses, 300.0, 1)

    def run(decomposed):
        pos = torch.tensor(sysd["coords"], dtype=torch.float64)[None].clone()
        vel = vel0.clone()
        F = torch.zeros_like(pos)
        m = par.masses.view(1, -1, 1)
        dt = 1.0 / refmd.TIMEFACTOR
        if decomposed:
            dec = SlabDecomposition(n, world, rank)
            buf, pos = dec.rehome(pos)
            send = torch.empty(dec.chunk * 3, dtype=pos.dtype)
            sl = slice(dec.lo, dec.hi)
        else:
            sl = slice(0, n)
        of.compute(pos, box, F)
        for _ in range(5):
            acc = F[:, sl] / m[:, sl]
            pos[:, sl] += vel[:, sl] * dt + 0.5 * acc * dt * dt
            vel[:, sl] += 0.5 * dt * acc
            if decomposed:
                dec.gather(buf, send)  # the exchange step
            of.compute(pos, box, F)  # the oracle computes all rows; only the owned ones are used
            vel[:, sl] += 0.5 * dt * (F[:, sl] / m[:, sl])
        return pos.clone(), vel.clone(), sl

    p1, v1, _ = run(False)
    p2, v2, sl = run(True)
    ok = torch.equal(p1, p2) and torch.equal(v1[:, sl], v2[:, sl])
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_decomposed_step_matches_single_process_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    results = dict(out.get(timeout=10) for _ in range(world))
    assert results == {0: True, 1: True}
