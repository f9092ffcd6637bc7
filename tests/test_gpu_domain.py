"""Decomposed step on the GPU with a single rank (world size 1 NCCL group): the owned-
range code paths, the gather buffer and the CUDA-graph replay must reproduce the plain
Integrator bit for bit.  Multi-rank runs are exercised by bench.py --gpus N."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(seed=3):
    from torchmd_b200 import Forces, System, maxwell_boltzmann, testsystems

    sysd = testsystems.water_box(1000, seed=seed)
    par = testsystems.water_parameters(sysd, device=DEV)
    n = len(sysd["coords"])
    system = System(n, 1, torch.float32, DEV)
    system.set_positions(sysd["coords"])
    system.set_box(sysd["box"])
    torch.manual_seed(5)
    system.set_velocities(maxwell_boltzmann(par.masses, 300.0, 1))
    forces = Forces(par, terms=["lj", "electrostatics", "bonds", "angles"], cutoff=9.0, rfa=True, switch_dist=7.5)
    forces.compute(system.pos, system.box, system.forces)
    return system, forces


def test_owned_subset_forces_match_full():
    """Forces of an owned sub-range equal the same atoms' forces in the full evaluation."""
    from torchmd_b200 import _lib

    system, forces = _setup()
    full = system.forces.clone()
    n = system.pos.shape[1]
    lo, cnt = n // 3, n // 2
    _lib.check(_lib.lib().tmd_set_owned_atoms(forces._ctx, lo, cnt))
    part = torch.zeros_like(full)
    e = forces.compute(system.pos, system.box, part, returnDetails=True)[0]
    assert torch.equal(part[0, lo : lo + cnt], full[0, lo : lo + cnt])
    _lib.check(_lib.lib().tmd_set_owned_atoms(forces._ctx, 0, n))
    efull = forces.compute(system.pos, system.box, part, returnDetails=True)[0]
    assert torch.equal(part, full)
    assert 0 < abs(e["lj"]) < abs(efull["lj"]) * 1.0001 + 1.0


def test_decomposed_world1_matches_integrator_bitwise():
    from torchmd_b200 import Integrator
    from torchmd_b200.domain import DecomposedIntegrator

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29571")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    sa, fa = _setup()
    torch.manual_seed(9)
    ia = Integrator(sa, fa, 1.0, DEV, gamma=0.1, T=300.0)
    sb, fb = _setup()
    torch.manual_seed(9)
    ib = DecomposedIntegrator(sb, fb, 1.0, DEV, gamma=0.1, T=300.0, use_graph=True)
    assert ib.integ.seed == ia.seed
    ea = ia.step(niter=40)
    eb = ib.step(niter=40)
    assert torch.equal(sa.pos, sb.pos) and torch.equal(sa.vel, sb.vel)
    np.testing.assert_allclose(ea[0], eb[0], rtol=1e-6)
    np.testing.assert_allclose(ea[1], eb[1], rtol=1e-9, atol=1e-6)
    assert fb.stats()["rebuilds"] >= 2
