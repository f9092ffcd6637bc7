"""Peer-to-peer position exchange (DecomposedIntegrator(exchange="p2p"), tmd_dd_*) with a
single rank: the double-buffered positions, the push kernel, the flag wait and the CUDA-graph
replay per parity must reproduce the plain Integrator bit for bit.  Multi-rank:
scripts/p2p_check.py under torchrun (scripts/gpu_validate_p2p.sh).

STATUS: first run on a B200 in round 2 (green); part of the standing GPU suite since.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

pytestmark = [
    pytest.mark.gpu,
]
DEV = "cuda:0"
# (tests/test_mirrors_on_interpreter.py runs this test on the host interpreter with a smaller system)
SIZE = dict(waters=1000, cutoff=9.0, switch=7.5, skin=None, steps=(1, 2, 37), backend="nccl")


def _setup(seed=3):
    from torchmd_b200 import Forces, System, maxwell_boltzmann, testsystems

    sysd = testsystems.water_box(SIZE["waters"], seed=seed)
    par = testsystems.water_parameters(sysd, device=DEV)
    n = len(sysd["coords"])
    system = System(n, 1, torch.float32, DEV)
    system.set_positions(sysd["coords"])
    system.set_box(sysd["box"])
    torch.manual_seed(5)
    system.set_velocities(maxwell_boltzmann(par.masses, 300.0, 1))
    forces = Forces(par, terms=["lj", "electrostatics", "bonds", "angles"], cutoff=SIZE["cutoff"], rfa=True, switch_dist=SIZE["switch"], skin=SIZE["skin"])
    forces.compute(system.pos, system.box, system.forces)
    return system, forces


@pytest.mark.parametrize("use_graph", [False, True])
def test_p2p_world1_matches_integrator_bitwise(use_graph):
    from torchmd_b200 import Integrator
    from torchmd_b200.domain import DecomposedIntegrator

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29573")
        if SIZE["backend"] == "nccl":
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
        else:
            dist.init_process_group(SIZE["backend"], rank=0, world_size=1)
    sa, fa = _setup()
    torch.manual_seed(9)
    ia = Integrator(sa, fa, 1.0, DEV, gamma=0.1, T=300.0)
    sb, fb = _setup()
    torch.manual_seed(9)
    ib = DecomposedIntegrator(sb, fb, 1.0, DEV, gamma=0.1, T=300.0, use_graph=use_graph, exchange="p2p")
    assert ib.integ.seed == ia.seed and ib.exchange == "p2p"
    for niter in SIZE["steps"]:  # odd and even counts: both buffer parities start and end a call
        ea = ia.step(niter=niter)
        eb = ib.step(niter=niter)
        assert torch.equal(sa.pos, sb.pos) and torch.equal(sa.vel, sb.vel)
        np.testing.assert_allclose(ea[0], eb[0], rtol=1e-6)
        np.testing.assert_allclose(ea[1], eb[1], rtol=1e-9, atol=1e-6)
    # positions edited by the caller between calls are picked up
    sa.pos.add_(0.001)
    sb.pos.add_(0.001)
    ia.step(niter=3)
    ib.step(niter=3)
    assert torch.equal(sa.pos, sb.pos)
    assert fb.stats()["rebuilds"] >= 2
