"""The differentiable / vmappable energy path of ``Forces.compute`` (reference forces.py:94-100,
328-346; reference tests test_replicas / test_vmap, tests/test_torchmd.py:469-605).

The plumbing -- the autograd.Function whose backward is -F, its vmap rule (batch folded into
replicas), the return formats, the external-potential hook -- is exercised here on the CPU by
substituting the kernel pass (``Forces._evaluate``, a CUDA call) with the oracle; the gpu-marked
test at the end runs the real kernels.
"""
import os

import numpy as np
import pytest
import torch

from conftest import golden_cfg, golden_system_tensors, load_golden, params_from_golden
from oracle import refmd


@pytest.fixture()
def cpu_forces(monkeypatch):
    from torchmd_b200 import Forces
    from torchmd_b200 import forces as fmod

    g = load_golden("water291_rf_switch")
    cfg = golden_cfg(g)
    terms = [str(t) for t in g["terms"]]
    par = params_from_golden(g, precision=torch.float32)
    of = refmd.OracleForces(params_from_golden(g, precision=torch.float64), terms, **cfg)
    calls = {"n": 0, "nrep": []}

    def evaluate(self, pos, box, forces, sync=True):
        calls["n"] += 1
        calls["nrep"].append(pos.shape[0])
        f64 = torch.zeros(pos.shape, dtype=torch.float64)
        det = of.compute(pos.double(), box.double(), f64)  # list of {term: value}
        forces.copy_(f64.to(forces.dtype))
        ene = torch.zeros(pos.shape[0], fmod.NUM_ENERGIES, dtype=torch.float64)
        for r, d in enumerate(det):
            for k, v in d.items():
                if k != "external":
                    ene[r, fmod.ENERGY_SLOTS.index(k)] = float(v)
        return ene

    monkeypatch.setattr(Forces, "_evaluate", evaluate)
    f = Forces(par, terms=terms, **cfg)
    pos, box = golden_system_tensors(g, torch.float32)
    return f, pos, box, g, calls


def test_autograd_forces_flag_runs_the_same_pass(cpu_forces):
    f, pos, box, g, calls = cpu_forces
    F = torch.zeros_like(pos)
    with pytest.raises(RuntimeError, match="require gradients"):
        f.compute(pos, box, F, explicit_forces=False)
    e = f.compute(pos.detach().requires_grad_(True), box, F, explicit_forces=False)  # reference test_replicas' call
    assert isinstance(e, list) and len(e) == pos.shape[0] and isinstance(e[0], float)
    assert abs(e[0] - g["energies_f64"][0].sum()) < 1e-6 * abs(g["energies_f64"][0]).sum()
    assert np.abs(F.numpy() - g["forces_f64"]).max() < 1e-4
    with pytest.raises(RuntimeError, match="forces tensor required"):
        f.compute(pos, box, None)


def test_energy_tensor_backward_is_minus_force(cpu_forces):
    f, pos, box, g, calls = cpu_forces
    p = pos.detach().clone().requires_grad_(True)
    E = f.compute(p, box, None, toNumpy=False, calculateForces=False, explicit_forces=False)
    assert torch.is_tensor(E) and E.shape == (pos.shape[0],) and E.requires_grad
    w = torch.tensor([1.0, -2.5])[: pos.shape[0]]
    (E * w).sum().backward()
    assert np.abs(p.grad[0].numpy() + w[0].item() * g["forces_f64"][0]).max() < 1e-4
    assert np.abs(p.grad[1].numpy() + w[1].item() * g["forces_f64"][1]).max() < 3e-4
    # details: per-term tensors; equal weights work, unequal weights are refused
    p2 = pos.detach().clone().requires_grad_(True)
    D = f.compute(p2, box, None, returnDetails=True, toNumpy=False, calculateForces=False)
    assert set(D[0]) == set(f.energies) | {"external"} and D[0]["lj"].shape == (1,)
    tot = sum(torch.cat(list(d.values())).sum() for d in D)
    tot.backward()
    assert np.abs(p2.grad.numpy() + g["forces_f64"]).max() < 1e-4
    p3 = pos.detach().clone().requires_grad_(True)
    D3 = f.compute(p3, box, None, returnDetails=True, toNumpy=False, calculateForces=False)
    with pytest.raises(NotImplementedError, match="per term"):
        (D3[0]["lj"].sum() + 2.0 * D3[0]["bonds"].sum()).backward()
    # without grad tracking the tensors are plain values
    E0 = f.compute(pos, box, torch.zeros_like(pos), toNumpy=False)
    assert torch.is_tensor(E0) and not E0.requires_grad


def test_forces_filled_alongside_differentiable_energy(cpu_forces):
    f, pos, box, g, calls = cpu_forces
    p = pos.detach().clone().requires_grad_(True)
    F = torch.full_like(pos, 3.0)
    E = f.compute(p, box, F, toNumpy=False, explicit_forces=False)
    assert E.requires_grad and np.abs(F.numpy() - g["forces_f64"]).max() < 1e-4


def test_vmap_folds_the_batch_into_replicas(cpu_forces):
    f, pos, box, g, calls = cpu_forces
    nb = 3
    batch = torch.stack([pos, pos + 0.01, pos - 0.02]).requires_grad_(True)  # (B,R,N,3)
    calls["n"] = 0
    E = torch.vmap(f.compute, in_dims=(0,))(batch, box=box, forces=None, returnDetails=False, explicit_forces=False,
                                             calculateForces=False, toNumpy=False)  # reference test_vmap's call
    assert E.shape == (nb, pos.shape[0])
    assert calls["n"] == 1 and calls["nrep"][-1] == nb * pos.shape[0]  # one kernel pass for the whole batch
    E.sum().backward()
    assert batch.grad.shape == batch.shape
    assert np.abs(batch.grad[0].numpy() + g["forces_f64"]).max() < 1e-4
    assert abs(float(E.detach()[0, 0]) - g["energies_f64"][0].sum()) < 1e-3


class Restraint:
    """Differentiable external potential: E = k |x - x0|^2 summed over atoms."""

    def __init__(self, x0, k=0.05):
        self.x0, self.k = x0, k

    def calculate(self, pos, box):
        d = pos - self.x0
        return self.k * (d * d).sum(dim=(1, 2)), -2.0 * self.k * d


def test_external_potential_in_both_force_modes(cpu_forces):
    f, pos, box, g, calls = cpu_forces
    f.external = Restraint(pos.detach() + 0.3)
    want = g["forces_f64"] + (-2.0 * 0.05 * (-0.3))
    F = torch.zeros_like(pos)
    e = f.compute(pos, box, F, returnDetails=True)
    assert abs(e[0]["external"] - 0.05 * 0.09 * pos[0].numel()) < 1e-3
    assert np.abs(F.numpy() - want).max() < 1e-4
    F2 = torch.zeros_like(pos)
    f.compute(pos.detach().requires_grad_(True), box, F2, explicit_forces=False)  # external force by autograd
    assert np.abs(F2.numpy() - want).max() < 1e-4
    p = pos.detach().clone().requires_grad_(True)
    E = f.compute(p, box, None, toNumpy=False, calculateForces=False)
    E.sum().backward()
    assert np.abs(p.grad.numpy() + want).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("TMD_B200_VALIDATE") != "1",
                    reason="autograd path on the kernels: plumbing checked on CPU, not yet run on a B200 (set TMD_B200_VALIDATE=1)")
def test_gpu_energy_backward_and_vmap():
    from torchmd_b200 import Forces

    dev = "cuda:0"
    g = load_golden("ala2_xsc_rf")
    par = params_from_golden(g, precision=torch.float32, device=dev)
    f = Forces(par, terms=[str(t) for t in g["terms"]], **golden_cfg(g))
    pos, box = golden_system_tensors(g, torch.float32, dev)
    p = pos.detach().clone().requires_grad_(True)
    E = f.compute(p, box, None, toNumpy=False, calculateForces=False, explicit_forces=False)
    E.sum().backward()
    scale = max(1.0, float(np.abs(g["forces_f64"]).max()) / 100.0)
    assert np.abs(p.grad.cpu().numpy().astype(np.float64) + g["forces_f64"]).max() < 1e-4 * scale
    e_ref = g["energies_f64"].sum(axis=1)
    assert np.abs(E.detach().cpu().numpy() - e_ref).max() < 1e-5 * np.abs(g["energies_f64"]).sum() + 2e-3
    batch = torch.stack([pos] * 4).requires_grad_(True)
    Ev = torch.vmap(f.compute, in_dims=(0,))(batch, box=box, forces=None, returnDetails=False, explicit_forces=False,
                                              calculateForces=False, toNumpy=False)
    Ev.sum().backward()
    assert Ev.shape == (4, pos.shape[0])
    assert np.abs(batch.grad[2].cpu().numpy().astype(np.float64) + g["forces_f64"]).max() < 1e-4 * scale
