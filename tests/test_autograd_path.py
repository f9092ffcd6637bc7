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


DEV = "cuda:0"  # (tests/test_mirrors_on_interpreter.py runs the gpu tests on the host interpreter with "cpu")

@pytest.fixture()
def cpu_forces(monkeypatch):
    from torchmd_b200 import Forces
    from torchmd_b200 import forces as fmod

    g = load_golden("water291_rf_switch")
    cfg = golden_cfg(g)
    terms = [str(t) for t in g["terms"]]
    par = params_from_golden(g, precision=torch.float32)
    ofs = {tg: refmd.OracleForces(params_from_golden(g, precision=torch.float64), terms, true_gradient=tg, **cfg) for tg in (False, True)}
    calls = {"n": 0, "nrep": [], "exact": []}

    def evaluate(self, pos, box, forces, sync=True, exact_gradient=False):
        calls["n"] += 1
        calls["nrep"].append(pos.shape[0])
        calls["exact"].append(bool(exact_gradient))
        f64 = torch.zeros(pos.shape, dtype=torch.float64)
        det = ofs[bool(exact_gradient)].compute(pos.double(), box.double(), f64)  # list of {term: value}
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
    g = dict(g)
    g["forces_autograd_f64"] = load_golden("water291_autograd")["forces_autograd_f64"]  # the reference's autograd forces
    return f, pos, box, g, calls


def test_autograd_forces_flag_runs_the_same_pass(cpu_forces):
    f, pos, box, g, calls = cpu_forces
    F = torch.zeros_like(pos)
    with pytest.raises(RuntimeError, match="require gradients"):
        f.compute(pos, box, F, explicit_forces=False)
    e = f.compute(pos.detach().requires_grad_(True), box, F, explicit_forces=False)  # reference test_replicas' call
    assert isinstance(e, list) and len(e) == pos.shape[0] and isinstance(e[0], float)
    assert abs(e[0] - g["energies_f64"][0].sum()) < 1e-6 * abs(g["energies_f64"][0]).sum()
    assert np.abs(F.numpy() - g["forces_autograd_f64"]).max() < 1e-4 and calls["exact"][-1] is True
    f.compute(pos, box, F)  # default: the explicit formulas
    assert np.abs(F.numpy() - g["forces_f64"]).max() < 1e-4 and calls["exact"][-1] is False
    with pytest.raises(RuntimeError, match="forces tensor required"):
        f.compute(pos, box, None)


def test_energy_tensor_backward_is_minus_force(cpu_forces):
    f, pos, box, g, calls = cpu_forces
    p = pos.detach().clone().requires_grad_(True)
    E = f.compute(p, box, None, toNumpy=False, calculateForces=False, explicit_forces=False)
    assert torch.is_tensor(E) and E.shape == (pos.shape[0],) and E.requires_grad
    w = torch.tensor([1.0, -2.5])[: pos.shape[0]]
    (E * w).sum().backward()
    assert np.abs(p.grad[0].numpy() + w[0].item() * g["forces_autograd_f64"][0]).max() < 1e-4
    assert np.abs(p.grad[1].numpy() + w[1].item() * g["forces_autograd_f64"][1]).max() < 3e-4
    # details: per-term tensors; equal weights work, unequal weights are refused
    p2 = pos.detach().clone().requires_grad_(True)
    D = f.compute(p2, box, None, returnDetails=True, toNumpy=False, calculateForces=False)
    assert set(D[0]) == set(f.energies) | {"external"} and D[0]["lj"].shape == (1,)
    tot = sum(torch.cat(list(d.values())).sum() for d in D)
    tot.backward()
    assert np.abs(p2.grad.numpy() + g["forces_autograd_f64"]).max() < 1e-4
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
    assert E.requires_grad and np.abs(F.numpy() - g["forces_autograd_f64"]).max() < 1e-4
    # explicit forces in the buffer, true gradient in the graph: two passes
    n0 = calls["n"]
    E2 = f.compute(p, box, F, toNumpy=False, explicit_forces=True)
    assert calls["n"] == n0 + 2 and np.abs(F.numpy() - g["forces_f64"]).max() < 1e-4
    E2.sum().backward()
    assert np.abs(p.grad.numpy() + g["forces_autograd_f64"]).max() < 1e-4


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
    assert np.abs(batch.grad[0].numpy() + g["forces_autograd_f64"]).max() < 1e-4
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
    want_auto = g["forces_autograd_f64"] + (-2.0 * 0.05 * (-0.3))
    F2 = torch.zeros_like(pos)
    f.compute(pos.detach().requires_grad_(True), box, F2, explicit_forces=False)  # external force by autograd
    assert np.abs(F2.numpy() - want_auto).max() < 1e-4
    p = pos.detach().clone().requires_grad_(True)
    E = f.compute(p, box, None, toNumpy=False, calculateForces=False)
    E.sum().backward()
    assert np.abs(p.grad.numpy() + want_auto).max() < 1e-4


@pytest.mark.gpu
def test_gpu_energy_backward_and_vmap():
    from torchmd_b200 import Forces

    dev = DEV
    g = load_golden("ala2_xsc_rf")
    par = params_from_golden(g, precision=torch.float32, device=dev)
    f = Forces(par, terms=[str(t) for t in g["terms"]], **golden_cfg(g))
    pos, box = golden_system_tensors(g, torch.float32, dev)
    p = pos.detach().clone().requires_grad_(True)
    E = f.compute(p, box, None, toNumpy=False, calculateForces=False, explicit_forces=False)
    E.sum().backward()
    # yardstick: the oracle in exact-gradient form (= the reference's autograd forces), fp64 values on the fp32 pair set
    of = refmd.OracleForces(params_from_golden(g, precision=torch.float64), [str(t) for t in g["terms"]], true_gradient=True,
                            decision_dtype=torch.float32, **golden_cfg(g))
    want = torch.zeros(pos.shape, dtype=torch.float64)
    of.compute(pos.cpu().double(), box.cpu().double(), want)
    want = want.numpy()
    scale = max(1.0, float(np.abs(want).max()) / 100.0)
    assert np.abs(p.grad.cpu().numpy().astype(np.float64) + want).max() < 1e-4 * scale
    e_ref = g["energies_f64"].sum(axis=1)
    assert np.abs(E.detach().cpu().numpy() - e_ref).max() < 1e-5 * np.abs(g["energies_f64"]).sum() + 2e-3
    batch = torch.stack([pos] * 4).requires_grad_(True)
    Ev = torch.vmap(f.compute, in_dims=(0,))(batch, box=box, forces=None, returnDetails=False, explicit_forces=False,
                                              calculateForces=False, toNumpy=False)
    Ev.sum().backward()
    assert Ev.shape == (4, pos.shape[0])
    assert np.abs(batch.grad[2].cpu().numpy().astype(np.float64) + want).max() < 1e-4 * scale


# ---- the force the autograd path returns: exact gradient of the switched LJ ------------------
def test_oracle_true_gradient_matches_reference_autograd_golden():
    """water291_autograd.npz: forces of the unmodified reference with explicit_forces=False."""
    g = load_golden("water291_rf_switch")
    ga = load_golden("water291_autograd")
    terms = [str(t) for t in g["terms"]]
    for tag, dtype, tol in (("f64", torch.float64, 1e-11), ("f32", torch.float32, 1e-4)):
        of = refmd.OracleForces(params_from_golden(g, precision=dtype), terms, true_gradient=True, **golden_cfg(g))
        pos, box = golden_system_tensors(g, dtype)
        F = torch.zeros_like(pos)
        of.compute(pos, box, F)
        assert np.abs(F.numpy() - ga["forces_autograd_" + tag]).max() < tol
        # and it is a different force field from the explicit one (the reference's 1/r quirk)
        assert np.abs(g["forces_" + tag] - ga["forces_autograd_" + tag]).max() > 1e-2


def test_kernel_switch_term_in_exact_gradient_form():
    """physics.cuh pair_terms with PairParams::true_gradient, compiled for the host."""
    import ctypes as C

    from conftest import ROOT

    hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    rng = np.random.default_rng(3)
    n, cutoff, switch = 20000, 9.0, 7.5
    r = rng.uniform(2.5, cutoff, n)
    s = (r * r).astype(np.float32)
    A = rng.uniform(1e4, 6e5, n).astype(np.float32)
    B = rng.uniform(10, 600, n).astype(np.float32)
    qq = np.zeros(n, np.float32)
    out = [np.zeros(n, np.float32) for _ in range(5)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    dist = torch.tensor(np.sqrt(s.astype(np.float64)))
    A64, B64 = torch.tensor(A.astype(np.float64)), torch.tensor(B.astype(np.float64))
    try:
        for exact in (1, 0):
            hc.hc_set_true_gradient(exact)
            hc.hc_pair_terms(n, p(s), p(qq), p(A), p(B), 1 << 6, 1, C.c_float(cutoff), 1, C.c_float(switch), 0,
                             C.c_float(0.0), C.c_float(0.0), *[p(o) for o in out])
            e_ref, f_ref = refmd.lj_pair(dist, A64, B64, 1, switch, cutoff, true_gradient=bool(exact))
            assert np.abs(out[4] - f_ref.numpy()).max() <= 3e-6 * max(1.0, np.abs(f_ref.numpy()).max())
            assert np.abs(out[1] - e_ref.numpy()).max() <= 3e-6 * max(1.0, np.abs(e_ref.numpy()).max())
    finally:
        hc.hc_set_true_gradient(0)
    # the exact form IS the derivative of the energy: central differences of the oracle's energy
    h = 1e-5
    ep, _ = refmd.lj_pair(dist + h, A64, B64, 1, switch, cutoff)
    em, _ = refmd.lj_pair(dist - h, A64, B64, 1, switch, cutoff)
    _, f_exact = refmd.lj_pair(dist, A64, B64, 1, switch, cutoff, true_gradient=True)
    inner = (dist > 2.6) & (dist < cutoff - 2 * h)
    num = ((ep - em) / (2 * h))[inner]
    assert ((num - f_exact[inner]).abs() <= 1e-6 * (1 + f_exact[inner].abs())).all()


@pytest.mark.gpu
def test_gpu_autograd_mode_returns_the_reference_autograd_forces():
    from torchmd_b200 import Forces

    dev = DEV
    g, ga = load_golden("water291_rf_switch"), load_golden("water291_autograd")
    par = params_from_golden(g, precision=torch.float32, device=dev)
    f = Forces(par, terms=[str(t) for t in g["terms"]], **golden_cfg(g))
    pos, box = golden_system_tensors(g, torch.float32, dev)
    F = torch.zeros_like(pos)
    e = f.compute(pos.detach().requires_grad_(True), box, F, explicit_forces=False)
    assert np.abs(F.cpu().numpy().astype(np.float64) - ga["forces_autograd_f64"]).max() < 1e-4
    assert abs(e[0] - g["energies_f64"][0].sum()) < 1e-5 * np.abs(g["energies_f64"][0]).sum() + 2e-3
    f.compute(pos, box, F)  # back to the explicit convention
    assert np.abs(F.cpu().numpy().astype(np.float64) - g["forces_f64"]).max() < 1e-4
