"""The CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Runs without a GPU."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden_cfg, golden_system_tensors, load_golden, params_from_golden
from oracle import refmd

FORCE_CASES = [
    "water291_rf_switch",
    "water291_plain",
    "argon100_nocut",
    "argon100_cut",
    "water999_eq",
    "chain_amber_vacuum",
    "chain_amber_periodic",
    "chain_charmm_periodic",
    "adversarial_cutoff",
]


@pytest.mark.parametrize("name", FORCE_CASES)
@pytest.mark.parametrize("tag,dtype", [("f32", torch.float32), ("f64", torch.float64)])
def test_oracle_forces_energies_pairs(name, tag, dtype):
    g = load_golden(name)
    par = params_from_golden(g, precision=dtype)
    terms = [str(t) for t in g["terms"]]
    of = refmd.OracleForces(par, terms, **golden_cfg(g))
    pos, box = golden_system_tensors(g, dtype)
    F = torch.zeros_like(pos)
    E = of.compute(pos, box, F)
    ref_F = g["forces_" + tag]
    scale = max(1.0, np.abs(ref_F).max())
    tol = 2e-6 if dtype == torch.float32 else 1e-12
    assert np.abs(F.numpy() - ref_F).max() <= tol * scale
    keys = [str(k) for k in g["energy_keys"]]
    for r in range(len(E)):
        for c, k in enumerate(keys):
            ref = g["energies_" + tag][r, c]
            assert abs(E[r][k] - ref) <= (2e-6 if dtype == torch.float32 else 1e-12) * max(1.0, abs(ref)), k
    if "npairs_" + tag in g:
        pairs = of.neighbour_pairs(pos[0], torch.diagonal(box[0])).numpy().astype(np.int32)
        assert len(pairs) == int(g["npairs_" + tag])
        assert hashlib.sha256(pairs.tobytes()).hexdigest() == str(g["pairs_sha256_" + tag])
        if tag == "f32" and "pairs_f32" in g:
            assert np.array_equal(pairs, g["pairs_f32"])


def test_water291_reference_energies_match_survey():
    """SURVEY.md section 8c quotes these fp64 values for the tests/water fixture."""
    g = load_golden("water291_rf_switch")
    e = dict(zip([str(k) for k in g["energy_keys"]], g["energies_f64"][0]))
    assert abs(e["bonds"] - 78.26866) < 1e-4
    assert abs(e["angles"] - 31.33205) < 1e-4
    assert abs(e["electrostatics"] + 755.0536) < 1e-3
    assert abs(e["lj"] - 74.69531) < 1e-4


@pytest.mark.parametrize("tag,dtype", [("f32", torch.float32), ("f64", torch.float64)])
def test_oracle_integrator_trajectories(tag, dtype):
    g = load_golden("water291_rf_switch")
    t = load_golden("water291_traj")
    par = params_from_golden(g, precision=dtype)
    terms = [str(x) for x in g["terms"]]
    cfg = golden_cfg(g)
    for mode in ("nve", "lan"):
        of = refmd.OracleForces(par, terms, **cfg)
        pos, box = golden_system_tensors(g, dtype)
        vel = torch.tensor(t["vel0_" + tag], dtype=dtype)
        F = torch.zeros_like(pos)
        fn = lambda p, b, f: [sum(e.values()) for e in of.compute(p, b, f)]  # noqa: E731
        fn(pos, box, F)
        oi = refmd.OracleIntegrator(
            pos, vel, box, F, par.masses.to(dtype), fn, 1.0,
            gamma_ps=0.1 if mode == "lan" else None, T=300.0 if mode == "lan" else None,
        )
        tol = 1e-5 if dtype == torch.float32 else 1e-11
        if mode == "nve":
            ek, ep, T = oi.step(1)
            assert np.abs(pos.numpy() - t["nve_pos1_" + tag]).max() < tol
            ek, ep, T = oi.step(9)
            assert np.abs(pos.numpy() - t["nve_pos10_" + tag]).max() < tol
            assert np.abs(vel.numpy() - t["nve_vel10_" + tag]).max() < tol
            assert np.allclose(ek, t["nve_ekin10_" + tag], rtol=1e-5)
        else:
            ek, ep, T = oi.step(4, noise=torch.tensor(t["lan_noise_" + tag], dtype=dtype))
            assert np.abs(pos.numpy() - t["lan_pos4_" + tag]).max() < tol
            assert np.abs(vel.numpy() - t["lan_vel4_" + tag]).max() < tol
            assert np.allclose(T, t["lan_T4_" + tag], rtol=1e-5)


def test_cutoff_predicate_is_fma_chain():
    """The fp32 distance the reference compares with the cutoff is
    sqrt_rn(fma(z,z,fma(y,y,x*x))) -- the kernels hard-code this order."""
    torch.manual_seed(3)
    v = (torch.randn(50001, 3) * 5).float()
    n = torch.norm(v, dim=1).numpy()
    vd = v.double()
    fl = lambda t: t.float().double()  # noqa: E731
    s = fl(vd[:, 2] * vd[:, 2] + fl(vd[:, 1] * vd[:, 1] + fl(vd[:, 0] * vd[:, 0]))).float().numpy()
    assert np.array_equal(np.sqrt(s), n)
