"""Parity of the fixed-point pair kernel (k_pair_fx, TMD_B200_FX=1) with the oracle and
with the default float kernel.

STATUS: first run on a B200 in round 2 (26 passed; profiles/r02_validation_call1.txt); part of
the standing GPU suite since.  Same tolerances as
test_gpu_forces.py: pairs bit-exact, forces < 1e-4 kcal/mol/A against the fp64 oracle on
the fp32 pair set.
"""
import os

import numpy as np
import pytest
import torch

from conftest import golden_cfg, golden_system_tensors, load_golden, params_from_golden
from oracle import refmd

pytestmark = [
    pytest.mark.gpu,
]
DEV = "cuda:0"
PERIODIC_CASES = ["water291_rf_switch", "water291_plain", "argon100_cut", "water999_eq", "chain_amber_periodic",
                  "chain_charmm_periodic", "adversarial_cutoff", "ala2_xsc_rf"]


def _lib_mod():
    from torchmd_b200 import _lib

    return _lib


# systems whose box only meets the guard-free image condition (cutoff + 2 skin < 0.45 L) with a thin skin
THIN_SKIN = {"water999_eq": 0.2}


def make_forces(g, fx, **kw):
    """fx: 0 float kernel, 1 fixed-point kernel, 2 fixed-point + packed fp32x2 arithmetic (k_pair_fx2, taken
    for every lj/electrostatics term set with <= 16 atom types; other systems fall back to the fixed-point kernel)."""
    from torchmd_b200 import Forces

    old = os.environ.get("TMD_B200_FX")
    old_cl = os.environ.get("TMD_B200_CLUSTER")
    os.environ["TMD_B200_FX"] = str(int(fx))
    os.environ["TMD_B200_CLUSTER"] = "0"  # these tests are about the full-row kernels
    try:
        par = params_from_golden(g, precision=torch.float32, device=DEV)
        f = Forces(par, terms=[str(t) for t in g["terms"]], **golden_cfg(g), **kw)
        pos, box = golden_system_tensors(g, torch.float32, DEV)
        F = torch.full_like(pos, 7.0)
        E = f.compute(pos, box, F, returnDetails=True)  # the context reads the switch when it is finalised here
        f.pair_kernel = _lib_mod().lib().tmd_pair_kernel(f._ctx)  # 0 float, 1 fixed point, 2 packed, 3 packed without a box
    finally:
        if old is None:
            os.environ.pop("TMD_B200_FX", None)
        else:
            os.environ["TMD_B200_FX"] = old
        if old_cl is None:
            os.environ.pop("TMD_B200_CLUSTER", None)
        else:
            os.environ["TMD_B200_CLUSTER"] = old_cl
    return f, pos, box, F, E


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", PERIODIC_CASES)
def test_fixed_point_kernel_matches_golden(name, mode):
    g = load_golden(name)
    kw = {"skin": THIN_SKIN[name]} if name in THIN_SKIN else {}
    f, pos, box, F, E = make_forces(g, mode, **kw)
    f0, _, _, F0, E0 = make_forces(g, 0, **kw)
    # boxes below (cutoff + 2 skin) / 0.45 keep the float kernel by design (guard-free image condition)
    expect = 0 if name in ("water291_plain", "water291_rf_switch", "ala2_xsc_rf") else mode
    assert f.pair_kernel == expect, f"pair kernel {f.pair_kernel} ran instead of {expect}"
    assert f0.pair_kernel == 0
    if mode == 2:  # also the force-only instantiation of the packed kernel
        os.environ["TMD_B200_FX"] = "2"
        try:
            from torchmd_b200 import _lib
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().tmd_forces(f._ctx, pos.data_ptr(), F.data_ptr(), None, stream))
            torch.cuda.synchronize()
        finally:
            os.environ.pop("TMD_B200_FX", None)
    ref = g["forces_f64"]
    scale = max(1.0, float(np.abs(ref).max()) / 100.0)
    err = np.abs(F.cpu().numpy().astype(np.float64) - ref).max()
    err0 = np.abs(F0.cpu().numpy().astype(np.float64) - ref).max()
    print(f"{name}: max|dF| vs fp64 reference: fixed-point {err:.3e}, float kernel {err0:.3e}")
    dev = np.abs(g["forces_f32"].astype(np.float64) - ref).max()
    assert err < max(1e-4 * scale, 1.2 * dev)
    assert err <= err0 * 1.3 + 1e-5 * scale, "the fixed-point path must not be less accurate than the float path"
    keys = [str(k) for k in g["energy_keys"]]
    for r in range(len(E)):
        for c, k in enumerate(keys):
            e_ref = g["energies_f64"][r, c]
            assert abs(E[r][k] - e_ref) <= 1e-5 * abs(e_ref) + 2e-3, (k, E[r][k], e_ref)
    if "pairs_f32" in g:
        assert np.array_equal(f.neighbour_pairs(pos, box).cpu().numpy(), g["pairs_f32"])
    # deterministic
    if mode == 1:
        F1 = F.clone()
        f.compute(pos, box, F)
        assert torch.equal(F, F1)


@pytest.mark.parametrize("name", ["ala2_nobox_rf", "thrombin_nobox_rf", "argon100_nocut", "chain_amber_vacuum", "benzamidine_amber_nocut", "ligand_amber_nocut"])
def test_packed_kernel_without_a_box(name):
    """TMD_B200_FX=2 on systems without a box: k_pair2_open (packed arithmetic on the float records; thrombin has
    45 atom types and reads the LJ table from global memory).  Decisions are exact by construction: pairs bit-exact."""
    g = load_golden(name)
    f, pos, box, F, E = make_forces(g, 2)
    f0, _, _, F0, E0 = make_forces(g, 0)
    assert f.pair_kernel == (3 if not np.any(g["box"]) else 0) and f0.pair_kernel == 0  # (argon100_nocut: box without cutoff)
    ref = g["forces_f64"]
    scale = max(1.0, float(np.abs(ref).max()) / 100.0)
    dev = np.abs(g["forces_f32"].astype(np.float64) - ref).max()
    for tag, Fx in (("with energies", F.clone()), ("forces only", None)):
        if Fx is None:
            from torchmd_b200 import _lib
            os.environ["TMD_B200_FX"] = "2"
            try:
                _lib.check(_lib.lib().tmd_forces(f._ctx, pos.data_ptr(), F.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
            finally:
                os.environ.pop("TMD_B200_FX", None)
            Fx = F
        err = np.abs(Fx.cpu().numpy().astype(np.float64) - ref).max()
        err0 = np.abs(F0.cpu().numpy().astype(np.float64) - ref).max()
        print(f"{name} ({tag}): max|dF| vs fp64 reference: packed {err:.3e}, float kernel {err0:.3e}")
        assert err < max(1e-4 * scale, 1.2 * dev)
    keys = [str(k) for k in g["energy_keys"]]
    for r in range(len(E)):
        for c, k in enumerate(keys):
            e_ref = g["energies_f64"][r, c]
            assert abs(E[r][k] - e_ref) <= 1e-5 * abs(e_ref) + 2e-3, (k, E[r][k], e_ref)
    if "pairs_f32" in g:
        assert np.array_equal(f.neighbour_pairs(pos, box).cpu().numpy(), g["pairs_f32"])


@pytest.mark.parametrize("mode", [1, 2])
def test_fixed_point_kernels_on_a_3000_atom_water_box(mode):
    """A generated 1000-water box (L = 31 A: eligible with the default skin) against the oracle."""
    from torchmd_b200 import Forces, System, testsystems

    sysd = testsystems.water_box(1000, seed=1)
    terms = ["lj", "electrostatics", "bonds", "angles"]
    cfg = dict(cutoff=9.0, rfa=True, switch_dist=7.5)
    os.environ["TMD_B200_FX"] = str(mode)
    try:
        par = testsystems.water_parameters(sysd, device=DEV)
        system = System(len(sysd["coords"]), 1, torch.float32, DEV)
        system.set_positions(sysd["coords"])
        system.set_box(sysd["box"])
        f = Forces(par, terms=terms, **cfg)
        e = f.compute(system.pos, system.box, system.forces, returnDetails=True)[0]
        assert _lib_mod().lib().tmd_pair_kernel(f._ctx) == mode
        pairs = f.neighbour_pairs(system.pos, system.box).cpu().numpy()
    finally:
        os.environ.pop("TMD_B200_FX", None)
    of = refmd.OracleForces(testsystems.water_parameters(sysd, precision=torch.float64), terms, decision_dtype=torch.float32, **cfg)
    f64 = torch.zeros(system.pos.shape, dtype=torch.float64)
    e_ref = of.compute(system.pos.cpu().double(), system.box.cpu().double(), f64)[0]
    err = (system.forces.cpu().double() - f64).abs().max().item()
    print(f"3000-atom water, TMD_B200_FX={mode}: max|dF| vs fp64 oracle {err:.3e}, max|F| {f64.abs().max():.0f}")
    assert err < 1e-4 * max(1.0, f64.abs().max().item() / 100.0)
    for k in terms:
        assert abs(e[k] - float(e_ref[k])) <= 2e-5 * max(1.0, abs(float(e_ref[k]))) + 2e-3, k
    of32 = refmd.OracleForces(testsystems.water_parameters(sysd, precision=torch.float32), terms, **cfg)
    ref_pairs = of32.neighbour_pairs(system.pos[0].cpu(), torch.diagonal(system.box[0]).cpu()).numpy().astype(np.int32)
    assert np.array_equal(pairs, ref_pairs)


def test_fixed_point_kernel_is_accurate_for_drifted_molecules():
    """Molecules several boxes away from the primary cell: the float path loses ~1e-3 there
    (fl(L*n) for |n|>=3), the fixed-point path must stay below 1e-4."""
    from torchmd_b200 import Forces

    g = load_golden("water999_eq")
    cfg = golden_cfg(g)
    rng = np.random.default_rng(7)
    box = g["box"].astype(np.float64)
    shift = rng.integers(-3, 4, (len(g["coords"]) // 3, 3)).repeat(3, axis=0)
    coords = (g["coords"].astype(np.float64) + shift * box).astype(np.float32)
    par64 = params_from_golden(g, precision=torch.float64)
    terms = [str(t) for t in g["terms"]]
    of = refmd.OracleForces(par64, terms, decision_dtype=torch.float32, **cfg)
    pos_c = torch.tensor(coords)[None]
    box_c = torch.diag(torch.tensor(g["box"]))[None]
    f64 = torch.zeros(1, len(coords), 3, dtype=torch.float64)
    of.compute(pos_c.double(), box_c.double(), f64)
    of32 = refmd.OracleForces(params_from_golden(g, precision=torch.float32), terms, **cfg)
    pairs_ref = of32.neighbour_pairs(pos_c[0], torch.tensor(g["box"])).numpy().astype(np.int32)

    for mode in ("1", "2"):
        os.environ["TMD_B200_FX"] = mode
        try:
            f = Forces(params_from_golden(g, precision=torch.float32, device=DEV), terms=terms, skin=0.2, **cfg)
            pos, bx = pos_c.to(DEV), box_c.to(DEV)
            F = torch.zeros_like(pos)
            f.compute(pos, bx, F)
            if mode == "2":  # also the force-only instantiation
                from torchmd_b200 import _lib
                _lib.check(_lib.lib().tmd_forces(f._ctx, pos.data_ptr(), F.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
        finally:
            os.environ.pop("TMD_B200_FX", None)
        assert _lib_mod().lib().tmd_pair_kernel(f._ctx) == int(mode)
        err = (F.cpu().double() - f64).abs().max().item()
        print(f"drifted molecules, TMD_B200_FX={mode}: max|dF| vs fp64 oracle {err:.3e}")
        assert err < 1e-4
    assert np.array_equal(f.neighbour_pairs(pos, bx).cpu().numpy(), pairs_ref)


def test_fixed_point_trajectory_tracks_float_kernel():
    """100 NVE steps: same neighbour decisions, forces equal to ~1e-5, so the trajectories stay
    within the fp32 divergence of two correct implementations."""
    from torchmd_b200 import Forces, Integrator, System

    g = load_golden("water999_eq")
    cfg = golden_cfg(g)
    out = []
    for fx in (0, 1, 2):
        os.environ["TMD_B200_FX"] = str(fx)
        try:
            par = params_from_golden(g, precision=torch.float32, device=DEV)
            s = System(len(g["coords"]), 1, torch.float32, DEV)
            s.set_positions(g["coords"])
            s.set_box(g["box"])
            s.set_velocities(torch.tensor(g["vel"])[None])
            f = Forces(par, terms=[str(t) for t in g["terms"]], skin=0.2, **cfg)
            integ = Integrator(s, f, 1.0, DEV, gamma=None, T=None)
            ek, ep, T = integ.step(niter=100)
            out.append((s.pos.cpu().clone(), float(ek[0]), float(ep[0])))
        finally:
            os.environ.pop("TMD_B200_FX", None)
    for k in (1, 2):
        dp = (out[0][0] - out[k][0]).abs().max().item()
        print(f"NVE 100 steps: max |dpos| TMD_B200_FX={k} vs float kernel {dp:.3e}; Epot {out[0][2]:.4f} / {out[k][2]:.4f}")
        assert dp < 5e-3
        assert abs(out[0][2] - out[k][2]) < 0.05
