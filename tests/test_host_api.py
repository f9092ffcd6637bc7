"""Host-side contract, no GPU needed: the C-ABI library loads and exports every symbol
include/tmd_b200.h declares, the Python shells keep the reference's constructor/argument
semantics and error cases, and the CUDA-only rule is enforced loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_cfg, load_golden, params_from_golden


def declared_functions():
    text = open(os.path.join(ROOT, "include", "tmd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from torchmd_b200 import _lib

    names = declared_functions()
    assert len(names) >= 20
    handle = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/tmd_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "ctypes binding and header disagree"
    assert _lib.lib().tmd_version() >= 100
    assert _lib.NUM_ENERGIES == 9 and _lib.ENERGY_SLOTS.index("lj") == 6


def test_c_abi_argument_errors_without_gpu():
    """Bad arguments are reported through return codes + tmd_last_error, never abort."""
    from torchmd_b200 import _lib

    L = _lib.lib()
    h = C.c_void_p()
    assert L.tmd_create(C.byref(h), 0, 0, 1) < 0  # natoms must be > 0
    assert b"tmd_create" in L.tmd_last_error()
    assert L.tmd_forces(None, None, None, None, None) < 0
    assert L.tmd_destroy(None) == 0


def test_system_setters_follow_reference_semantics():
    from torchmd_b200 import System

    s = System(4, 2, torch.float32, "cpu")
    assert s.natoms == 4 and s.nreplicas == 2
    assert s.pos.shape == (2, 4, 3) and s.box.shape == (2, 3, 3) and s.masses.shape == (4, 1)
    xyz = np.arange(12, dtype=np.float32).reshape(4, 3)
    s.set_positions(xyz)  # (N,3) broadcast to both replicas
    assert torch.equal(s.pos[0], s.pos[1]) and torch.equal(s.pos[0], torch.tensor(xyz))
    s.set_positions(np.stack([xyz, xyz + 1], axis=2))  # (N,3,R)
    assert torch.equal(s.pos[1], torch.tensor(xyz + 1))
    with pytest.raises(RuntimeError):
        s.set_positions(np.zeros((4, 2)))
    s.set_box(np.array([10.0, 11.0, 12.0]))
    assert torch.equal(torch.diagonal(s.box[1]), torch.tensor([10.0, 11.0, 12.0]))
    assert s.box[0, 0, 1] == 0
    s.set_box(np.array([[1.0, 4.0], [2.0, 5.0], [3.0, 6.0]]))  # (3,R)
    assert torch.equal(torch.diagonal(s.box[1]), torch.tensor([4.0, 5.0, 6.0]))
    with pytest.raises(RuntimeError):
        s.set_box(np.zeros(2))
    with pytest.raises(RuntimeError):
        s.set_box(np.zeros((2, 2)))
    with pytest.raises(RuntimeError):
        s.set_velocities(torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError):
        s.set_forces(np.zeros((2, 4, 2)))
    with pytest.raises(RuntimeError):
        s.set_masses(torch.zeros(3))
    s.set_masses(torch.tensor([1.0, 2.0, 3.0, 4.0]))
    assert s.masses[:, 0].tolist() == [1.0, 2.0, 3.0, 4.0]
    s.precision_(torch.float64)
    assert s.pos.dtype == torch.float64


def test_forces_constructor_and_cuda_only_rule():
    from torchmd_b200 import Forces

    g = load_golden("water291_rf_switch")
    par = params_from_golden(g)
    with pytest.raises(RuntimeError):
        Forces(par, terms=None)
    with pytest.raises(ValueError):
        Forces(par, terms=["lj", "nonsense"])
    with pytest.raises(RuntimeError):
        Forces(par, terms=["1-4"])  # needs dihedrals
    f = Forces(par, terms=["LJ", "Electrostatics", "Bonds"], **golden_cfg(g))
    assert f.energies == ["lj", "electrostatics", "bonds"] and f.require_distances
    assert f.cutoff == 7.3 and f.rfa and f.switch_dist == 6.0 and f.natoms == 291
    assert Forces.terms == Forces.bonded + Forces.nonbonded
    pos = torch.zeros(1, 291, 3)
    with pytest.raises(RuntimeError, match="CUDA"):  # no CPU path, no silent fallback
        f.compute(pos, torch.zeros(1, 3, 3), torch.zeros_like(pos))
    with pytest.raises(RuntimeError):  # same check as the reference (forces.py:94-98)
        f.compute(pos, torch.zeros(1, 3, 3), torch.zeros_like(pos), explicit_forces=False)
    # the lazily built pair table equals the reference's definition
    ava = f.ava_idx.numpy()
    excl = {tuple(sorted(p)) for p in par.get_exclusions(("bonds", "angles", "1-4"))}
    assert len(ava) == 291 * 290 // 2 - len(excl)
    assert (ava[:, 0] < ava[:, 1]).all() and not any((a, b) in excl for a, b in ava[:2000])


def test_exclusion_csr_is_symmetric_and_sorted():
    from torchmd_b200.forces import _exclusion_csr

    ptr, cols = _exclusion_csr(6, [[0, 1], [1, 0], [4, 2], [2, 5], [3, 3]])
    assert ptr.tolist() == [0, 1, 2, 4, 4, 5, 6]
    assert cols.tolist() == [1, 0, 4, 5, 2, 2]
    ptr, cols = _exclusion_csr(3, [])
    assert ptr.tolist() == [0, 0, 0, 0] and len(cols) == 0


def test_integrator_host_helpers_match_reference_definitions():
    """kinetic_energy / kinetic_to_temp / maxwell_boltzmann (integrator.py:8-58) on CPU tensors."""
    from torchmd_b200 import kinetic_energy, kinetic_to_temp, maxwell_boltzmann

    masses = torch.tensor([[1.0], [2.0], [3.0]])
    vel = torch.tensor([[[1.0, 0, 0], [0, 2.0, 0], [0, 0, 3.0]], [[1.0, 1, 1], [1, 1, 1], [0, 0, 0]]])
    ke = kinetic_energy(masses, vel)
    assert ke.shape == (2, 1)
    np.testing.assert_allclose(ke[:, 0].numpy(), [0.5 * 1 + 0.5 * 2 * 4 + 0.5 * 3 * 9, 0.5 * 3 + 0.5 * 2 * 3], rtol=1e-6)
    batch = torch.tensor([0, 0, 1])
    kb = kinetic_energy(masses, vel, batch)
    assert kb.shape == (2, 2)
    np.testing.assert_allclose(kb.numpy(), [[4.5, 13.5], [4.5, 0.0]], rtol=1e-6)
    with pytest.raises(ValueError):
        kinetic_energy(masses, vel[0])
    np.testing.assert_allclose(kinetic_to_temp(np.array([3.0]), 2), 2.0 / (3 * 2 * 0.001987191) * 3.0)
    torch.manual_seed(0)
    v = maxwell_boltzmann(torch.full((20000, 1), 4.0), 300.0, replicas=2)
    assert v.shape == (2, 20000, 3)
    assert abs(v.var().item() - 300.0 * 0.001987191 / 4.0) < 0.003


def test_integrator_requires_cuda_state():
    from torchmd_b200 import Integrator, System

    s = System(2, 1, torch.float32, "cpu")
    s.set_masses(torch.tensor([1.0, 2.0]))

    class Mock:
        def compute(self, pos, box, forces):
            return 0.0

    integ = Integrator(s, Mock(), 1.0, "cpu", gamma=0.1, T=300.0)
    assert abs(integ.dt - 1.0 / 48.88821) < 1e-12 and integ.T == 300.0
    assert abs(integ.gamma - 0.1 / (1000.0 / 48.88821)) < 1e-15
    with pytest.raises(RuntimeError, match="CUDA"):
        integ.step(1)


def test_synthetic_inputs_are_deterministic():
    from torchmd_b200 import testsystems

    a, b = testsystems.water_box(64, seed=3), testsystems.water_box(64, seed=3)
    assert np.array_equal(a["coords"], b["coords"]) and a["coords"].shape == (192, 3)
    assert not np.array_equal(a["coords"], testsystems.water_box(64, seed=4)["coords"])
    d = np.linalg.norm(a["coords"][0::3] - a["coords"][1::3], axis=1)
    np.testing.assert_allclose(d, 0.9572, atol=1e-5)
    assert abs(a["box"][0] - (64 / 0.0334) ** (1 / 3)) < 1e-4
    par = testsystems.water_parameters(a)
    assert par.bond_params["idx"].shape == (128, 2) and par.angle_params["idx"].shape == (64, 3)
    assert len(par.get_exclusions()) == 128 + 64
    ar = testsystems.argon_box(50, seed=1)
    assert ar["coords"].shape == (50, 3)


def test_cell_grid_matches_the_library_rules():
    from torchmd_b200.neighbourlist import cell_grid

    g = cell_grid([99.93, 99.93, 99.93], cutoff=9.0, skin=1.0)
    assert g["ncell"] == (19, 19, 19) and g["reach"] == (2, 2, 2)
    assert all(w >= g["rlist"] / 2 for w in g["width"])
    small = cell_grid([16.919, 16.633, 16.639], cutoff=7.3)  # tests/water: too short for 5 cells
    assert small["ncell"] == (1, 1, 1) and small["reach"] == (0, 0, 0)
    slab = cell_grid([200.0, 30.0, 12.0], cutoff=9.0, skin=1.0)
    assert slab["ncell"][0] == 39 and slab["ncell"][1] == 5 and slab["ncell"][2] == 1
    # the reference stub would leave a 0.93 A sliver as the last bin of a 99.93 A box;
    # here every cell has the same width and the widths tile the box
    assert abs(g["width"][0] * g["ncell"][0] - 99.93) < 1e-9
