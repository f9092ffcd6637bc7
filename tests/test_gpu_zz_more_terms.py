"""Less common pair terms through the CUDA path: repulsion and repulsionCG (forces.py:418-450) alone
mixed with LJ (the reference builds the A/B tables only when "lj" is enabled, forces.py:45-46, so that
is the runnable combination) -- the run-time-flag variant of the pair kernel.  Golden vectors from the
unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
import test_gpu_forces
from test_gpu_forces import force_tol, run_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["argon100_lj_rep_mix", "argon100_lj_rep_nocut"])
def test_repulsion_terms(name):
    g = load_golden(name)
    f, pos, box, F, E = run_gpu(g)
    ref = g["forces_f64"]
    err = np.abs(F.cpu().numpy().astype(np.float64) - ref).max()
    scale = max(np.abs(ref).max(), 1e-30)
    print(f"{name}: max|dF| {err:.3e} (max|F| {scale:.3e})")
    assert err <= max(force_tol(ref) * 1e-2, 2e-6 * scale)  # tiny forces: relative fp32 bound
    keys = [str(k) for k in g["energy_keys"]]
    for c, k in enumerate(keys):
        e_ref = g["energies_f64"][0, c]
        assert abs(E[0][k] - e_ref) <= 1e-5 * abs(e_ref) + 1e-9, (k, E[0][k], e_ref)
    pairs = f.neighbour_pairs(pos, box).cpu().numpy()
    assert len(pairs) == int(g["npairs_f32"])
    if "pairs_f32" in g:
        assert np.array_equal(pairs, g["pairs_f32"])


def test_device_cell_grid_matches_host_description():
    """The grid the library builds on the device equals torchmd_b200.neighbourlist.cell_grid."""
    from torchmd_b200 import Forces, System, testsystems
    from torchmd_b200.neighbourlist import cell_grid, neighbour_list

    sysd = testsystems.water_box(1000, seed=5)
    par = testsystems.water_parameters(sysd, device=test_gpu_forces.DEV)
    system = System(len(sysd["coords"]), 1, torch.float32, test_gpu_forces.DEV)
    system.set_positions(sysd["coords"])
    system.set_box(sysd["box"])
    f = Forces(par, terms=["lj", "electrostatics"], cutoff=9.0, rfa=True, switch_dist=7.5, skin=1.0)
    pairs = neighbour_list(f, system.pos, system.box)
    want = cell_grid([float(x) for x in sysd["box"]], cutoff=9.0, skin=1.0)
    assert f.stats()["ncells"] == want["ncell"]
    assert pairs.shape[1] == 2 and bool((pairs[:, 0] < pairs[:, 1]).all())
