"""Deterministic synthetic inputs for the parity tests and ``bench.py``.

SURVEY.md section 8d fixes the generators so the GPU run, the CPU oracle and the
golden fixtures all see the same atoms:

* ``water_box(n_waters, seed)`` -- flexible TIP3P water at 0.0334 molecules/A^3,
  the parameters of the reference's ``tests/water/water_forcefield.yaml``
  (OT/HT, CHARMM-style negative epsilons kept as they are).
* ``argon_box(n_atoms, seed)`` -- LJ-only argon at the density of the
  reference's ``tests/argon/argon_start.pdb`` fixture (100 atoms in 77.395^3 A^3).

Both return a plain dict of numpy arrays plus a ``TopologyParameters`` factory,
so neither moleculekit nor the reference checkout is needed.
"""
import math
import os

import numpy as np
import torch

from .parameters import TopologyParameters

# tests/water/water_forcefield.yaml (reference), atom types sorted like
# np.unique does in torchmd/parameters.py:110 -> HT=0, OT=1
WATER_TYPES = ("HT", "OT")
WATER_SIGMA = (0.40001352444501237, 3.150574226831496)
WATER_EPSILON = (-0.046, -0.1521)
WATER_CHARGE = {"OT": -0.834, "HT": 0.417}
WATER_MASS = {"OT": 15.9994, "HT": 1.008}
WATER_BOND = (450.0, 0.9572)  # (OT, HT): k0, req
WATER_HH_BOND = (0.0, 1.5139)  # (HT, HT)
WATER_ANGLE = (55.0, math.radians(104.52))  # (HT, OT, HT): k0, theta0
WATER_DENSITY = 0.0334  # molecules / A^3

# tests/argon/argon_forcefield.yaml (reference)
ARGON_SIGMA = 3.345
ARGON_EPSILON = 0.238
ARGON_MASS = 39.95
ARGON_DENSITY = 100.0 / 77.395**3


def _random_rotations(rng, n):
    """n uniformly distributed rotation matrices (from unit quaternions)."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack(
        [
            np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
            np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
            np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
        ],
        axis=1,
    )


def water_box(n_waters, seed=0, hh_bonds=False):
    """Cubic box of ``n_waters`` TIP3P molecules, atom order O,H,H per molecule.

    Oxygens sit on a seeded random subset of an m^3 lattice (m = ceil(n^(1/3)))
    with +-0.2 A jitter; each rigid molecule gets a seeded random orientation.
    The start is a lattice, not a liquid: equilibrate before measuring.
    """
    rng = np.random.default_rng(seed)
    L = (n_waters / WATER_DENSITY) ** (1.0 / 3.0)
    m = int(math.ceil(n_waters ** (1.0 / 3.0) - 1e-9))
    sites = np.sort(rng.permutation(m**3)[:n_waters])
    grid = np.stack(np.unravel_index(sites, (m, m, m)), axis=1).astype(np.float64)
    oxy = (grid + 0.5) * (L / m) + rng.uniform(-0.2, 0.2, size=(n_waters, 3))

    r_oh, ang = WATER_BOND[1], WATER_ANGLE[1]
    local = np.array(
        [
            [0.0, 0.0, 0.0],
            [r_oh * math.sin(ang / 2), r_oh * math.cos(ang / 2), 0.0],
            [-r_oh * math.sin(ang / 2), r_oh * math.cos(ang / 2), 0.0],
        ]
    )
    rot = _random_rotations(rng, n_waters)
    coords = oxy[:, None, :] + np.einsum("nij,aj->nai", rot, local)
    coords = coords.reshape(-1, 3).astype(np.float32)

    n = 3 * n_waters
    o = np.arange(0, n, 3)
    names = np.tile(np.array(["OT", "HT", "HT"], dtype=object), n_waters)
    bonds = np.concatenate([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1)])
    if hh_bonds:
        bonds = np.concatenate([bonds, np.stack([o + 1, o + 2], 1)])
    angles = np.stack([o + 1, o, o + 2], 1)
    return {
        "name": f"water{n}",
        "coords": coords,
        "box": np.array([L, L, L], dtype=np.float32),
        "atomtype": names,
        "charge": np.array([WATER_CHARGE[t] for t in names], dtype=np.float32),
        "masses": np.array([WATER_MASS[t] for t in names], dtype=np.float32),
        "bonds": bonds.astype(np.int64),
        "angles": angles.astype(np.int64),
    }


def water_parameters(sysd, precision=torch.float32, device="cpu"):
    """``TopologyParameters`` for a ``water_box`` dict.

    Row order follows the reference builder (``torchmd/parameters.py:165-205``):
    bonds unique-sorted with i<j, angles with first<last, parameter rows in
    order of first appearance.
    """
    types = np.array([WATER_TYPES.index(t) for t in sysd["atomtype"]])
    bonds = np.unique(np.sort(sysd["bonds"], axis=1), axis=0)
    is_hh = (types[bonds[:, 0]] == 0) & (types[bonds[:, 1]] == 0)
    bond_rows = [WATER_BOND]
    bond_map = np.zeros(len(bonds), dtype=np.int64)
    if is_hh.any():
        bond_rows.append(WATER_HH_BOND)
        bond_map[is_hh] = 1
    ang = sysd["angles"].copy()
    flip = ang[:, 0] > ang[:, 2]
    ang[flip] = ang[flip][:, ::-1]
    ang = np.unique(ang, axis=0)
    return TopologyParameters(
        atom_types=types,
        type_sigma=WATER_SIGMA,
        type_epsilon=WATER_EPSILON,
        charges=sysd["charge"],
        masses=sysd["masses"],
        bonds=(bonds, np.stack([np.arange(len(bonds)), bond_map], 1), bond_rows),
        angles=(ang, np.stack([np.arange(len(ang)), np.zeros(len(ang), np.int64)], 1), [WATER_ANGLE]),
        precision=precision,
        device=device,
    )


def argon_box(n_atoms, seed=0, min_dist=3.4):
    """``n_atoms`` argon atoms, uniform random with a minimum separation."""
    rng = np.random.default_rng(seed)
    L = (n_atoms / ARGON_DENSITY) ** (1.0 / 3.0)
    pts = np.empty((0, 3))
    while len(pts) < n_atoms:
        cand = rng.uniform(0, L, size=(n_atoms, 3))
        for c in cand:
            if len(pts) == n_atoms:
                break
            if len(pts):
                d = pts - c
                d -= L * np.round(d / L)
                if (np.einsum("ij,ij->i", d, d) < min_dist**2).any():
                    continue
            pts = np.vstack([pts, c])
    return {
        "name": f"argon{n_atoms}",
        "coords": pts.astype(np.float32),
        "box": np.array([L, L, L], dtype=np.float32),
        "atomtype": np.array(["AR"] * n_atoms, dtype=object),
        "charge": np.zeros(n_atoms, dtype=np.float32),
        "masses": np.full(n_atoms, ARGON_MASS, dtype=np.float32),
        "bonds": np.zeros((0, 2), dtype=np.int64),
        "angles": np.zeros((0, 3), dtype=np.int64),
    }


def argon_parameters(sysd, precision=torch.float32, device="cpu"):
    n = len(sysd["coords"])
    return TopologyParameters(
        atom_types=np.zeros(n, dtype=np.int64),
        type_sigma=[ARGON_SIGMA],
        type_epsilon=[ARGON_EPSILON],
        charges=sysd["charge"],
        masses=sysd["masses"],
        precision=precision,
        device=device,
    )


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_system(name, precision=torch.float32, device="cpu"):
    """A system of the committed parity fixtures (``tests/golden/<name>.npz``: coordinates, box, run settings and the
    parameter tables the reference's ``Parameters`` held when the fixture was made) as
    ``(TopologyParameters, coords (N,3) float32, box (3,) float32, terms, cfg)`` -- the reference's own test systems
    (alanine dipeptide, thrombin-ligand, the water fixture) for ``bench.py --workload`` without the reference checkout."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))

    def term(key):
        if "par_" + key + "_idx" not in g:
            return None
        return (g["par_" + key + "_idx"], g["par_" + key + "_map"], g["par_" + key + "_params"])

    se = g["par_lj_sigma_eps"]
    par = TopologyParameters(
        atom_types=g["par_types"], type_sigma=se[:, 0], type_epsilon=se[:, 1], charges=g["par_charges"], masses=g["par_masses"],
        bonds=term("bond"), angles=term("angle"), dihedrals=term("dihedral"), impropers=term("improper"),
        pairs14=term("nonbonded_14"), precision=precision, device=device,
    )

    def opt(k):
        v = float(g["cfg_" + k])
        return None if np.isnan(v) else v

    cfg = dict(cutoff=opt("cutoff"), rfa=bool(g["cfg_rfa"]), switch_dist=opt("switch_dist"))
    return par, np.asarray(g["coords"], dtype=np.float32), np.asarray(g["box"], dtype=np.float32), [str(t) for t in g["terms"]], cfg
