"""Golden vectors for Wrapper.wrap FROM THE UNMODIFIED REFERENCE (torchmd/wrapper.py).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_wrap.py

Build container only (needs /root/reference and networkx).  Writes wrap_cases.npz: inputs,
the reference's output, and the reference's molecule groups; asserts that the oracle
restatement (oracle/refmd.py wrap_positions / molecule_groups) reproduces both bit for bit.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)

from torchmd.wrapper import Wrapper as RefWrapper  # noqa: E402  (reference)

from oracle import refmd  # noqa: E402


def run_case(name, natoms, bonds, pos, box_diag, out):
    pos = torch.tensor(pos, dtype=torch.float32)
    nrep = pos.shape[0]
    box = torch.zeros(nrep, 3, 3, dtype=torch.float32)
    for r in range(nrep):
        box[r] = torch.diag(torch.tensor(box_diag[r], dtype=torch.float32))
    w = RefWrapper(natoms, bonds, "cpu")
    after = pos.clone()
    w.wrap(after, box)
    # oracle == reference, bit for bit
    groups, single = refmd.molecule_groups(natoms, bonds)
    assert sorted(map(tuple, (sorted(g.tolist()) for g in w.groups))) == sorted(map(tuple, groups)), name
    assert sorted(w.nongrouped.tolist()) == sorted(single), name
    mine = pos.clone()
    refmd.wrap_positions(mine, box, groups, single)
    nbad = int((mine != after).sum())
    assert nbad == 0, f"{name}: oracle differs from the reference in {nbad} coordinates"
    moved = int((after != pos).any(dim=2).sum())
    print(f"{name}: {natoms} atoms, {len(groups)} groups, {len(single)} single atoms, {moved} atom positions moved")
    out[name + "_natoms"] = np.int64(natoms)
    out[name + "_bonds"] = np.zeros((0, 2), np.int64) if bonds is None else np.asarray(bonds, np.int64)
    out[name + "_pos"] = pos.numpy()
    out[name + "_box"] = box.numpy()
    out[name + "_after"] = after.numpy()
    out[name + "_ngroups"] = np.int64(len(groups))
    out[name + "_nsingle"] = np.int64(len(single))


def main():
    out = {}
    rng = np.random.default_rng(11)
    # A: the equilibrated 999-atom water box, molecules thrown up to 3 boxes away, 2 replicas
    g = np.load(os.path.join(HERE, "water999_eq.npz"))
    coords, box = g["coords"].astype(np.float64), g["box"].astype(np.float64)
    n = len(coords)
    pos = np.stack([coords + rng.integers(-3, 4, (n // 3, 3)).repeat(3, axis=0) * box for _ in range(2)])
    pos[1] += rng.normal(scale=0.3, size=pos[1].shape)
    run_case("water", n, g["par_bond_idx"], pos, [box, box * [1.0, 0.9, 1.1]], out)
    # B: mixed topology in a rectangular box: a 300-atom chain with branches, 40 dimers, 25 ions
    nchain, ndim, nion = 300, 40, 25
    bonds = [(i, i + 1) for i in range(nchain - 1)] + [(i, i + 7) for i in range(0, nchain - 7, 50)]
    base = nchain
    for d in range(ndim):
        bonds.append((base + 2 * d, base + 2 * d + 1))
    natoms = nchain + 2 * ndim + nion
    box_b = np.array([31.7, 44.2, 27.9])
    walk = np.cumsum(rng.normal(scale=0.9, size=(nchain, 3)), axis=0) + rng.uniform(-60, 60, 3)
    dim = np.repeat(rng.uniform(-80, 120, (ndim, 3)), 2, axis=0) + rng.normal(scale=0.7, size=(2 * ndim, 3))
    ions = rng.uniform(-100, 150, (nion, 3))
    pos_b = np.concatenate([walk, dim, ions])[None]
    perm = rng.permutation(natoms)  # scramble the atom numbering: groups are not contiguous ranges
    inv = np.argsort(perm)
    bonds_p = np.array([(inv[i], inv[j]) for i, j in bonds])
    run_case("mixed", natoms, bonds_p, pos_b[:, perm], [box_b], out)
    # C: no bonds at all; D: all-zero box (nothing may move)
    run_case("nobonds", 64, None, rng.uniform(-50, 90, (1, 64, 3)), [np.array([20.0, 21.0, 22.0])], out)
    run_case("zerobox", 64, np.array([(0, 1), (1, 2), (5, 6)]), rng.uniform(-50, 90, (1, 64, 3)), [np.zeros(3)], out)
    path = os.path.join(HERE, "wrap_cases.npz")
    np.savez_compressed(path, **out)
    print(f"wrote wrap_cases.npz {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
