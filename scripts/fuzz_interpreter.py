"""Differential fuzzing of the kernels in the host SIMT interpreter against the CPU oracle (no GPU): random small
systems -- non-cubic boxes, few cells per dimension, atoms outside the box, random exclusions, with and without
switching / reaction field / cutoff, several replicas with different configurations -- through torchmd_b200.Forces
(the real host path) on the interpreter build.  Pairs must be bit-exact, forces within the parity tolerance.

    python scripts/fuzz_interpreter.py [ncases] [first_seed] [variant-tag]     # e.g. 200 0 _r2

Environment switches (each leaves the default sequence of cases unchanged): BONDED=1 random angles / dihedrals / impropers /
1-4 pairs; FAR=<k> atoms up to k boxes away (default 1); STRICT=1 plain 1e-4 yardstick; REPBOX=1 every replica in its own box;
NOCUT=1 periodic boxes without a cutoff; VERBOSE=1 print passing cases too.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def install(tag):
    import test_simt_kernels as T
    from torchmd_b200 import _lib

    class _Stream:
        cuda_stream = None

    _lib._lib = T.load(T.build_simt(tag, T.VARIANTS[tag]))
    _lib.on_device = lambda t: True
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.current_device = lambda: 0
    torch.cuda.synchronize = lambda *a, **k: None


def one_case(seed):
    from oracle import refmd
    from torchmd_b200 import Forces
    from torchmd_b200.parameters import TopologyParameters

    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 160))
    nrep = int(rng.integers(1, 4))
    periodic = rng.random() < 0.75
    has_cut = periodic or rng.random() < 0.6
    cutoff = float(rng.uniform(3.5, 9.0)) if has_cut else None
    L = rng.uniform(2.05 * (cutoff or 5.0), 2.05 * (cutoff or 5.0) + 25.0, size=3) if periodic else np.zeros(3)
    if os.environ.get("NOCUT") == "1":  # periodic boxes without a cutoff: every pair at its minimum image
        cutoff = None
    extent = L if periodic else np.full(3, rng.uniform(8.0, 30.0))
    n = max(2, min(n, int(np.prod(extent) / 80.0)))  # room for the minimum distance below
    # positions: minimum distance 2.6 A (contacts down to 0.7 sigma) inside the extent, then some atoms moved whole box lengths away
    pos = np.zeros((nrep, n, 3))
    # REPBOX=1 in the environment: every replica in its own box (up to 15 % longer per axis; own generator, so that the
    # default sequence of cases does not move)
    Ls = np.tile(L, (nrep, 1))
    if periodic and os.environ.get("REPBOX") == "1":
        Ls[1:] *= 1.0 + 0.15 * np.random.default_rng(seed + 10**6).random((nrep - 1, 3))
    for r in range(nrep):
        L = Ls[r]
        if periodic:
            extent = L
        pts = []
        while len(pts) < n:
            p = rng.uniform(0, extent)
            d = np.array(pts) - p if pts else np.zeros((0, 3))
            if periodic:
                d = d - L * np.round(d / L)
            if len(pts) == 0 or np.min(np.linalg.norm(d, axis=1)) > 2.6:
                pts.append(p)
        pos[r] = np.array(pts)
        if periodic:
            far = rng.random(n) < 0.15
            # one box either way by default; FAR=<k> in the environment moves atoms up to k boxes away (image counts of
            # 3, 5, 6, 7 ... for which fl(L * count) is inexact: the float kernel's VALUES use the unrounded product,
            # physics.cuh straddle_value, so its forces stay at the 1e-4 yardstick where the reference's fp32 does not;
            # STRICT=1 drops the allowance for the reference's own fp32 deviation from the tolerance below)
            reach = int(os.environ.get("FAR", "1"))
            pos[r][far] += L * rng.integers(-reach, reach + 1, size=(int(far.sum()), 3))
    ntypes = int(rng.integers(1, 5))
    types = rng.integers(0, ntypes, size=n)
    sigma, eps = rng.uniform(2.0, 3.6, ntypes), rng.uniform(0.02, 0.3, ntypes)
    charges = rng.uniform(-0.8, 0.8, n) * (rng.random() < 0.8)
    nb = int(rng.integers(0, n))
    bonds = np.unique(np.sort(np.stack([rng.integers(0, n, nb), rng.integers(0, n, nb)], 1), axis=1), axis=0) if nb else np.zeros((0, 2), int)
    bonds = bonds[bonds[:, 0] != bonds[:, 1]]
    terms = ["lj", "electrostatics"]
    bonded = None
    if len(bonds):
        terms.append("bonds")
        bonded = (bonds, np.stack([np.arange(len(bonds)), np.zeros(len(bonds), int)], 1), np.array([[30.0, 2.5]]))
    # BONDED=1 in the environment: random angles, multi-term dihedrals (AMBER cosine series, or a harmonic CHARMM term
    # where every period of the list is zero), impropers of either kind and scaled 1-4 pairs over random atom tuples
    angles = dihedrals = impropers = pairs14 = None
    if os.environ.get("BONDED") == "1" and n >= 6:

        def tuples(count, k):
            out = [rng.choice(n, size=k, replace=False) for _ in range(count)]
            return np.unique(np.array(out, dtype=np.int64), axis=0)

        if rng.random() < 0.8:
            idx = tuples(int(rng.integers(1, n)), 3)
            prm = np.stack([rng.uniform(10.0, 60.0, 3), rng.uniform(1.6, 2.2, 3)], 1)
            angles = (idx, np.stack([np.arange(len(idx)), rng.integers(0, 3, len(idx))], 1), prm)
            terms.append("angles")

        def torsion_set(count):
            idx = tuples(count, 4)
            harmonic = rng.random() < 0.35  # the reference switches on `all periods > 0` for the WHOLE list
            if harmonic:
                prm = np.stack([rng.uniform(5.0, 50.0, 3), rng.uniform(-np.pi, np.pi, 3), np.zeros(3)], 1)
                mp = np.stack([np.arange(len(idx)), rng.integers(0, 3, len(idx))], 1)
            else:
                prm = np.stack([rng.uniform(0.05, 2.0, 5), rng.choice([0.0, np.pi], 5), rng.integers(1, 5, 5).astype(float)], 1)
                rows = np.repeat(np.arange(len(idx)), rng.integers(1, 4, len(idx)))  # up to three series terms per dihedral
                mp = np.stack([rows, rng.integers(0, 5, len(rows))], 1)
            return idx, mp, prm

        if rng.random() < 0.8:
            dihedrals = torsion_set(int(rng.integers(1, n)))
            terms.append("dihedrals")
            if rng.random() < 0.7:
                idx = np.unique(np.sort(dihedrals[0][:, [0, 3]], axis=1), axis=0)
                prm = np.stack([rng.uniform(1e4, 6e5, 3), rng.uniform(50.0, 700.0, 3), np.full(3, 2.0), np.full(3, 1.2)], 1)
                pairs14 = (idx, np.stack([np.arange(len(idx)), rng.integers(0, 3, len(idx))], 1), prm)
                terms.append("1-4")
        if rng.random() < 0.6:
            impropers = torsion_set(int(rng.integers(1, max(2, n // 3))))
            terms.append("impropers")
    switch = float(rng.uniform(0.5 * cutoff, 0.95 * cutoff)) if cutoff and rng.random() < 0.6 else None
    rfa = bool(cutoff and rng.random() < 0.6)
    skin = float(rng.choice([0.0, 0.3, 1.0, 2.0]))

    def params(prec):
        return TopologyParameters(atom_types=types, type_sigma=sigma, type_epsilon=eps, charges=charges.astype(np.float32),
                                  masses=np.full(n, 12.0, np.float32), bonds=bonded, angles=angles, dihedrals=dihedrals,
                                  impropers=impropers, pairs14=pairs14, precision=prec, device="cpu")

    cfg = dict(cutoff=cutoff, rfa=rfa, switch_dist=switch)
    f = Forces(params(torch.float32), terms=terms, skin=skin, **cfg)
    p32 = torch.tensor(pos, dtype=torch.float32)
    box = torch.zeros(nrep, 3, 3)
    for r in range(nrep):
        for k in range(3):
            box[r, k, k] = float(Ls[r, k])
    L = Ls[0]
    F = torch.zeros_like(p32)
    E = f.compute(p32, box, F, returnDetails=True)
    of = refmd.OracleForces(params(torch.float64), terms, decision_dtype=torch.float32, **cfg)
    p64 = p32.double()
    F64 = torch.zeros_like(p64)
    E64 = of.compute(p64, box.double(), F64)
    fmax = float(F64.abs().max())
    err = float((F.double() - F64).abs().max())
    # the yardstick of tests/test_gpu_forces.py: 1e-4 scaled with the force magnitude, or the reference's own fp32
    # deviation where that is larger (atoms whole boxes away cost the reference -- and the float kernel -- bits)
    of32f = refmd.OracleForces(params(torch.float32), terms, **cfg)
    F32 = torch.zeros_like(p32)
    of32f.compute(p32, box, F32)
    dev = float((F32.double() - F64).abs().max())
    tol = 1e-4 * max(1.0, fmax / 100.0)
    if os.environ.get("STRICT") != "1":
        tol = max(tol, 1.2 * dev)
    ok = err < tol
    for r in range(nrep):
        for k in terms:
            ok = ok and abs(E[r][k] - E64[r][k]) <= 1e-5 * abs(E64[r][k]) + 2e-3
    of32 = refmd.OracleForces(params(torch.float32), terms, **cfg)
    for r in range(nrep):
        want = of32.neighbour_pairs(p32[r], torch.diagonal(box[r])).numpy().astype(np.int32)
        got = f.neighbour_pairs(p32, box, replica=r).cpu().numpy() if "replica" in f.neighbour_pairs.__code__.co_varnames else None
        if got is not None:
            ok = ok and got.shape == want.shape and np.array_equal(got, want)
    if len(terms) > 3:
        desc_terms = " terms=" + ",".join(terms[2:])
    else:
        desc_terms = ""
    desc = f"seed {seed}:{desc_terms} n={n} R={nrep} periodic={periodic} L={np.round(L, 2)} cutoff={cutoff} switch={switch} rfa={rfa} skin={skin} bonds={len(bonds)}"
    return ok, desc + f" | max|dF| {err:.2e} (tol {tol:.1e}, max|F| {fmax:.1f})"


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    install(sys.argv[3] if len(sys.argv) > 3 else "")
    bad = 0
    for s in range(first, first + ncases):
        try:
            ok, desc = one_case(s)
        except Exception as e:  # noqa: BLE001
            ok, desc = False, f"seed {s}: raised {type(e).__name__}: {e}"
        if not ok:
            bad += 1
            print("FAIL", desc, flush=True)
        elif os.environ.get("VERBOSE") == "1":
            print("ok  ", desc, flush=True)
    print(f"{ncases - bad} of {ncases} cases passed")
    sys.exit(1 if bad else 0)
