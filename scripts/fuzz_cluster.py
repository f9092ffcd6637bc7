"""Differential fuzzing of the CLUSTER path (torchmd_b200/csrc/cluster.cuh) in the host SIMT interpreter against the CPU
oracle (no GPU): jittered, partly filled lattices dense enough for the cluster lists -- periodic boxes just above the
size the path needs and open systems -- with random types, charges, local bonds (exclusions inside and across
clusters), atoms moved whole boxes away, one or two replicas; after the first evaluation the atoms are moved a little
(list reuse) and a lot (rebuild).  Pairs must be bit-exact every time, forces within the parity tolerance (3x relaxed
after the moves: the large move produces contacts with |F| > 1000, where fp32 is at 1e-6 relative).

    python scripts/fuzz_cluster.py [ncases] [first_seed]        # VERBOSE=1 prints passing cases and the kernel they took

scripts/fuzz_interpreter.py is the companion for small sparse systems (mostly the full-row kernels).
"""
import collections
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def one(seed):
    from oracle import refmd
    from torchmd_b200 import Forces, _lib
    from torchmd_b200.parameters import TopologyParameters

    rng = np.random.default_rng(seed)
    periodic = rng.random() < 0.7
    cutoff = float(rng.uniform(3.5, 5.5))
    skin = float(rng.choice([0.0, 0.3, 0.6]))
    rl = cutoff + skin
    nrep = int(rng.integers(1, 3))
    if periodic:
        L = rng.uniform(2 * (rl + 8.0 + skin) + 0.3, 2 * (rl + 8.0 + skin) + 9.0, size=3)
    else:
        L = rng.uniform(14.0, 30.0, size=3)
    a = float(rng.uniform(2.6, 3.4))
    g = [max(2, int(L[k] / a)) for k in range(3)]
    grid = np.stack(np.meshgrid(*[np.arange(m) for m in g], indexing="ij"), -1).reshape(-1, 3).astype(float)
    grid = grid * (L / np.array(g)) + 0.3
    keep = rng.random(len(grid)) < rng.uniform(0.7, 1.0)
    base = grid[keep]
    n = len(base)
    order = rng.permutation(n)
    base = base[order]
    pos = np.stack([base + rng.normal(0, 0.35, base.shape) for _ in range(nrep)])
    Ls = np.tile(L if periodic else np.zeros(3), (nrep, 1))
    if periodic:
        far = rng.random(n) < 0.1
        for r in range(nrep):
            pos[r][far] += L * rng.integers(-2, 3, size=(int(far.sum()), 3))
    ntypes = int(rng.integers(1, 7))
    types = rng.integers(0, ntypes, size=n)
    sigma, eps = rng.uniform(1.6, 2.6, ntypes), rng.uniform(0.02, 0.3, ntypes)
    charges = rng.uniform(-0.8, 0.8, n) * (rng.random() < 0.8)
    # bonds between spatial neighbours (exclusions stay local, like molecules) + a few random long ones
    nb = int(rng.integers(0, n))
    i = rng.integers(0, n, nb)
    d = np.linalg.norm(base[i][:, None, :] - base[None, :, :], axis=2) if nb and n < 1500 else None
    bonds = np.zeros((0, 2), int)
    if nb and d is not None:
        d[np.arange(nb), i] = 1e9
        j = np.argmin(d, axis=1)
        bonds = np.stack([i, j], 1)
        extra = np.stack([rng.integers(0, n, 5), rng.integers(0, n, 5)], 1)
        bonds = np.concatenate([bonds, extra])
        bonds = np.unique(np.sort(bonds, axis=1), axis=0)
        bonds = bonds[bonds[:, 0] != bonds[:, 1]]
    terms = ["lj", "electrostatics"]
    bonded = None
    if len(bonds):
        terms.append("bonds")
        bonded = (bonds, np.stack([np.arange(len(bonds)), np.zeros(len(bonds), int)], 1), np.array([[30.0, 2.5]]))
    switch = float(rng.uniform(0.5 * cutoff, 0.95 * cutoff)) if rng.random() < 0.6 else None
    rfa = bool(rng.random() < 0.6)
    def params(prec):
        return TopologyParameters(atom_types=types, type_sigma=sigma, type_epsilon=eps, charges=charges.astype(np.float32),
                                  masses=np.full(n, 12.0, np.float32), bonds=bonded, precision=prec, device="cpu")
    cfg = dict(cutoff=cutoff, rfa=rfa, switch_dist=switch)
    f = Forces(params(torch.float32), terms=terms, skin=skin, **cfg)
    p32 = torch.tensor(pos, dtype=torch.float32)
    box = torch.zeros(nrep, 3, 3)
    for r in range(nrep):
        for k in range(3):
            box[r, k, k] = float(Ls[r, k])
    F = torch.zeros_like(p32)
    E = f.compute(p32, box, F, returnDetails=True)
    kern = int(_lib.lib().tmd_pair_kernel(f._ctx))
    of = refmd.OracleForces(params(torch.float64), terms, decision_dtype=torch.float32, **cfg)
    F64 = torch.zeros(nrep, n, 3, dtype=torch.float64)
    E64 = of.compute(p32.double(), box.double(), F64)
    fmax = float(F64.abs().max()); err = float((F.double() - F64).abs().max())
    of32 = refmd.OracleForces(params(torch.float32), terms, **cfg)
    F32 = torch.zeros_like(p32); of32.compute(p32, box, F32)
    dev = float((F32.double() - F64).abs().max())
    tol = max(1e-4 * max(1.0, fmax / 100.0), 1.2 * dev)
    ok = err < tol
    why = [] if ok else ['forces']
    for r in range(nrep):
        for k in terms:
            if not abs(E[r][k] - E64[r][k]) <= 1e-5 * abs(E64[r][k]) + 2e-3: ok = False; why.append(f'E {k} r{r}: {E[r][k]} vs {E64[r][k]}')
        want = of32.neighbour_pairs(p32[r], torch.diagonal(box[r])).numpy().astype(np.int32)
        got = f.neighbour_pairs(p32, box, replica=r).cpu().numpy()
        if not (got.shape == want.shape and np.array_equal(got, want)): ok = False; why.append(f'pairs r{r}: {got.shape} vs {want.shape}')
    # second evaluation after a small move (list reuse) and after a large one (rebuild)
    for amp in (0.05, 0.8):
        p2 = p32 + torch.tensor(rng.normal(0, amp, pos.shape), dtype=torch.float32)
        f.compute(p2, box, F)
        of.compute(p2.double(), box.double(), F64)
        of32.compute(p2, box, F32)
        dev2 = float((F32.double() - F64).abs().max()); fm = float(F64.abs().max())
        e2 = float((F.double() - F64).abs().max())
        if not e2 < max(3e-4 * max(1.0, fm / 100.0), 3.0 * dev2): ok = False; why.append(f'move {amp}: err {e2:.2e} fmax {fm:.1f} dev {dev2:.2e}')
        for r in range(nrep):
            want = of32.neighbour_pairs(p2[r], torch.diagonal(box[r])).numpy().astype(np.int32)
            got = f.neighbour_pairs(p2, box, replica=r).cpu().numpy()
            if not (got.shape == want.shape and np.array_equal(got, want)): ok = False; why.append(f'move {amp} pairs r{r}: {got.shape} vs {want.shape}')
        # where is the largest error, and how big is the force there?
        if e2 >= max(3e-4 * max(1.0, fm / 100.0), 3.0 * dev2):
            dd = (F.double() - F64).abs().amax(dim=2); idx = int(dd.argmax()); r_, a_ = divmod(idx, n)
            why.append(f'  worst atom r{r_} a{a_}: |F| {float(F64[r_, a_].abs().max()):.1f} dF {float(dd[r_, a_]):.2e} ref32 dF {float((F32.double()-F64)[r_, a_].abs().max()):.2e}')
    return ok, kern, f"seed {seed}: n={n} R={nrep} periodic={periodic} L={np.round(L,1)} cutoff={cutoff:.2f} switch={switch} rfa={rfa} skin={skin} bonds={len(bonds)} types={ntypes} | max|dF| {err:.2e} (tol {tol:.1e}, max|F| {fmax:.1f}) {why}"


if __name__ == "__main__":
    import fuzz_interpreter as Z

    Z.install("_cl")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    kinds = collections.Counter()
    bad = 0
    for s in range(first, first + n):
        try:
            ok, kern, desc = one(s)
        except Exception as e:  # noqa: BLE001
            ok, kern, desc = False, None, f"seed {s}: raised {type(e).__name__}: {e}"
        kinds[kern] += 1
        bad += not ok
        if not ok or os.environ.get("VERBOSE") == "1":
            print("FAIL" if not ok else "ok  ", "[kernel %s]" % kern, desc, flush=True)
    print(f"{n - bad} of {n} passed; pair kernels used: {dict(kinds)}")
    sys.exit(1 if bad else 0)
