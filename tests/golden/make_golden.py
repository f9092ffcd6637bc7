"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Run in the build container only (``/root/reference`` must exist):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the reference's own ``torchmd.forces/integrator/systems/parameters``
and ``forcefields.ff_yaml`` (pure torch, importable without moleculekit), drives
them with a duck-typed ``mol``, and writes small ``.npz``
files holding inputs *and* reference outputs, so the GPU box (which has no
reference checkout) can check both ``oracle/refmd.py`` and the CUDA path against
them.  While doing so it also asserts that the oracle restatement is
bit-identical to the reference on every case (fp32 and fp64) and records which
rounding sequence ``torch.norm`` uses for the cutoff predicate.

Nothing here is imported by the product or by the tests; the tests only read
the ``.npz`` outputs.
"""
import hashlib
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

from torchmd.forcefields.forcefield import ForceField  # noqa: E402  (reference)
from torchmd.forces import Forces as RefForces  # noqa: E402
from torchmd.integrator import Integrator as RefIntegrator  # noqa: E402
from torchmd.integrator import maxwell_boltzmann as ref_mb  # noqa: E402
from torchmd.parameters import Parameters as RefParameters  # noqa: E402
from torchmd.systems import System as RefSystem  # noqa: E402

from oracle import refmd  # noqa: E402
from torchmd_b200 import testsystems  # noqa: E402

torch.set_num_threads(8)


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def make_mol(d):
    """Duck-typed moleculekit.Molecule: just what Parameters/YamlForcefield read."""
    n = len(d["coords"])
    return types.SimpleNamespace(
        numAtoms=n,
        atomtype=np.asarray(d["atomtype"], dtype=object),
        charge=np.asarray(d["charge"], dtype=np.float32),
        masses=np.asarray(d["masses"], dtype=np.float32),
        bonds=np.asarray(d.get("bonds", np.zeros((0, 2), int))),
        angles=np.asarray(d.get("angles", np.zeros((0, 3), int))),
        dihedrals=np.asarray(d.get("dihedrals", np.zeros((0, 4), int))),
        impropers=np.asarray(d.get("impropers", np.zeros((0, 4), int))),
        coords=np.asarray(d["coords"], dtype=np.float32)[:, :, None],
        box=np.asarray(d["box"], dtype=np.float32)[:, None],
    )


def ref_setup(d, ff_file, terms, precision, nrep=1, **fkw):
    mol = make_mol(d)
    ff = ForceField.create(mol, ff_file)
    par = RefParameters(ff, mol, terms, precision=precision, device="cpu")
    sysm = RefSystem(mol.numAtoms, nrep, precision, "cpu")
    sysm.set_positions(mol.coords)
    sysm.set_box(mol.box)
    frc = RefForces(par, terms=terms, **fkw)
    return mol, par, sysm, frc


def pack_params(par):
    """Reference Parameters -> flat dict of arrays (the kernel input contract)."""
    out = {
        "types": par.mapped_atom_types.numpy().astype(np.int64),
        "charges": par.charges.double().numpy(),
        "masses": par.masses.double().numpy()[:, 0],
    }
    if par.nonbonded_params is not None:
        out["lj_sigma_eps"] = par.nonbonded_params["params"].double().numpy()
    for name in ("bond", "angle", "dihedral", "improper", "nonbonded_14"):
        t = getattr(par, name + "_params")
        if t is not None and len(t["idx"]):
            out[name + "_idx"] = t["idx"].numpy()
            out[name + "_map"] = t["map"].numpy()
            out[name + "_params"] = t["params"].double().numpy()
    return out


def oracle_matches(par, terms, sysm, fkw, ref_E, ref_F, ref_pairs):
    """Assert the restatement reproduces the reference bit for bit."""
    of = refmd.OracleForces(par, terms, **fkw)
    F = torch.zeros_like(sysm.pos)
    E = of.compute(sysm.pos, sysm.box, F)
    assert torch.equal(F, ref_F), f"oracle forces differ: {(F - ref_F).abs().max()}"
    for r in range(len(E)):
        for k, v in E[r].items():
            assert v == ref_E[r][k], (k, v, ref_E[r][k])
    if ref_pairs is not None:
        bd = torch.diagonal(sysm.box[0])
        assert torch.equal(of.neighbour_pairs(sysm.pos[0], bd), ref_pairs)


def ref_pairs_in_cutoff(frc, sysm):
    """The reference 'neighbour list': ava_idx[dist <= cutoff] (forces.py:264-269)."""
    from torchmd.forces import calculate_distances

    bd = sysm.box[0][torch.eye(3).bool()]
    dist, _, _ = calculate_distances(sysm.pos[0], frc.ava_idx, bd)
    if frc.cutoff is None:
        return frc.ava_idx, dist
    return frc.ava_idx[dist <= frc.cutoff], dist


def force_case(d, ff_file, terms, nrep=1, store_pairs=True, **fkw):
    """Run the reference in fp32 and fp64 on one configuration."""
    res = {}
    for tag, prec in (("f32", torch.float32), ("f64", torch.float64)):
        mol, par, sysm, frc = ref_setup(d, ff_file, terms, prec, nrep, **fkw)
        E = frc.compute(sysm.pos, sysm.box, sysm.forces, returnDetails=True)
        pairs = None
        if frc.require_distances:
            pairs, _ = ref_pairs_in_cutoff(frc, sysm)
        oracle_matches(par, terms, sysm, fkw, E, sysm.forces, pairs)
        keys = [k for k in E[0] if k != "external"]
        res[f"energy_keys"] = np.array(keys)
        res[f"energies_{tag}"] = np.array([[e[k] for k in keys] for e in E])
        res[f"forces_{tag}"] = sysm.forces.numpy().copy()
        if pairs is not None:
            p32 = pairs.numpy().astype(np.int32)
            res[f"npairs_{tag}"] = np.int64(len(p32))
            res[f"pairs_sha256_{tag}"] = np.array(hashlib.sha256(p32.tobytes()).hexdigest())
            if store_pairs and tag == "f32":
                res["pairs_f32"] = p32
        if tag == "f32":
            res.update({"par_" + k: v for k, v in pack_params(par).items()})
    res["coords"] = np.asarray(d["coords"], dtype=np.float32)
    res["box"] = np.asarray(d["box"], dtype=np.float32)
    res["terms"] = np.array(terms)
    for k, v in fkw.items():
        res["cfg_" + k] = np.array(np.nan if v is None else v)
    res["cfg_nrep"] = np.int64(nrep)
    return res


def save(name, res):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **res)
    print(f"wrote {name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


# ----------------------------------------------------------------------------
# fixture readers (reference test data -> arrays)
# ----------------------------------------------------------------------------
def read_pdb(path):
    xyz, box = [], None
    for line in open(path):
        if line.startswith("CRYST1"):
            box = [float(line[6:15]), float(line[15:24]), float(line[24:33])]
        elif line.startswith(("ATOM", "HETATM")):
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
        elif line.startswith("ENDMDL"):
            break
    return np.array(xyz, dtype=np.float32), np.array(box, dtype=np.float32)


def read_psf(path):
    lines = open(path).read().split("\n")
    sec = {}
    i = 0
    while i < len(lines):
        if "!N" in lines[i]:
            key = lines[i].split("!")[1].split(":")[0].split()[0]
            n = int(lines[i].split()[0])
            j = i + 1
            body = []
            while j < len(lines) and lines[j].strip():
                body.append(lines[j])
                j += 1
            sec[key] = (n, body)
            i = j
        else:
            i += 1
    atoms = [ln.split() for ln in sec["NATOM"][1]]
    flat = lambda k: np.array(" ".join(sec[k][1]).split(), dtype=np.int64) - 1  # noqa: E731
    return {
        "atomtype": np.array([a[5] for a in atoms], dtype=object),
        "charge": np.array([float(a[6]) for a in atoms], dtype=np.float32),
        "masses": np.array([float(a[7]) for a in atoms], dtype=np.float32),
        "bonds": flat("NBOND").reshape(-1, 2),
        "angles": flat("NTHETA").reshape(-1, 3),
    }


# ----------------------------------------------------------------------------
# a small molecule with every bonded term, from a YAML force field of our own
# ----------------------------------------------------------------------------
CHAIN_FF = """atomtypes: [C, CT, N, O]

bonds:
  (CT, CT): {k0: 310.0, req: 1.526}
  (CT, C): {k0: 317.0, req: 1.522}
  (C, O): {k0: 570.0, req: 1.229}
  (C, N): {k0: 490.0, req: 1.335}
  (N, CT): {k0: 337.0, req: 1.449}

angles:
  (X, CT, X): {k0: 50.0, theta0: 109.5}
  (X, C, X): {k0: 70.0, theta0: 120.0}
  (X, N, X): {k0: 50.0, theta0: 121.9}

dihedrals:
  (X, X, X, X):
    terms:
      - {phi_k: 0.156, phase: 0.0, per: 3}
    scnb: 2.0
    scee: 1.2
  (X, C, N, X):
    terms:
      - {phi_k: 2.5, phase: 180.0, per: 2}
      - {phi_k: 0.3, phase: 0.0, per: 1}
    scnb: 2.0
    scee: 1.2
  (CT, CT, C, N):
    terms:
      - {phi_k: 0.07, phase: 0.0, per: 2}
      - {phi_k: 0.1, phase: 0.0, per: 4}
      - {phi_k: 0.45, phase: 180.0, per: 1}
    scnb: 2.0
    scee: 1.2

impropers:
  (X, X, C, O): {phi_k: %(imp_k)s, phase: %(imp_phase)s, per: %(imp_per)s}

lj:
  C: {sigma: 3.3997, epsilon: 0.086, sigma14: 3.3997, epsilon14: 0.086}
  CT: {sigma: 3.3997, epsilon: 0.1094, sigma14: 3.3997, epsilon14: 0.1094}
  N: {sigma: 3.25, epsilon: 0.17, sigma14: 3.25, epsilon14: 0.17}
  O: {sigma: 2.96, epsilon: 0.21, sigma14: 2.96, epsilon14: 0.21}

electrostatics:
  C: {charge: 0.5973}
  CT: {charge: 0.0337}
  N: {charge: -0.4157}
  O: {charge: -0.5679}

masses: {C: 12.01, CT: 12.01, N: 14.01, O: 16.0}
"""


def chain_molecules(nmol, nres, seed, box_len):
    """nmol copies of CT-[CT-C(=O)-N-CT]xnres-CT backbones, random-walk geometry."""
    rng = np.random.default_rng(seed)
    types_, bonds, coords, charges = [], [], [], []
    placed = np.zeros((0, 3))  # atoms of the molecules already built

    def clear_of_others(c):
        return len(placed) == 0 or np.min(np.linalg.norm(placed - c, axis=1)) > 3.2

    q = {"C": 0.5973, "CT": 0.0337, "N": -0.4157, "O": -0.5679}
    for _ in range(nmol):
        base = len(types_)
        seq = ["CT"]
        bb = [base]  # backbone atom indices
        carbonyls = []
        for _r in range(nres):
            for t in ("CT", "C", "N", "CT"):
                seq.append(t)
                bb.append(base + len(seq) - 1)
                if t == "C":
                    seq.append("O")
                    carbonyls.append((base + len(seq) - 2, base + len(seq) - 1))
        seq.append("CT")
        bb.append(base + len(seq) - 1)
        types_ += seq
        for a, b in zip(bb[:-1], bb[1:]):
            bonds.append((a, b))
        bonds += carbonyls
        # geometry: self-avoiding walk for the backbone, O placed off its C
        pos = {}
        while True:
            start = rng.uniform(6.0, box_len - 6.0, 3)
            if clear_of_others(start):
                break
        pos[bb[0]] = start
        prev_dir = rng.normal(size=3)
        prev_dir /= np.linalg.norm(prev_dir)
        for a, b in zip(bb[:-1], bb[1:]):
            for _try in range(1000):
                d = prev_dir * 0.45 + rng.normal(size=3)
                d /= np.linalg.norm(d)
                cand = pos[a] + 1.5 * d
                inside = np.all(cand > 2.0) and np.all(cand < box_len - 2.0)
                if inside and clear_of_others(cand) and all(np.linalg.norm(cand - p) > 2.3 for k, p in pos.items() if k != a):
                    break
            pos[b] = cand
            prev_dir = d
        for c, o in carbonyls:
            for _try in range(1000):
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                cand = pos[c] + 1.23 * d
                if clear_of_others(cand) and all(np.linalg.norm(cand - p) > 2.1 for k, p in pos.items() if k != c):
                    break
            pos[o] = cand
        coords += [pos[base + i] for i in range(len(seq))]
        placed = np.array(coords)
        charges += [q[t] for t in seq]
    n = len(types_)
    bonds = np.array(bonds, dtype=np.int64)
    nb = [[] for _ in range(n)]
    for a, b in bonds:
        nb[a].append(b)
        nb[b].append(a)
    angles = [(a, j, c) for j in range(n) for a in nb[j] for c in nb[j] if a < c]
    dihedrals = []
    for j, k in bonds:
        for i in nb[j]:
            for l in nb[k]:  # noqa: E741
                if i != k and l != j and i != l:
                    dihedrals.append((i, j, k, l))
    # impropers: carbonyl carbon (3 bonded neighbours) as third atom, O last
    impropers = []
    tarr = np.array(types_, dtype=object)
    for c in range(n):
        if tarr[c] == "C" and len(nb[c]) == 3:
            o = [x for x in nb[c] if tarr[x] == "O"][0]
            others = sorted(x for x in nb[c] if x != o)
            impropers.append((others[0], others[1], c, o))
    mass = {"C": 12.01, "CT": 12.01, "N": 14.01, "O": 16.0}
    return {
        "coords": np.array(coords, dtype=np.float32),
        "atomtype": tarr,
        "charge": np.array(charges, dtype=np.float32),
        "masses": np.array([mass[t] for t in types_], dtype=np.float32),
        "bonds": bonds,
        "angles": np.array(angles, dtype=np.int64),
        "dihedrals": np.array(dihedrals, dtype=np.int64),
        "impropers": np.array(impropers, dtype=np.int64),
    }


# ----------------------------------------------------------------------------
def main():
    torch.manual_seed(1)
    wff = os.path.join(REF, "tests/water/water_forcefield.yaml")
    aff = os.path.join(REF, "tests/argon/argon_forcefield.yaml")
    report = []

    # ---- G0: which rounding sequence does torch.norm(dim=1) use in fp32? ------
    v = (torch.randn(200003, 3) * 5).float()
    n = torch.norm(v, dim=1)
    vd = v.double()
    fl = lambda t: t.float().double()  # noqa: E731
    s = fl(vd[:, 2] * vd[:, 2] + fl(vd[:, 1] * vd[:, 1] + fl(vd[:, 0] * vd[:, 0]))).float()
    assert np.array_equal(np.sqrt(s.numpy()), n.numpy()), "torch.norm != sqrt_rn(fma(z,z,fma(y,y,x*x)))"
    report.append("torch.norm(fp32,(P,3),dim=1) == sqrt_rn(fma(z,z,fma(y,y,x*x))) on %s, torch %s, cap %s"
                  % (os.uname().machine, torch.__version__, torch.backends.cpu.get_cpu_capability()))

    # ---- G1: water fixture, 291 atoms (tests/water) ----------------------------
    xyz, box = read_pdb(os.path.join(REF, "tests/water/structure.pdb"))
    top = read_psf(os.path.join(REF, "tests/water/structure.psf"))
    w291 = dict(top, coords=xyz, box=box)
    terms = ["lj", "electrostatics", "bonds", "angles"]
    res = force_case(w291, wff, terms, nrep=2, cutoff=7.3, rfa=True, switch_dist=6.0)
    for k in ("atomtype", "charge", "masses", "bonds", "angles"):
        res["top_" + k] = np.array([str(x) for x in w291[k]]) if k == "atomtype" else w291[k]
    save("water291_rf_switch", res)
    # the fixture's own run recipe (tests/water/water_conf.yaml): cutoff 7.3, no RF, no switch
    save("water291_plain", force_case(w291, wff, terms, nrep=1, store_pairs=False, cutoff=7.3, rfa=False, switch_dist=None))
    print("water291 fp64 energies", dict(zip(res["energy_keys"], res["energies_f64"][0])))

    # ---- G2: argon fixture, 100 atoms, LJ only ---------------------------------
    axyz, abox = read_pdb(os.path.join(REF, "tests/argon/argon_start.pdb"))
    a100 = dict(coords=axyz, box=abox, atomtype=np.array(["AR"] * len(axyz), dtype=object),
                charge=np.zeros(len(axyz), np.float32), masses=np.full(len(axyz), 39.95, np.float32))
    save("argon100_nocut", force_case(a100, aff, ["lj"], cutoff=None, rfa=False, switch_dist=None))
    save("argon100_cut", force_case(a100, aff, ["lj"], cutoff=9.0, rfa=False, switch_dist=7.5))

    # repulsion / repulsionCG terms (forces.py:418-450).  The reference only builds the A/B tables
    # when "lj" is among the terms (forces.py:45-46) -- repulsion alone raises AttributeError there --
    # so the runnable combination is the mix.
    save("argon100_lj_rep_mix", force_case(a100, aff, ["lj", "repulsion", "repulsioncg"], cutoff=12.0, rfa=False, switch_dist=10.0))
    save("argon100_lj_rep_nocut", force_case(a100, aff, ["lj", "repulsioncg", "repulsion"], cutoff=None, rfa=False, switch_dist=None))

    # ---- G3: synthetic water, 999 atoms, equilibrated with the reference --------
    d = testsystems.water_box(333, seed=0)
    # our array-built parameters must equal the reference builder's output
    mol, par, sysm, frc = ref_setup(d, wff, terms, torch.float32, cutoff=9.0, rfa=True, switch_dist=7.5)
    ours = testsystems.water_parameters(d)
    assert torch.equal(ours.mapped_atom_types, par.mapped_atom_types)
    assert torch.equal(ours.charges, par.charges) and torch.equal(ours.masses, par.masses)
    for nm in ("bond_params", "angle_params"):
        for k in ("idx", "map", "params"):
            assert torch.equal(getattr(ours, nm)[k], getattr(par, nm)[k]), (nm, k)
    A, B = ours.get_AB()
    assert torch.equal(A, par.A) and torch.equal(B, par.B)
    report.append("testsystems.water_parameters == reference Parameters(YamlForcefield) (bitwise)")
    torch.manual_seed(1)
    sysm.set_velocities(ref_mb(par.masses, 300.0, 1))
    integ = RefIntegrator(sysm, frc, 1.0, "cpu", gamma=5.0, T=300.0)
    frc.compute(sysm.pos, sysm.box, sysm.forces)
    for _ in range(40):
        ek, ep, T = integ.step(niter=10)
    print("water999 after 400 steps: T=%.1f Epot=%.1f" % (T[0], ep[0]))
    d_eq = dict(d, coords=sysm.pos[0].numpy().copy())
    res = force_case(d_eq, wff, terms, cutoff=9.0, rfa=True, switch_dist=7.5)
    res["vel"] = sysm.vel[0].numpy().copy()
    save("water999_eq", res)

    # ---- G4: chain molecules with dihedrals / impropers / 1-4 -------------------
    allterms = ["bonds", "angles", "dihedrals", "impropers", "1-4", "lj", "electrostatics"]
    for tag, imp, boxlen, fkw in (
        ("chain_amber_vacuum", dict(imp_k=10.5, imp_phase=180.0, imp_per=2), 0.0,
         dict(cutoff=None, rfa=False, switch_dist=None)),
        ("chain_amber_periodic", dict(imp_k=10.5, imp_phase=180.0, imp_per=2), 30.0,
         dict(cutoff=9.0, rfa=True, switch_dist=7.5)),
        ("chain_charmm_periodic", dict(imp_k=45.0, imp_phase=0.0, imp_per=0), 30.0,
         dict(cutoff=9.0, rfa=True, switch_dist=7.5)),
    ):
        ffp = os.path.join(HERE, tag + "_ff.yaml")
        with open(ffp, "w") as fh:
            fh.write(CHAIN_FF % imp)
        d = chain_molecules(nmol=6, nres=3, seed=3, box_len=boxlen if boxlen else 30.0)
        d["box"] = np.array([boxlen] * 3, dtype=np.float32)
        res = force_case(d, ffp, allterms, **fkw)
        save(tag, res)
        print(tag, dict(zip(res["energy_keys"], np.round(res["energies_f64"][0], 4))))

    # ---- G5: integrator trajectories on the water fixture -----------------------
    traj = {}
    for tag, prec in (("f32", torch.float32), ("f64", torch.float64)):
        mol, par, sysm, frc = ref_setup(w291, wff, terms, prec, nrep=2, cutoff=7.3, rfa=True, switch_dist=6.0)
        torch.manual_seed(7)
        v0 = ref_mb(par.masses, 300.0, 2)
        sysm.set_velocities(v0)
        traj["vel0_" + tag] = sysm.vel.numpy().copy()
        frc.compute(sysm.pos, sysm.box, sysm.forces)
        integ = RefIntegrator(sysm, frc, 1.0, "cpu")  # NVE
        for nst in (1, 9):
            ek, ep, T = integ.step(niter=nst)
            k = "1" if nst == 1 else "10"
            traj[f"nve_pos{k}_{tag}"] = sysm.pos.numpy().copy()
            traj[f"nve_vel{k}_{tag}"] = sysm.vel.numpy().copy()
            traj[f"nve_ekin{k}_{tag}"] = ek
            traj[f"nve_epot{k}_{tag}"] = np.array(ep)
        # Langevin: record the N(0,1) draws the reference consumed so they can be injected
        mol, par, sysm, frc = ref_setup(w291, wff, terms, prec, nrep=2, cutoff=7.3, rfa=True, switch_dist=6.0)
        sysm.set_velocities(v0)
        frc.compute(sysm.pos, sysm.box, sysm.forces)
        integ = RefIntegrator(sysm, frc, 1.0, "cpu", gamma=0.1, T=300.0)
        torch.manual_seed(11)
        g = torch.random.get_rng_state()
        ek, ep, T = integ.step(niter=4)
        torch.random.set_rng_state(g)
        traj["lan_noise_" + tag] = torch.stack([torch.randn_like(sysm.vel) for _ in range(4)]).numpy()
        traj[f"lan_pos4_{tag}"] = sysm.pos.numpy().copy()
        traj[f"lan_vel4_{tag}"] = sysm.vel.numpy().copy()
        traj[f"lan_ekin4_{tag}"] = ek
        traj[f"lan_T4_{tag}"] = T
        traj[f"lan_epot4_{tag}"] = np.array(ep)
        # the oracle integrator must reproduce both runs bit for bit
        for mode in ("nve", "lan"):
            mol, par, s2, f2 = ref_setup(w291, wff, terms, prec, nrep=2, cutoff=7.3, rfa=True, switch_dist=6.0)
            s2.set_velocities(v0)
            of = refmd.OracleForces(par, terms, cutoff=7.3, rfa=True, switch_dist=6.0)
            fn = lambda p, b, f: [sum(e.values()) for e in of.compute(p, b, f)]  # noqa: E731
            fn(s2.pos, s2.box, s2.forces)
            oi = refmd.OracleIntegrator(s2.pos, s2.vel, s2.box, s2.forces, par.masses.to(prec), fn, 1.0,
                                        gamma_ps=0.1 if mode == "lan" else None, T=300.0 if mode == "lan" else None)
            if mode == "nve":
                oi.step(1)
                oi.step(9)
                assert np.array_equal(s2.pos.numpy(), traj["nve_pos10_" + tag])
                assert np.array_equal(s2.vel.numpy(), traj["nve_vel10_" + tag])
            else:
                oi.step(4, noise=torch.from_numpy(traj["lan_noise_" + tag]))
                assert np.array_equal(s2.pos.numpy(), traj["lan_pos4_" + tag])
                assert np.array_equal(s2.vel.numpy(), traj["lan_vel4_" + tag])
    report.append("oracle integrator == reference Integrator (NVE 10 steps, Langevin 4 steps, fp32+fp64, bitwise)")
    save("water291_traj", traj)

    # ---- G6: adversarial cutoff pairs -------------------------------------------
    # 2 x 600 atoms: pair k sits at r_c * (1 + m*2^-23), m in [-6, 6], random direction,
    # pairs far apart from each other; some straddle the periodic boundary; positions
    # are NOT wrapped (offsets of whole box lengths) as happens in a running simulation.
    rng = np.random.default_rng(5)
    rc = 9.0
    npair = 600
    L = 60.0
    grid = np.stack(np.meshgrid(*[np.arange(10)] * 3, indexing="ij"), -1).reshape(-1, 3)[:npair]
    centers = (grid * 6.0 + 0.5).astype(np.float64)  # pair centres >= 6 A apart... pairs themselves 9 A long
    # spread pair centres on a coarse lattice in a big box so different pairs never come within r_c
    L = 10 * 30.0
    centers = grid * 30.0 + 3.0 + rng.uniform(0, 1, size=(npair, 3))
    dirs = rng.normal(size=(npair, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    m = rng.integers(-6, 7, size=npair)
    r = rc * (1.0 + m * 2.0**-23)
    a = centers - 0.5 * r[:, None] * dirs
    b = centers + 0.5 * r[:, None] * dirs
    shift = rng.integers(-2, 3, size=(npair, 3)) * L  # unwrapped images of atom b
    b_unwrapped = b + shift * (rng.uniform(size=(npair, 1)) < 0.5)
    coords = np.empty((2 * npair, 3), dtype=np.float32)
    coords[0::2] = a
    coords[1::2] = b_unwrapped
    adv = dict(coords=coords, box=np.array([L, L, L], np.float32), atomtype=np.array(["AR"] * (2 * npair), dtype=object),
               charge=np.zeros(2 * npair, np.float32), masses=np.full(2 * npair, 39.95, np.float32))
    res = force_case(adv, aff, ["lj"], cutoff=rc, rfa=False, switch_dist=7.5)
    print("adversarial: %d of %d near-cutoff pairs inside (fp32), %d (fp64)" % (res["npairs_f32"], npair, res["npairs_f64"]))
    save("adversarial_cutoff", res)

    # ---- G7/G8: AMBER prmtop fixtures (configs 3 and 5 of BASELINE.json) ------------------
    # The reference reads these through moleculekit + parmed, which this image lacks; they are
    # read with repo torchmd_b200/amber.py and evaluated with the oracle.  The alanine-dipeptide
    # values are pinned by the reference's STORED known answers (tests/test_torchmd.py:517,605,
    # examples/tutorial.ipynb:105); CODATA revisions of scipy.constants move them by ~2e-4.
    from torchmd_b200 import amber

    allterms = ["bonds", "angles", "dihedrals", "impropers", "1-4", "lj", "electrostatics"]

    def amber_case(name, prmtop, xyz, box, nrep, fkw):
        sysm = amber.AmberSystem(prmtop)
        res = {}
        for tag, prec in (("f32", torch.float32), ("f64", torch.float64)):
            par = amber.amber_parameters(sysm, allterms, precision=prec)
            # fp64 values on the fp32 in/out decisions: the yardstick for an fp32 kernel
            of = refmd.OracleForces(par, allterms, decision_dtype=torch.float32, **fkw)
            pos = torch.tensor(xyz, dtype=prec)[None].repeat(nrep, 1, 1).contiguous()
            bx = torch.zeros(nrep, 3, 3, dtype=prec)
            for k in range(3):
                bx[:, k, k] = float(box[k])
            F = torch.zeros_like(pos)
            E = of.compute(pos, bx, F)
            keys = list(E[0])
            res["energy_keys"] = np.array(keys)
            res[f"energies_{tag}"] = np.array([[e[k] for k in keys] for e in E])
            res[f"forces_{tag}"] = F.numpy().copy()
            p32 = of.neighbour_pairs(pos[0], torch.diagonal(bx[0])).numpy().astype(np.int32)
            res[f"npairs_{tag}"] = np.int64(len(p32))
            res[f"pairs_sha256_{tag}"] = np.array(hashlib.sha256(p32.tobytes()).hexdigest())
            if tag == "f64":  # full-precision parameters; an fp32 run rounds them the same way
                res.update({"par_" + k: v for k, v in pack_params(par).items()})
        res["coords"] = np.asarray(xyz, dtype=np.float32)
        res["box"] = np.asarray(box, dtype=np.float32)
        res["terms"] = np.array(allterms)
        for k, v in fkw.items():
            res["cfg_" + k] = np.array(np.nan if v is None else v)
        res["cfg_nrep"] = np.int64(nrep)
        res["source"] = np.array("oracle/refmd.py on parameters read by torchmd_b200/amber.py (reference needs parmed)")
        save(name, res)
        return res

    ala = os.path.join(REF, "tests/data/prod_alanine_dipeptide_amber")
    axyz = amber.read_bincoor(os.path.join(ala, "input.coor"))
    r1 = amber_case("ala2_nobox_rf", os.path.join(ala, "structure.prmtop"), axyz, np.zeros(3), 2,
                    dict(cutoff=9.0, rfa=True, switch_dist=7.5))
    assert abs(r1["energies_f64"][0].sum() + 1722.3569) < 5e-4, r1["energies_f64"][0].sum()  # test_torchmd.py:517
    r2 = amber_case("ala2_xsc_rf", os.path.join(ala, "structure.prmtop"), axyz, amber.read_xsc(os.path.join(ala, "input.xsc")), 1,
                    dict(cutoff=9.0, rfa=True, switch_dist=7.5))
    e2 = dict(zip([str(k) for k in r2["energy_keys"]], r2["energies_f64"][0]))
    for k, v in dict(electrostatics=-2568.498, lj=359.251, bonds=3.9577, angles=2.8446, dihedrals=10.5799, impropers=1.2417).items():
        assert abs(e2[k] - v) < 2e-3, (k, e2[k], v)  # examples/tutorial.ipynb:105 (fp32 GPU run, 4-7 digits)
    report.append("torchmd_b200/amber.py + oracle reproduce the reference's stored alanine-dipeptide energies: "
                  "-1722.3569 (test_torchmd.py:517) to %.1e, tutorial.ipynb:105 per-term vector to 2e-3"
                  % abs(r1["energies_f64"][0].sum() + 1722.3569))
    thr = os.path.join(REF, "tests/data/thrombin-ligand-amber")
    txyz, _ = read_pdb(os.path.join(thr, "structure.pdb")) if False else (None, None)
    pdb_xyz = []
    for line in open(os.path.join(thr, "structure.pdb")):
        if line.startswith(("ATOM", "HETATM")):
            pdb_xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    amber_case("thrombin_nobox_rf", os.path.join(thr, "structure.prmtop"), np.array(pdb_xyz, dtype=np.float32), np.zeros(3), 2,
               dict(cutoff=7.3, rfa=True, switch_dist=None))

    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as fh:
        fh.write("Generated by tests/golden/make_golden.py from the reference at /root/reference\n")
        fh.write("(torchmd/torchmd @ 09484183e34af78bb69b27a0149aeccf1318dbc7), torch %s, numpy %s.\n" % (torch.__version__, np.__version__))
        fh.write("Checks passed while generating:\n")
        for line in report:
            fh.write("  * " + line + "\n")
        fh.write("  * oracle/refmd.py OracleForces == reference Forces.compute (energies, forces, neighbour pairs; fp32+fp64; bitwise) on every case\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
