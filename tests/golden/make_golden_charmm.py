"""Golden vectors for the reference's ten CHARMM fixtures (tests/data/{1water,2ions,3ions,1dihedral,singledihedral,
4dihedrals,benzamidine,2watersperiodic,sodiumperiodic,waterbox}: PSF + PDB + CHARMM .prm [+ .xtc]), configured as
tests/test_torchmd.py:363-369 configures them: all terms; without a box no cutoff, no switching, plain Coulomb; with a
box cutoff = min(box)/2 - 0.01, switch_dist 6, reaction field.

The values come from the UNMODIFIED reference: its Parameters (torchmd/parameters.py) builds the tables, its
Forces.compute evaluates them, in fp32 and fp64.  What the reference cannot do here is read the files (it needs
moleculekit + parmed): the molecule comes from repo torchmd_b200/charmm.py (load_molecule, read_xtc_first_frame) and the
force-field object the reference queries is repo CharmmPrmForceField, a restatement of its parmed adapter's look-ups.

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden_charmm.py
"""
import glob
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

from make_golden import pack_params  # noqa: E402
from oracle import refmd  # noqa: E402
from torchmd.forces import Forces as RefForces  # noqa: E402
from torchmd.parameters import Parameters as RefParameters  # noqa: E402
from torchmd_b200 import charmm  # noqa: E402

ALLTERMS = ["bonds", "angles", "dihedrals", "impropers", "1-4", "electrostatics", "lj"]
CASES = ["1water", "2ions", "3ions", "1dihedral", "singledihedral", "4dihedrals", "benzamidine"]
# box from the .xtc header as the reference's test reads it (tests/test_torchmd.py:349-351); periodic fixtures are configured
# as :366-369 does: cutoff = min(box)/2 - 0.01, switch_dist 6, reaction field.  waterbox's trajectory has a zero box (no
# cutoff); its 293-atom frames are compressed, so the coordinates of its structure.pdb are used.
XTC_CASES = ["2watersperiodic", "sodiumperiodic", "waterbox"]


def case(name, nrep=2, xtc=False):
    folder = os.path.join(REF, "tests/data", name)
    psf, pdb, prm = (glob.glob(os.path.join(folder, e))[0] for e in ("*.psf", "*.pdb", "*.prm"))
    mol = charmm.load_molecule(psf, pdb)
    ff = charmm.CharmmPrmForceField(prm, mol)
    xyz = np.asarray(mol.coords, dtype=np.float32).reshape(-1, 3)
    boxv = np.zeros(3, dtype=np.float32)
    if xtc:
        frame, boxv = charmm.read_xtc_first_frame(glob.glob(os.path.join(folder, "*.xtc"))[0])
        if frame is not None:
            assert np.abs(frame - xyz).max() < 2e-3, name  # the trajectory's first frame is the PDB structure
    cfg = dict(cutoff=None, rfa=False, switch_dist=None)
    if np.any(boxv != 0):
        cfg = dict(cutoff=float(np.min(boxv)) / 2 - 0.01, rfa=True, switch_dist=6.0)
    start = torch.tensor(xyz)[None].repeat(nrep, 1, 1).contiguous()
    g = torch.Generator().manual_seed(11)
    start[1] += (0.02 * torch.randn(start[1].shape, generator=g, dtype=torch.float64)).float()
    terms = list(ALLTERMS)
    res = {}
    for tag, prec in (("f32", torch.float32), ("f64", torch.float64)):
        par = RefParameters(ff, mol, terms, precision=prec, device="cpu")
        use = [t for t in terms if not (t == "impropers" and par.improper_params is None)]
        f = RefForces(par, terms=use, **cfg)
        pos = start.to(prec)
        box = torch.zeros(nrep, 3, 3, dtype=prec)
        for k in range(3):
            box[:, k, k] = float(boxv[k])
        F = torch.zeros_like(pos)
        E = f.compute(pos, box, F, returnDetails=True)
        keys = [k for k in E[0] if k != "external"]
        res["energy_keys"] = np.array(keys)
        res[f"energies_{tag}"] = np.array([[e[k] for k in keys] for e in E])
        res[f"forces_{tag}"] = F.numpy().copy()
        # the oracle must agree bit for bit with the reference on the same tables
        of = refmd.OracleForces(par, use, **cfg)
        Fo = torch.zeros_like(pos)
        Eo = of.compute(pos, box, Fo)
        assert torch.equal(F, Fo), name
        assert all(Eo[r][k] == E[r][k] for r in range(nrep) for k in keys), name
        p32 = of.neighbour_pairs(pos[0], torch.diagonal(box[0])).numpy().astype(np.int32)
        res[f"npairs_{tag}"] = np.int64(len(p32))
        res[f"pairs_sha256_{tag}"] = np.array(hashlib.sha256(p32.tobytes()).hexdigest())
        if tag == "f32":
            res["pairs_f32"] = p32
        if tag == "f64":
            res.update({"par_" + k: v for k, v in pack_params(par).items()})
    res["coords"] = xyz
    res["coords_replicas"] = start.numpy().copy()
    res["box"] = boxv
    res["terms"] = np.array(use)
    for k, v in cfg.items():
        res["cfg_" + k] = np.array(np.nan if v is None else v)
    res["cfg_nrep"] = np.int64(nrep)
    res["source"] = np.array("reference Parameters + Forces.compute on a molecule / force field read by torchmd_b200/charmm.py")
    np.savez_compressed(os.path.join(HERE, "charmm_" + name + ".npz"), **res)
    print("charmm_" + name, "atoms", len(xyz), "terms", use, "E[0]", dict(zip(keys, np.round(res["energies_f64"][0], 4))))


if __name__ == "__main__":
    for c in CASES:
        case(c)
    for c in XTC_CASES:
        case(c, xtc=True)
