"""Golden vectors for the two small AMBER fixtures of the reference's own test matrix that round 1's generator did
not cover: tests/data/benzamidine-amber and tests/data/ligand-amber, configured as tests/test_torchmd.py:363-365
configures a fixture without a box (no cutoff, no switching, plain Coulomb, all terms).

The reference reads prmtop files through moleculekit + parmed, which this image lacks; as for the alanine-dipeptide
and thrombin cases (make_golden.py) the parameters come from repo torchmd_b200/amber.py and the values from
oracle/refmd.py, which is pinned bitwise to the reference's Forces.compute on every case the reference can run here.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_amber_small.py
"""
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

from make_golden import pack_params  # noqa: E402
from oracle import refmd  # noqa: E402
from torchmd_b200 import amber  # noqa: E402

ALLTERMS = ["bonds", "angles", "dihedrals", "impropers", "1-4", "lj", "electrostatics"]


def pdb_coords(path):
    xyz = []
    for line in open(path):
        if line.startswith(("ATOM", "HETATM")):
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.array(xyz, dtype=np.float32)


def case(name, folder, nrep=2):
    sysm = amber.AmberSystem(os.path.join(folder, "structure.prmtop"))
    xyz = pdb_coords(os.path.join(folder, "structure.pdb"))
    fkw = dict(cutoff=None, rfa=False, switch_dist=None)
    res = {}
    # replica 1 is the same molecule slightly deformed; the coordinates are float32 values in both precisions
    start = torch.tensor(xyz, dtype=torch.float32)[None].repeat(nrep, 1, 1).contiguous()
    if nrep > 1:
        g = torch.Generator().manual_seed(7)
        start[1] += (0.02 * torch.randn(start[1].shape, generator=g, dtype=torch.float64)).float()
    for tag, prec in (("f32", torch.float32), ("f64", torch.float64)):
        par = amber.amber_parameters(sysm, ALLTERMS, precision=prec)
        of = refmd.OracleForces(par, ALLTERMS, decision_dtype=torch.float32, **fkw)
        pos = start.to(prec)
        bx = torch.zeros(nrep, 3, 3, dtype=prec)
        F = torch.zeros_like(pos)
        E = of.compute(pos, bx, F)
        keys = list(E[0])
        res["energy_keys"] = np.array(keys)
        res[f"energies_{tag}"] = np.array([[e[k] for k in keys] for e in E])
        res[f"forces_{tag}"] = F.numpy().copy()
        p32 = of.neighbour_pairs(pos[0], torch.diagonal(bx[0])).numpy().astype(np.int32)
        res[f"npairs_{tag}"] = np.int64(len(p32))
        res[f"pairs_sha256_{tag}"] = np.array(hashlib.sha256(p32.tobytes()).hexdigest())
        if tag == "f32":
            res["pairs_f32"] = p32
            res["coords_replicas"] = start.numpy().copy()
        if tag == "f64":
            res.update({"par_" + k: v for k, v in pack_params(par).items()})
    res["coords"] = xyz
    res["box"] = np.zeros(3, dtype=np.float32)
    res["terms"] = np.array(ALLTERMS)
    for k, v in fkw.items():
        res["cfg_" + k] = np.array(np.nan if v is None else v)
    res["cfg_nrep"] = np.int64(nrep)
    res["source"] = np.array("oracle/refmd.py on parameters read by torchmd_b200/amber.py (reference needs parmed)")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
    print(name, "atoms", len(xyz), "pairs", int(res["npairs_f32"]), "E[0]", dict(zip(keys, np.round(res["energies_f64"][0], 4))),
          "max|F|", float(np.abs(res["forces_f64"]).max()))


if __name__ == "__main__":
    case("benzamidine_amber_nocut", os.path.join(REF, "tests/data/benzamidine-amber"))
    case("ligand_amber_nocut", os.path.join(REF, "tests/data/ligand-amber"))
