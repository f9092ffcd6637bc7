"""Minimal CHARMM/X-PLOR PSF and PDB readers -- just what the reference takes from a
moleculekit ``Molecule`` for these formats (parameters.py:25,110-133; npzmol.py:11-39):
atom types, charges, masses, bonds, angles, dihedrals, impropers, coordinates and box.
The reference needs moleculekit for this; the hot path does not.
"""
import types

import numpy as np

_SECTIONS = {"NATOM": 0, "NBOND": 2, "NTHETA": 3, "NPHI": 4, "NIMPHI": 4}


def read_psf(path):
    """dict with atomtype (object array), charge, masses (fp32), bonds (M,2), angles (K,3),
    dihedrals (D,4), impropers (I,4) as zero-based int64 arrays, plus name/resname/resid/segid."""
    with open(path) as fh:
        lines = fh.read().split("\n")
    if not lines or not lines[0].startswith("PSF"):
        raise ValueError(f"{path}: not a PSF file")
    sec = {}
    i = 0
    while i < len(lines):
        ln = lines[i]
        if "!N" in ln:
            key = ln.split("!")[1].split(":")[0].split()[0]
            count = int(ln.split()[0])
            body, i = [], i + 1
            while i < len(lines) and lines[i].strip():
                body.append(lines[i])
                i += 1
            sec[key] = (count, body)
        else:
            i += 1
    if "NATOM" not in sec:
        raise ValueError(f"{path}: no !NATOM section")
    natom, rows = sec["NATOM"]
    atoms = [r.split() for r in rows]
    if len(atoms) != natom:
        raise ValueError(f"{path}: !NATOM says {natom}, found {len(atoms)} atom records")

    def index_section(key, width):
        if key not in sec:
            return np.zeros((0, width), dtype=np.int64)
        count, body = sec[key]
        flat = np.array(" ".join(body).split(), dtype=np.int64) - 1
        if len(flat) != count * width:
            raise ValueError(f"{path}: !{key} says {count} entries, found {len(flat) / width:g}")
        if len(flat) and (flat.min() < 0 or flat.max() >= natom):
            raise ValueError(f"{path}: !{key} refers to an atom outside 1..{natom}")
        return flat.reshape(-1, width)

    return {
        "segid": np.array([a[1] for a in atoms], dtype=object),
        "resid": np.array([int(a[2]) for a in atoms], dtype=np.int64),
        "resname": np.array([a[3] for a in atoms], dtype=object),
        "name": np.array([a[4] for a in atoms], dtype=object),
        "atomtype": np.array([a[5] for a in atoms], dtype=object),
        "charge": np.array([float(a[6]) for a in atoms], dtype=np.float32),
        "masses": np.array([float(a[7]) for a in atoms], dtype=np.float32),
        "bonds": index_section("NBOND", 2),
        "angles": index_section("NTHETA", 3),
        "dihedrals": index_section("NPHI", 4),
        "impropers": index_section("NIMPHI", 4),
    }


def read_pdb(path):
    """(coords (N,3) fp32 of the first model, box (3,) fp32 from CRYST1 or zeros)."""
    xyz, box = [], np.zeros(3, dtype=np.float32)
    with open(path) as fh:
        for line in fh:
            if line.startswith("CRYST1"):
                box = np.array([float(line[6:15]), float(line[15:24]), float(line[24:33])], dtype=np.float32)
            elif line.startswith(("ATOM", "HETATM")):
                xyz.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
            elif line.startswith("ENDMDL"):
                break
    return np.array(xyz, dtype=np.float32).reshape(-1, 3), box


def load_molecule(psf_path, pdb_path=None):
    """The duck-typed ``mol`` the reference's Parameters / System set-up reads
    (SURVEY.md section 8c): numAtoms, atomtype, charge, masses, bonds, angles, dihedrals,
    impropers, coords (N,3,1), box (3,1)."""
    top = read_psf(psf_path)
    mol = types.SimpleNamespace(**top)
    mol.numAtoms = len(top["atomtype"])
    mol.coords = np.zeros((mol.numAtoms, 3, 1), dtype=np.float32)
    mol.box = np.zeros((3, 1), dtype=np.float32)
    if pdb_path is not None:
        xyz, box = read_pdb(pdb_path)
        if len(xyz) != mol.numAtoms:
            raise ValueError(f"{pdb_path}: {len(xyz)} atoms, the PSF has {mol.numAtoms}")
        mol.coords = xyz[:, :, None].copy()
        mol.box = box[:, None].copy()
    return mol
