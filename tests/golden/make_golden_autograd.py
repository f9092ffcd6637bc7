"""Golden forces of the reference's AUTOGRAD path (Forces.compute(explicit_forces=False),
forces.py:328-336) FROM THE UNMODIFIED REFERENCE, for the tests/water fixture state stored
in water291_rf_switch.npz (LJ with switch + reaction field + bonds + angles).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_autograd.py

Build container only.  Asserts that oracle/refmd.py with true_gradient=True reproduces the
reference's autograd forces (to autograd's own rounding) and its energies (bit for bit).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from torchmd.forcefields.forcefield import ForceField  # noqa: E402  (reference)
from torchmd.forces import Forces as RefForces  # noqa: E402
from torchmd.parameters import Parameters as RefParameters  # noqa: E402

from conftest import golden_cfg, golden_system_tensors, params_from_golden  # noqa: E402
from oracle import refmd  # noqa: E402
from torchmd_b200 import charmm  # noqa: E402


def main():
    g = dict(np.load(os.path.join(HERE, "water291_rf_switch.npz")))
    cfg = golden_cfg(g)
    terms = [str(t) for t in g["terms"]]
    d = "/root/reference/tests/water"
    mol = charmm.load_molecule(os.path.join(d, "structure.psf"), os.path.join(d, "structure.pdb"))
    mol.element = None
    out = {}
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        ff = ForceField.create(mol, os.path.join(d, "water_forcefield.yaml"))
        par = RefParameters(ff, mol, terms, precision=dtype, device="cpu")
        frc = RefForces(par, terms=terms, cutoff=cfg["cutoff"], rfa=cfg["rfa"], switch_dist=cfg["switch_dist"])
        pos, box = golden_system_tensors(g, dtype)
        F_auto = torch.zeros_like(pos)
        E = frc.compute(pos.detach().requires_grad_(True), box, F_auto, returnDetails=True, explicit_forces=False)
        F_expl = torch.zeros_like(pos)
        frc.compute(pos, box, F_expl, returnDetails=True)
        assert np.array_equal(F_expl.numpy(), g["forces_" + tag]), "state differs from the stored fixture"
        of = refmd.OracleForces(params_from_golden(g, precision=dtype), terms, true_gradient=True, **cfg)
        F_or = torch.zeros_like(pos)
        E_or = of.compute(pos, box, F_or)
        dev = (F_or - F_auto).abs().max().item()
        quirk = (F_expl - F_auto).abs().max().item()
        print(f"{tag}: reference autograd vs explicit forces differ by {quirk:.3e}; oracle(true_gradient) vs reference autograd {dev:.3e}")
        assert dev < (1e-4 if dtype == torch.float32 else 1e-11)  # fp32: autograd rounds differently from the explicit formulas
        for r in range(len(E)):
            for k in terms:
                assert float(E_or[r][k]) == E[r][k], (k, float(E_or[r][k]), E[r][k])
        out["forces_autograd_" + tag] = F_auto.numpy()
        out["explicit_minus_autograd_" + tag] = np.float64(quirk)
    path = os.path.join(HERE, "water291_autograd.npz")
    np.savez_compressed(path, **out)
    print(f"wrote water291_autograd.npz {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
