import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native artefacts exist (compiles with nvcc/g++ when stale)."""
    import __graft_entry__ as g

    g.build()


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def params_from_golden(g, precision=torch.float32, device="cpu"):
    """TopologyParameters from the par_* arrays a golden file stores."""
    from torchmd_b200.parameters import TopologyParameters

    def term(name):
        if "par_" + name + "_idx" not in g:
            return None
        return (g["par_" + name + "_idx"], g["par_" + name + "_map"], g["par_" + name + "_params"])

    se = g["par_lj_sigma_eps"]
    return TopologyParameters(
        atom_types=g["par_types"],
        type_sigma=se[:, 0],
        type_epsilon=se[:, 1],
        charges=g["par_charges"],
        masses=g["par_masses"],
        bonds=term("bond"),
        angles=term("angle"),
        dihedrals=term("dihedral"),
        impropers=term("improper"),
        pairs14=term("nonbonded_14"),
        precision=precision,
        device=device,
    )


def golden_cfg(g):
    def opt(k):
        v = float(g["cfg_" + k])
        return None if np.isnan(v) else v

    return dict(cutoff=opt("cutoff"), rfa=bool(g["cfg_rfa"]), switch_dist=opt("switch_dist"))


def golden_system_tensors(g, dtype, device="cpu"):
    nrep = int(g["cfg_nrep"])
    if "coords_replicas" in g:  # fixtures whose replicas are distinct configurations
        pos = torch.tensor(g["coords_replicas"], dtype=dtype, device=device).contiguous()
    else:
        pos = torch.tensor(g["coords"], dtype=dtype, device=device)[None].repeat(nrep, 1, 1).contiguous()
    box = torch.zeros(nrep, 3, 3, dtype=dtype, device=device)
    for k in range(3):
        box[:, k, k] = float(g["box"][k])
    return pos, box
