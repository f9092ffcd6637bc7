"""Wrapper.wrap (reference torchmd/wrapper.py): the oracle restatement and the arithmetic of
the CUDA kernel (csrc/wrap.cuh, compiled for the host by the hostcheck shim) against golden
vectors produced by the unmodified reference (tests/golden/make_golden_wrap.py); the GPU
kernel itself in the gpu-marked test below.

Tolerance: bit-exact.  (The image a group is moved by depends on floor(com / box); a
different summation order could only change it for a centre within an ulp of a box face --
none of the fixtures has one, and the test would say so.)
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import refmd

DEV = "cuda:0"  # (tests/test_mirrors_on_interpreter.py runs the gpu tests on the host interpreter with "cpu")
CASES = ["water", "mixed", "nobonds", "zerobox"]


def case(g, name):
    bonds = g[name + "_bonds"]
    return int(g[name + "_natoms"]), (bonds if len(bonds) else None), g[name + "_pos"], g[name + "_box"], g[name + "_after"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_wrap_matches_reference_golden(name):
    g = load_golden("wrap_cases")
    natoms, bonds, pos, box, after = case(g, name)
    groups, single = refmd.molecule_groups(natoms, bonds)
    assert len(groups) == int(g[name + "_ngroups"]) and len(single) == int(g[name + "_nsingle"])
    p = torch.tensor(pos)
    refmd.wrap_positions(p, torch.tensor(box), groups, single)
    assert np.array_equal(p.numpy(), after)


@pytest.mark.parametrize("name", CASES)
def test_molecule_groups_match_oracle(name):
    from torchmd_b200.wrapper import Wrapper, calculate_molecule_groups

    g = load_golden("wrap_cases")
    natoms, bonds, pos, box, after = case(g, name)
    groups, single = refmd.molecule_groups(natoms, bonds)
    mg, ng = calculate_molecule_groups(natoms, bonds)
    assert [t.tolist() for t in mg] == groups and ng.tolist() == single
    w = Wrapper(natoms, bonds, "cpu")  # construction is host-only; wrap() needs CUDA
    assert [t.tolist() for t in w.groups] == groups and w.nongrouped.tolist() == single
    # CSR handed to the kernel: every atom exactly once
    assert sorted(w._atoms.tolist()) == list(range(natoms)) and w._ptr[-1] == natoms
    with pytest.raises(RuntimeError):
        w.wrap(torch.tensor(pos), torch.tensor(box))  # CPU tensors: no fallback


@pytest.mark.parametrize("name", CASES)
def test_kernel_arithmetic_on_host_matches_reference_golden(name):
    from torchmd_b200.wrapper import Wrapper

    hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    g = load_golden("wrap_cases")
    natoms, bonds, pos, box, after = case(g, name)
    w = Wrapper(natoms, bonds, "cpu")
    p = np.ascontiguousarray(pos, np.float32).copy()
    b = np.ascontiguousarray(box, np.float32)
    hc.hc_wrap(natoms, len(w._ptr) - 1, w._ptr.ctypes.data_as(C.c_void_p), w._atoms.ctypes.data_as(C.c_void_p),
               p.shape[0], p.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(p, after)


def test_wrap_index_branch_leaves_caller_tensor_alone():
    """wrapper.py:17-21 rebinds the local `pos`: with wrapidx the caller's tensor is unchanged."""
    from torchmd_b200.wrapper import Wrapper

    g = load_golden("wrap_cases")
    natoms, bonds, pos, box, after = case(g, "mixed")
    w = Wrapper(natoms, bonds, "cpu")
    p = torch.tensor(pos)
    w.wrap(p, torch.tensor(box), wrapidx=[0, 1, 2])
    assert np.array_equal(p.numpy(), pos)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_wrap_matches_reference_golden(name):
    from torchmd_b200.wrapper import Wrapper

    g = load_golden("wrap_cases")
    natoms, bonds, pos, box, after = case(g, name)
    w = Wrapper(natoms, bonds, DEV)
    p = torch.tensor(pos, device=DEV)
    w.wrap(p, torch.tensor(box, device=DEV))
    assert np.array_equal(p.cpu().numpy(), after)


def random_wrap_case(seed):
    """Random bond graph (chains, branched groups, lone atoms), one to three replicas each in its own box, atoms up to four
    boxes outside; returns (natoms, bonds, pos, box, oracle result)."""
    rng = np.random.default_rng(seed)
    natoms = int(rng.integers(1, 200))
    nb = int(rng.integers(0, natoms))
    bonds = None
    if nb and natoms > 1:
        a = rng.integers(0, natoms - 1, nb)
        b = np.minimum(natoms - 1, a + rng.integers(1, 4, nb))  # short-range partners: groups of a few to a few dozen atoms
        bonds = np.stack([a, b], 1)
    nrep = int(rng.integers(1, 4))
    box = np.zeros((nrep, 3, 3), np.float32)
    for r in range(nrep):
        box[r][np.eye(3, dtype=bool)] = rng.uniform(10.0, 40.0, 3)
    diag = box[:, np.eye(3, dtype=bool)]
    pos = (rng.uniform(-3.0, 4.0, (nrep, natoms, 3)) * diag[:, None, :]).astype(np.float32)
    groups, single = refmd.molecule_groups(natoms, bonds)
    want = torch.tensor(pos.copy())
    refmd.wrap_positions(want, torch.tensor(box), groups, single)
    return natoms, bonds, pos, box, want.numpy()


@pytest.mark.parametrize("seed", range(24))
def test_kernel_arithmetic_on_host_matches_oracle_on_random_topologies(seed):
    """The kernel's arithmetic (hostcheck build of csrc/wrap.cuh) against the oracle on random_wrap_case, bit for bit."""
    from torchmd_b200.wrapper import Wrapper

    hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    natoms, bonds, pos, box, want = random_wrap_case(seed)
    w = Wrapper(natoms, bonds, "cpu")
    p = np.ascontiguousarray(pos).copy()
    hc.hc_wrap(natoms, len(w._ptr) - 1, w._ptr.ctypes.data_as(C.c_void_p), w._atoms.ctypes.data_as(C.c_void_p),
               p.shape[0], p.ctypes.data_as(C.c_void_p), np.ascontiguousarray(box).ctypes.data_as(C.c_void_p))
    assert np.array_equal(p, want), f"max diff {np.abs(p - want).max()}"


@pytest.mark.gpu
def test_gpu_wrap_random_topologies():
    from torchmd_b200.wrapper import Wrapper

    for seed in range(24):
        natoms, bonds, pos, box, want = random_wrap_case(seed)
        p = torch.tensor(pos, device=DEV)
        Wrapper(natoms, bonds, DEV).wrap(p, torch.tensor(box, device=DEV))
        assert np.array_equal(p.cpu().numpy(), want), f"seed {seed}: max diff {np.abs(p.cpu().numpy() - want).max()}"
