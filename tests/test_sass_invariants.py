"""Static properties of the compiled kernels (cuobjdump on the built library, no GPU): register
budgets that the occupancy tuning relies on, no local-memory traffic inside the pair loops, and
the packed fp32x2 instructions where they are supposed to be."""
import os
import shutil
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "scripts"))

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="CUDA toolkit (cuobjdump) not on PATH")


@pytest.fixture(scope="module")
def sass():
    import sass_budget

    return sass_budget.functions(), sass_budget.resources(), sass_budget


def find(table, needle):
    names = [n for n in table if needle in n]
    assert names, f"kernel {needle} not in the library"
    return names[0]


def test_default_pair_kernel_register_budget_and_clean_loop(sass):
    table, res, sb = sass
    name = find(table, "k_pairILb0ELb1ELb1ELi1E")  # <ENERGY=0, PERIODIC, SAFE, MODE=1>: the production instantiation
    regs, stack, _ = res[name]
    assert regs <= 40, f"{regs} registers: 6 CTAs of 256 threads per SM need <= 40"
    lo, hi = sb.main_loop(table[name])
    body = [t for _, t in table[name][lo : hi + 1]]
    assert not any("STL" in t or "LDL" in t for t in body), "local-memory traffic inside the neighbour loop"
    assert sum("LDG.E.128" in t for t in body) == 2 and sum(t.startswith("MUFU") for t in body) == 4


def test_fixed_point_kernels_have_clean_loops(sass):
    table, res, sb = sass
    for needle, max_regs in (("k_pair_fxILb0ELi1ELb1E", 40), ("k_pair_fx2ILb0E", 64)):
        name = find(table, needle)
        regs, _, _ = res[name]
        assert regs <= max_regs, (needle, regs)
        lo, hi = sb.main_loop(table[name])
        body = [t for _, t in table[name][lo : hi + 1]]
        assert not any("STL" in t or "LDL" in t for t in body), needle
        assert len(body) / 2 < 85, (needle, len(body))  # instructions per list entry: 79 / 59 when written


def test_packed_kernel_uses_packed_instructions(sass):
    table, res, sb = sass
    name = find(table, "k_pair_fx2ILb0E")
    lo, hi = sb.main_loop(table[name])
    body = [t for _, t in table[name][lo : hi + 1]]
    packed = sum(any(p in t for p in ("FFMA2", "FMUL2", "FADD2")) for t in body)
    assert packed >= 38, f"only {packed} packed fp32x2 instructions in the loop: the compiler scalarised the arithmetic"
    scalar_fp = sum(t.split()[0] in ("FFMA", "FMUL", "FADD") for t in body)
    assert scalar_fp <= 4, f"{scalar_fp} scalar FP instructions left in the packed loop"
