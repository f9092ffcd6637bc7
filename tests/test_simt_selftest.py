"""The SIMT interpreter of tests/simt tested on its own: the scheduling orders (SIMT_SCHEDULE) exist to expose
kernels that only work because thread 3 happens to run before thread 5, so they must (a) really change the
order and (b) make a toy kernel with a missing barrier misbehave while leaving the correct one alone."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT_DIR = os.path.join(ROOT, "tests", "simt")

PROBE = """
import ctypes as C, sys
import numpy as np
h = C.CDLL(sys.argv[1])
order = np.zeros(64, np.int32)
h.simt_selftest_order(order.ctypes.data_as(C.c_void_p), 64)
print(h.simt_selftest_neighbours(1, 8), h.simt_selftest_neighbours(0, 8), " ".join(map(str, order)))
"""


@pytest.fixture(scope="module")
def selftest_lib():
    out = os.path.join(SIMT_DIR, "libsimt_selftest.so")
    srcs = [os.path.join(SIMT_DIR, f) for f in ("selftest.cpp", "simt.h", os.path.join("stub", "cuda_runtime.h"))]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-U_FORTIFY_SOURCE", "-I", os.path.join(SIMT_DIR, "stub"), "-I", SIMT_DIR,
                        "-o", out, os.path.join(SIMT_DIR, "selftest.cpp")], check=True, cwd=ROOT)
    return out


def probe(lib, schedule):
    env = dict(os.environ)
    env.pop("SIMT_SCHEDULE", None)
    if schedule:
        env["SIMT_SCHEDULE"] = schedule
    r = subprocess.run([sys.executable, "-c", PROBE, lib], env=env, capture_output=True, text=True, check=True)
    f = r.stdout.split()
    return int(f[0]), int(f[1]), np.array(f[2:], int)


def test_schedules_change_the_order_and_expose_a_missing_barrier(selftest_lib):
    total = 8 * 64
    synced, racy, order = probe(selftest_lib, None)
    assert synced == total and list(order) == list(range(64))
    assert racy == 8  # forward: only the last thread of a block finds its neighbour's value already written
    synced, racy, order = probe(selftest_lib, "reverse")
    assert synced == total and list(order) == list(range(63, -1, -1))
    assert racy == total - 8  # reverse: everyone but the first thread to run
    seen = set()
    for seed in (1, 2, 3):
        synced, racy, order = probe(selftest_lib, f"random:{seed}")
        assert synced == total
        assert sorted(order) == list(range(64)) and list(order) != list(range(64))
        assert 0 < racy < total
        seen.add(tuple(order))
    assert len(seen) == 3


@pytest.mark.parametrize("schedule", ["random:11"])
def test_kernel_results_do_not_depend_on_the_thread_order(schedule):
    """The quick half of tests/test_simt_kernels.py again under another thread and block order (the whole file
    passes under SIMT_SCHEDULE=reverse and random orders too; that takes four minutes per order and is run by hand)."""
    pick = "chain or argon or wrap_kernel or tiny or owned or exclusions or overlap or replicas or in_one_process"
    env = dict(os.environ, SIMT_SCHEDULE=schedule)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_simt_kernels.py"), "-q", "-x", "-k", pick,
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
