"""Scalar device functions of csrc/physics.cuh, compiled for the host by the
tests/hostcheck shim, against the oracle.  Catches formula and rounding-order
mistakes without a GPU; the GPU parity tests repeat the comparison end to end."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_cfg, load_golden
from oracle import refmd

F32 = np.float32


@pytest.fixture(scope="module")
def hc():
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    lib.hc_squared_threshold.restype = C.c_float
    lib.hc_squared_threshold.argtypes = [C.c_float]
    return lib


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def decide(hc, pi, pj, box, cutoff):
    n = len(pi)
    w = np.zeros((n, 3), F32)
    s = np.zeros(n, F32)
    inside = np.zeros(n, np.int32)
    periodic = int(np.all(box != 0))
    smax = hc.hc_squared_threshold(F32(cutoff)) if cutoff is not None else F32(np.inf)
    bx = np.where(box == 0, 1.0, box).astype(F32)
    hc.hc_decide(n, p(np.ascontiguousarray(pi, F32)), p(np.ascontiguousarray(pj, F32)), p(bx), periodic,
                 C.c_float(smax), p(w), p(s), p(inside))
    return w, s, inside.astype(bool)


def test_squared_threshold_is_exact(hc):
    for rc in (7.3, 9.0, 12.0, 0.37, 33.333):
        rc32 = F32(rc)
        t = F32(hc.hc_squared_threshold(rc32))
        assert np.sqrt(t) <= rc32
        assert np.sqrt(np.nextafter(t, F32(np.inf))) > rc32


def test_decision_matches_reference_on_adversarial_pairs(hc):
    g = load_golden("adversarial_cutoff")
    xyz = g["coords"]
    box = g["box"]
    pi, pj = xyz[0::2], xyz[1::2]
    w, s, inside = decide(hc, pi, pj, box, 9.0)
    ref = np.zeros(len(pi), bool)
    ref[g["pairs_f32"][:, 0] // 2] = True  # pair k is atoms (2k, 2k+1)
    assert np.array_equal(inside, ref)
    # and the wrapped vector / distance are bit-identical to the oracle's
    dist, _, vec = refmd.pair_geometry(torch.tensor(xyz), torch.tensor(np.stack([np.arange(0, len(xyz), 2), np.arange(1, len(xyz), 2)], 1)), torch.tensor(box))
    assert np.array_equal(vec.numpy(), w)
    assert np.array_equal(dist.numpy(), np.sqrt(s))


def test_decision_random_pairs_small_box(hc):
    """Small box (cutoff close to L/2): the slow exact-division branch is exercised."""
    rng = np.random.default_rng(0)
    box = np.array([16.919, 16.633, 16.639], F32)
    n = 200000
    pi = rng.uniform(-40, 60, size=(n, 3)).astype(F32)
    pj = rng.uniform(-40, 60, size=(n, 3)).astype(F32)
    w, s, inside = decide(hc, pi, pj, box, 7.3)
    d = torch.tensor(pi) - torch.tensor(pj)
    b = torch.tensor(box)[None]
    wref = d - b * torch.round(d / b)
    dist = torch.norm(wref, dim=1)
    assert np.array_equal(wref.numpy(), w)
    assert np.array_equal((dist <= 7.3).numpy(), inside)


def test_decision_half_box_ties(hc):
    """d/L exactly on .5: round-half-even must be reproduced."""
    box = np.array([16.0, 32.0, 8.0], F32)
    ks = np.arange(-7, 8).astype(F32)
    pi = np.stack([ks * 8.0, ks * 16.0, ks * 4.0], 1).astype(F32)
    pj = np.zeros_like(pi)
    w, s, _ = decide(hc, pi, pj, box, 7.3)
    d = torch.tensor(pi)
    b = torch.tensor(box)[None]
    assert np.array_equal((d - b * torch.round(d / b)).numpy(), w)


@pytest.mark.parametrize("rfa,switch", [(True, 7.5), (False, None), (True, None), (False, 6.0)])
def test_pair_terms_match_oracle(hc, rfa, switch):
    rng = np.random.default_rng(1)
    n = 20000
    cutoff = 9.0
    r = rng.uniform(1.6, cutoff, n)
    s = (r * r).astype(F32)
    qi, qj = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    A = rng.uniform(1e4, 6e5, n).astype(F32)
    B = rng.uniform(10, 600, n).astype(F32)
    qq = (refmd.COULOMB * qi * qj).astype(F32)
    terms = (1 << 5) | (1 << 6) | (1 << 7) | (1 << 8)
    eps = 78.5
    denom = 2 * eps + 1
    krf = (1 / cutoff**3) * (eps - 1) / denom
    crf = (1 / cutoff) * (3 * eps) / denom
    out = [np.zeros(n, F32) for _ in range(5)]
    hc.hc_pair_terms(n, p(s), p(qq), p(A), p(B), terms, 1, C.c_float(cutoff), int(switch is not None),
                     C.c_float(switch or 0.0), int(rfa), C.c_float(krf), C.c_float(crf), *[p(o) for o in out])
    e_el, e_lj, e_rep, e_cg, dedr = out
    dist = torch.tensor(np.sqrt(s.astype(np.float64)))
    A64, B64 = torch.tensor(A.astype(np.float64)), torch.tensor(B.astype(np.float64))
    lj_e, lj_f = refmd.lj_pair(dist, A64, B64, 1, switch, cutoff)
    el_e, el_f = refmd.coulomb_pair(dist, torch.tensor(qq.astype(np.float64) / refmd.COULOMB), torch.ones(n, dtype=torch.float64), 1, cutoff, rfa, eps)
    rp_e, rp_f = refmd.repulsion_pair(dist, A64)
    cg_e, cg_f = refmd.repulsion_cg_pair(dist, B64)
    for mine, ref in ((e_lj, lj_e), (e_el, el_e), (e_rep, rp_e), (e_cg, cg_e)):
        ref = ref.numpy()
        assert np.abs(mine - ref).max() <= 3e-6 * max(1.0, np.abs(ref).max())
    tot = (lj_f + el_f + rp_f + cg_f).numpy()
    assert np.abs(dedr - tot).max() <= 3e-6 * max(1.0, np.abs(tot).max())


def test_bond_angle_torsion_match_oracle(hc):
    rng = np.random.default_rng(2)
    n = 5000
    r = rng.uniform(0.8, 2.0, n).astype(F32)
    k0 = rng.uniform(100, 600, n).astype(F32)
    r0 = rng.uniform(0.9, 1.6, n).astype(F32)
    e, f = np.zeros(n, F32), np.zeros(n, F32)
    hc.hc_bond(n, p(r), p(k0), p(r0), p(e), p(f))
    eo, fo = refmd.harmonic_bond(torch.tensor(r.astype(np.float64)), torch.tensor(np.stack([k0, r0], 1).astype(np.float64)))
    assert np.allclose(e, eo.numpy(), rtol=2e-6, atol=1e-5) and np.allclose(f, fo.numpy(), rtol=2e-6, atol=1e-4)

    r21 = rng.normal(size=(n, 3)).astype(F32)
    r23 = rng.normal(size=(n, 3)).astype(F32)
    th0 = rng.uniform(1.5, 2.2, n).astype(F32)
    ka = rng.uniform(30, 80, n).astype(F32)
    e = np.zeros(n, F32)
    fa = np.zeros((n, 9), F32)
    hc.hc_angle(n, p(r21), p(r23), p(ka), p(th0), p(e), p(fa))
    eo, fs = refmd.harmonic_angle(torch.tensor(r21.astype(np.float64)), torch.tensor(r23.astype(np.float64)), torch.tensor(np.stack([ka, th0], 1).astype(np.float64)))
    fo = torch.cat(fs, dim=1).numpy()
    # near-collinear or very short vectors amplify fp32 rounding (1/sin, 1/|r|): test the rest
    n21, n23 = np.linalg.norm(r21, axis=1), np.linalg.norm(r23, axis=1)
    sin = np.linalg.norm(np.cross(r21, r23), axis=1) / (n21 * n23)
    ok = (sin > 0.2) & (n21 > 0.5) & (n23 > 0.5)
    assert ok.sum() > 2000
    assert np.abs(e[ok] - eo.numpy()[ok]).max() < 2e-5 * max(1, eo[ok].abs().max().item())
    assert np.abs(fa[ok] - fo[ok]).max() < 2e-5 * max(1, np.abs(fo[ok]).max())

    for amber in (1, 0):
        r12 = rng.normal(size=(n, 3)).astype(F32) * 1.5
        r23 = rng.normal(size=(n, 3)).astype(F32) * 1.5
        r34 = rng.normal(size=(n, 3)).astype(F32) * 1.5
        nterm = rng.integers(1, 4, n)
        tp = np.concatenate([[0], np.cumsum(nterm)]).astype(np.int32)
        nt = int(tp[-1])
        terms = np.stack([rng.uniform(0.05, 3.0, nt), rng.choice([0.0, math.pi, 0.6], nt),
                          rng.integers(1, 5, nt) if amber else np.zeros(nt)], 1).astype(F32)
        e = np.zeros(n, F32)
        ft = np.zeros((n, 12), F32)
        hc.hc_torsion(n, p(r12), p(r23), p(r34), p(tp), p(terms), amber, p(e), p(ft))
        rows = torch.tensor(np.repeat(np.arange(n), nterm))
        eo, fs = refmd.torsion(torch.tensor(r12.astype(np.float64)), torch.tensor(r23.astype(np.float64)),
                               torch.tensor(r34.astype(np.float64)), rows, torch.tensor(terms.astype(np.float64)))
        fo = torch.cat(fs, dim=1).numpy()
        # ill-conditioned geometries (nearly collinear) amplify fp32 rounding: compare where well conditioned
        cA = np.linalg.norm(np.cross(r12, r23), axis=1)
        cB = np.linalg.norm(np.cross(r23, r34), axis=1)
        ok = (cA > 0.3) & (cB > 0.3)
        assert np.abs(e[ok] - eo.numpy()[ok]).max() < 5e-5 * max(1, eo.abs().max().item())
        assert np.abs(ft[ok] - fo[ok]).max() < 2e-4 * max(1, np.abs(fo[ok]).max())


# ---- fixed-point separations of the periodic pair kernel (k_pair_fx) -----------------------
def fx_decide(hc, pi, pj, box, cutoff, rmax, pmax):
    n = len(pi)
    w = np.zeros((n, 3), F32)
    s = np.zeros(n, F32)
    cls = np.zeros(n, np.int32)
    margin = C.c_float(0)
    smax = hc.hc_squared_threshold(F32(cutoff))
    hc.hc_fx_decide(n, p(np.ascontiguousarray(pi, F32)), p(np.ascontiguousarray(pj, F32)), p(box.astype(F32)),
                    C.c_float(smax), C.c_double(rmax), C.c_float(pmax), p(w), p(s), p(cls), C.byref(margin))
    return w, s, cls, margin.value


def test_fx_separation_is_the_exact_minimum_image(hc):
    """Integer subtraction of the fixed-point coordinates == fp64 minimum image of the fp32
    positions, also for atoms that drifted several boxes away and across the boundary."""
    rng = np.random.default_rng(5)
    box = np.array([99.93, 87.41, 120.07], F32)
    n = 200000
    pi = (rng.uniform(0, 1, (n, 3)) * box + rng.integers(-6, 7, (n, 3)) * box).astype(F32)
    d = rng.normal(size=(n, 3))
    d *= (rng.uniform(0.5, 11.0, n) / np.linalg.norm(d, axis=1))[:, None]
    pj = (pi.astype(np.float64) - d + rng.integers(-6, 7, (n, 3)) * box.astype(np.float64)).astype(F32)
    w, s, _, _ = fx_decide(hc, pi, pj, box, 9.0, 11.0, 8 * 120.07)
    L = box.astype(np.float64)
    d64 = pi.astype(np.float64) - pj.astype(np.float64)
    w64 = d64 - L * np.round(d64 / L)
    err = np.abs(w - w64)
    # quantisation L/2^32 per coordinate (x2) + three fp32 roundings of |w|
    bound = 2 * L.max() / 2**32 + 3 * 2.0**-24 * np.abs(w64) + 1e-12
    assert (err <= bound).all(), (err - bound).max()
    assert err.max() < 3e-6


def test_fx_decision_band_is_sound(hc):
    """Outside the band the fixed-point decision IS the reference's; the band is thin."""
    g = load_golden("adversarial_cutoff")
    xyz, box = g["coords"], g["box"]
    pi, pj = xyz[0::2], xyz[1::2]
    ref = np.zeros(len(pi), bool)
    ref[g["pairs_f32"][:, 0] // 2] = True
    pmax = float(np.abs(xyz).max())
    _, _, cls, margin = fx_decide(hc, pi, pj, box, 9.0, 11.0, pmax)
    assert ((cls == 2) | ((cls == 1) == ref)).all()
    assert (cls == 2).any()  # the adversarial pairs sit within ulps of the cutoff: they must land in the band
    assert margin < 0.02

    # random pairs in a production-size box, atoms anywhere within +-4 boxes
    rng = np.random.default_rng(6)
    box = np.array([99.93, 99.93, 99.93], F32)
    n = 400000
    pi = (rng.uniform(-4, 5, (n, 3)) * box).astype(F32)
    d = rng.normal(size=(n, 3))
    d *= (rng.uniform(8.9, 9.1, n) / np.linalg.norm(d, axis=1))[:, None]  # all near the cutoff
    pj = (pi.astype(np.float64) - d + rng.integers(-2, 3, (n, 3)) * box.astype(np.float64)).astype(F32)
    pmax = float(max(np.abs(pi).max(), np.abs(pj).max()))
    _, s, cls, margin = fx_decide(hc, pi, pj, box, 9.0, 11.0, pmax)
    _, sref, inside = decide(hc, pi, pj, box, 9.0)
    assert ((cls == 2) | ((cls == 1) == inside)).all()
    assert np.abs(s - sref).max() <= margin  # the bound the band is built from
    # band half-width in r is margin / (2 rc): a 0.2 A window around the cutoff holds a few % of it
    assert (cls == 2).mean() < 0.05


def _water_pair_setup(name="water999_eq"):
    g = load_golden(name)
    cfg = golden_cfg(g)
    par64 = __import__("conftest").params_from_golden(g, precision=torch.float64)
    terms = ["lj", "electrostatics"]
    of = refmd.OracleForces(par64, terms, decision_dtype=torch.float32, cutoff=cfg["cutoff"], rfa=cfg["rfa"],
                            switch_dist=cfg["switch_dist"])
    return g, cfg, par64, of


@pytest.mark.parametrize("drift", [False, True])
def test_fx_pair_forces_meet_the_tolerance(hc, drift):
    """Host emulation of both pair kernels' value arithmetic on the equilibrated 999-atom water
    box against the fp64 oracle (same fp32 pair set): the fixed-point path must be at least as
    accurate as the float path, also when molecules have drifted out of the primary box."""
    g, cfg, par64, of = _water_pair_setup()
    pos = g["coords"].astype(F32).copy()
    box = g["box"].astype(F32)
    if drift:
        rng = np.random.default_rng(7)
        shift = rng.integers(-3, 4, (len(pos) // 3, 3)).repeat(3, axis=0)  # whole molecules
        pos = (pos.astype(np.float64) + shift * box.astype(np.float64)).astype(F32)
    pos_t = torch.tensor(pos)[None]
    box_t = torch.diag(torch.tensor(box))[None]
    f64 = torch.zeros(1, len(pos), 3, dtype=torch.float64)
    of.compute(pos_t.double(), box_t.double(), f64)
    pairs = of.neighbour_pairs(pos_t[0], torch.tensor(box)).numpy().astype(np.int32)
    types = g["par_types"].astype(np.int32)
    nt = int(types.max()) + 1
    AB = np.stack([par64.A.numpy(), par64.B.numpy()], -1).astype(F32).reshape(nt, nt, 2)
    qs = (g["par_charges"] * math.sqrt(refmd.COULOMB)).astype(F32)
    eps, rc = 78.5, cfg["cutoff"]
    krf = (1 / rc**3) * (eps - 1) / (2 * eps + 1)
    crf = (1 / rc) * (3 * eps) / (2 * eps + 1)
    errs = []
    for variant in (0, 1, 2):
        out = np.zeros((len(pos), 3), F32)
        hc.hc_pair_forces(variant, len(pos), len(pairs), p(np.ascontiguousarray(pairs)), p(pos), p(qs), p(types), nt,
                          p(np.ascontiguousarray(AB)), p(box), (1 << 5) | (1 << 6), C.c_float(rc), 1,
                          C.c_float(cfg["switch_dist"]), 1, C.c_float(krf), C.c_float(crf), p(out))
        errs.append(np.abs(out.astype(np.float64) - f64[0].numpy()).max())
    print(f"max |dF| vs fp64 oracle: float path {errs[0]:.2e}, fixed-point path {errs[1]:.2e}, "
          f"fixed-point + packed arithmetic {errs[2]:.2e} (drift={drift})")
    assert errs[1] < 1e-4 and errs[2] < 1e-4
    assert errs[1] <= errs[0] * 1.25 + 5e-6
    assert errs[2] <= errs[1] * 1.25 + 5e-6


def test_packed_two_partner_coefficient_matches_oracle(hc):
    """physics.cuh pair_coef2 (k_pair_fx2): LJ with switch + reaction field, explicit-force
    convention, two partners per packed evaluation."""
    rng = np.random.default_rng(8)
    n, cutoff, switch, eps = 40000, 9.0, 7.5, 78.5
    r = rng.uniform(1.6, cutoff, n)
    s = (r * r).astype(F32)
    qq = (refmd.COULOMB * rng.uniform(-1, 1, n) * rng.uniform(-1, 1, n)).astype(F32)
    A = rng.uniform(1e4, 6e5, n).astype(F32)
    B = rng.uniform(10, 600, n).astype(F32)
    krf = (1 / cutoff**3) * (eps - 1) / (2 * eps + 1)
    out = np.zeros(n, F32)
    hc.hc_pair_coef2(n, p(s), p(qq), p(A), p(B), C.c_float(cutoff), C.c_float(switch), C.c_float(krf), p(out))
    dist = torch.tensor(np.sqrt(s.astype(np.float64)))
    _, lj_f = refmd.lj_pair(dist, torch.tensor(A.astype(np.float64)), torch.tensor(B.astype(np.float64)), 1, switch, cutoff)
    _, el_f = refmd.coulomb_pair(dist, torch.tensor(qq.astype(np.float64) / refmd.COULOMB), torch.ones(n, dtype=torch.float64), 1, cutoff, True, eps)
    want = ((lj_f + el_f) / dist).numpy()
    assert np.abs(out - want).max() <= 4e-6 * max(1.0, np.abs(want).max())
    # and agrees with the scalar kernel arithmetic (pair_terms<1>) to rounding
    o5 = [np.zeros(n, F32) for _ in range(5)]
    crf = (1 / cutoff) * (3 * eps) / (2 * eps + 1)
    hc.hc_pair_terms(n, p(s), p(qq), p(A), p(B), (1 << 5) | (1 << 6), 1, C.c_float(cutoff), 1, C.c_float(switch), 1,
                     C.c_float(krf), C.c_float(crf), *[p(o) for o in o5])
    scalar = o5[4].astype(np.float64) / np.sqrt(s.astype(np.float64))
    assert np.abs(out - scalar).max() <= 4e-6 * max(1.0, np.abs(scalar).max())
    # energies of the ENERGY instantiation
    c2, elj, eel = (np.zeros(n, F32) for _ in range(3))
    hc.hc_pair_coef2e(n, p(s), p(qq), p(A), p(B), C.c_float(cutoff), C.c_float(switch), C.c_float(krf), C.c_float(crf),
                      p(c2), p(elj), p(eel))
    assert np.array_equal(c2, out)
    lj_e, _ = refmd.lj_pair(dist, torch.tensor(A.astype(np.float64)), torch.tensor(B.astype(np.float64)), 1, switch, cutoff)
    el_e, _ = refmd.coulomb_pair(dist, torch.tensor(qq.astype(np.float64) / refmd.COULOMB), torch.ones(n, dtype=torch.float64), 1, cutoff, True, eps)
    assert np.abs(elj - lj_e.numpy()).max() <= 3e-6 * max(1.0, np.abs(lj_e.numpy()).max())
    assert np.abs(eel - el_e.numpy()).max() <= 3e-6 * max(1.0, np.abs(el_e.numpy()).max())


@pytest.mark.parametrize("scale", [1, 50, 1900])
def test_fx_decision_bound_holds_far_from_the_origin(hc, scale):
    """The margin grows with the largest |coordinate| (the reference's own rounding of p_i - p_j and
    L*n does): the bound must hold and the classification stay sound up to the 2000-box limit."""
    rng = np.random.default_rng(1)
    box = np.array([99.93, 87.41, 120.07], F32)
    n = 100000
    pi = (rng.uniform(-scale, scale, (n, 3)) * box).astype(F32)
    d = rng.normal(size=(n, 3))
    d *= (rng.uniform(8.95, 9.05, n) / np.linalg.norm(d, axis=1))[:, None]
    pj = (pi.astype(np.float64) - d + rng.integers(-3, 4, (n, 3)) * box.astype(np.float64)).astype(F32)
    pmax = float(max(np.abs(pi).max(), np.abs(pj).max()))
    _, s, cls, margin = fx_decide(hc, pi, pj, box, 9.0, 11.0, pmax)
    _, sref, inside = decide(hc, pi, pj, box, 9.0)
    assert np.abs(s - sref).max() <= margin
    assert ((cls == 2) | ((cls == 1) == inside)).all()


@pytest.mark.parametrize("rfa,switch", [(True, 7.5), (False, None), (True, None), (False, 6.0)])
def test_packed_coefficient_covers_every_lj_electrostatics_combination(hc, rfa, switch):
    """make_switch_consts turns switching / the reaction field off through its constants."""
    rng = np.random.default_rng(9)
    n, cutoff, eps = 20000, 9.0, 78.5
    r = rng.uniform(1.6, cutoff, n)
    s = (r * r).astype(F32)
    qq = (refmd.COULOMB * rng.uniform(-1, 1, n) * rng.uniform(-1, 1, n)).astype(F32)
    A = rng.uniform(1e4, 6e5, n).astype(F32)
    B = rng.uniform(10, 600, n).astype(F32)
    krf = (1 / cutoff**3) * (eps - 1) / (2 * eps + 1)
    crf = (1 / cutoff) * (3 * eps) / (2 * eps + 1)
    c, elj, eel = (np.zeros(n, F32) for _ in range(3))
    hc.hc_pair_coef2g(n, p(s), p(qq), p(A), p(B), C.c_float(cutoff), int(switch is not None), C.c_float(switch or 0.0), int(rfa),
                      C.c_float(krf), C.c_float(crf), p(c), p(elj), p(eel))
    dist = torch.tensor(np.sqrt(s.astype(np.float64)))
    lj_e, lj_f = refmd.lj_pair(dist, torch.tensor(A.astype(np.float64)), torch.tensor(B.astype(np.float64)), 1, switch, cutoff)
    el_e, el_f = refmd.coulomb_pair(dist, torch.tensor(qq.astype(np.float64) / refmd.COULOMB), torch.ones(n, dtype=torch.float64), 1, cutoff, rfa, eps)
    want = ((lj_f + el_f) / dist).numpy()
    assert np.abs(c - want).max() <= 4e-6 * max(1.0, np.abs(want).max())
    assert np.abs(elj - lj_e.numpy()).max() <= 3e-6 * max(1.0, np.abs(lj_e.numpy()).max())
    assert np.abs(eel - el_e.numpy()).max() <= 3e-6 * max(1.0, np.abs(el_e.numpy()).max())


def test_packed_arithmetic_without_a_box(hc):
    """k_pair2_open's arithmetic (float records, reference differences, packed evaluation) on the alanine-dipeptide
    fixture without a box: non-bonded forces against the fp64 oracle on the fp32 pair set."""
    g = load_golden("ala2_nobox_rf")
    cfg = golden_cfg(g)
    par64 = __import__("conftest").params_from_golden(g, precision=torch.float64)
    of = refmd.OracleForces(par64, ["lj", "electrostatics"], decision_dtype=torch.float32, **cfg)
    pos = g["coords"].astype(F32).copy()
    pos_t = torch.tensor(pos)[None]
    f64 = torch.zeros(1, len(pos), 3, dtype=torch.float64)
    of.compute(pos_t.double(), torch.zeros(1, 3, 3, dtype=torch.float64), f64)
    pairs = of.neighbour_pairs(pos_t[0], torch.zeros(3)).numpy().astype(np.int32)
    types = g["par_types"].astype(np.int32)
    nt = int(types.max()) + 1
    A, B = par64.get_AB()
    AB = np.stack([A.numpy(), B.numpy()], -1).astype(F32).reshape(nt, nt, 2)
    qs = (g["par_charges"] * math.sqrt(refmd.COULOMB)).astype(F32)
    eps, rc = 78.5, cfg["cutoff"]
    krf = (1 / rc**3) * (eps - 1) / (2 * eps + 1)
    crf = (1 / rc) * (3 * eps) / (2 * eps + 1)
    out = np.zeros((len(pos), 3), F32)
    hc.hc_pair_forces(3, len(pos), len(pairs), p(np.ascontiguousarray(pairs)), p(pos), p(qs), p(types), nt, p(np.ascontiguousarray(AB)),
                      p(np.ones(3, F32)), (1 << 5) | (1 << 6), C.c_float(rc), int(cfg["switch_dist"] is not None),
                      C.c_float(cfg["switch_dist"] or 0.0), int(cfg["rfa"]), C.c_float(krf), C.c_float(crf), p(out))
    err = np.abs(out.astype(np.float64) - f64[0].numpy()).max()
    assert err < 1e-4, err
