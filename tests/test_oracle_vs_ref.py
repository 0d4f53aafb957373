"""Pins the oracle.  (1) the real reference built in oracle/_ref reproduces the known answers recorded at
survey time; (2) the plain-C restatement oracle/cutout_oracle.c is byte-identical to the
real reference on seeded shells, edge sets and halo set-ups.  CPU only."""
import hashlib

import numpy as np
import pytest

import harness as H

needs_ref = pytest.mark.skipif(not H.ref_available(), reason="oracle/_ref not built (needs /root/reference)")

KAT = {  # K and sha256[:16] of id, theta, phi, redshift, x
    "narrow": (30889, "4ce6a39d99f49613", "2047f5c587dafe29", "eae9b8243c48a6f3", "884131f3e7d4b1c7", "8c51614734111ec0"),
    "wide": (759498, "182f9c350d689719", "64e730ea83c6f615", "ba1a55a5c63ee561", "6e22101aea8c830d", "59abcebeb5ecb773"),
    "halo": (58, "9f081aebb94897c7", "ef0c7e02d2ac39c5", "86ceb0d5153a37db", "8e10fe518765701d", "af5695107465eb98"),
}


def _sig(out):
    return (out["id"].shape[0],) + tuple(hashlib.sha256(out[c].tobytes()).hexdigest()[:16]
                                         for c in ("id", "theta", "phi", "redshift", "x"))


@pytest.fixture(scope="module")
def kat_cols():
    cols = H.shell_fixture(10_000_000, seed=20261019)
    assert hashlib.sha256(cols["x"].tobytes()).hexdigest()[:16] == "e36442e5e3b62853"
    assert hashlib.sha256(cols["id"].tobytes()).hexdigest()[:16] == "0379cc26255dc5d3"
    return cols


def test_kat_oracle_const(kat_cols):
    assert _sig(H.oracle_cutout_const(kat_cols, H.oracle_cli_bounds(87, 2), H.oracle_cli_bounds(2, 2))) == KAT["narrow"]
    assert _sig(H.oracle_cutout_const(kat_cols, H.oracle_cli_bounds(90, 20), H.oracle_cli_bounds(0, 20))) == KAT["wide"]


def test_kat_oracle_halo(kat_cols):
    out, rough = H.oracle_cutout_halo(kat_cols, H.oracle_halo_setup([150, 20, 12], 10.0))
    assert _sig(out) == KAT["halo"] and rough == 516


@needs_ref
def test_kat_reference(kat_cols):
    with H.RefRun(kat_cols) as r:
        assert _sig(r.const(H.oracle_cli_bounds(87, 2), H.oracle_cli_bounds(2, 2))) == KAT["narrow"]
    with H.RefRun(kat_cols) as r:
        assert _sig(r.const(H.oracle_cli_bounds(90, 20), H.oracle_cli_bounds(0, 20))) == KAT["wide"]
    with H.RefRun(kat_cols) as r:
        assert _sig(r.halo([150, 20, 12], 10.0)) == KAT["halo"]
        assert (r.last_count, r.last_rough_count) == (58, 516)


@needs_ref
@pytest.mark.parametrize("tc,pc", [((87, 2), (2, 2)), ((90, 20), (0, 20)), ((45, 1), (45, 1)), ((3, 3), (60, 30)),
                                   ((89.9, 0.2), (89.5, 0.6))])
def test_const_oracle_equals_reference(tc, pc):
    theta_cut = H.oracle_cli_bounds(*tc); phi_cut = H.oracle_cli_bounds(*pc)
    cols = H.concat_cols(H.shell_fixture(400_000, seed=11), H.edge_fixture(theta_cut, phi_cut, octant_only=True),
                         H.shell_fixture(100_000, seed=12, octant=False))
    want = None
    with H.RefRun(cols) as r:
        want = r.const(theta_cut, phi_cut)
    got = H.oracle_cutout_const(cols, theta_cut, phi_cut)
    assert want["id"].shape[0] > 0
    H.assert_cols_equal(got, want)


HALOS = [((150, 20, 12), 10.0), ((150, 20, 12), 120.0), ((30, 170, 100), 30.0), ((100, 100, 100), 5.0),
         ((10, 10, 250), 60.0), ((200, 1, 1), 20.0), ((120, -60, 40), 45.0), ((-90, 80, -150), 60.0)]


@needs_ref
@pytest.mark.parametrize("pos,box", HALOS)
@pytest.mark.parametrize("pos_only", [False, True])
def test_halo_oracle_equals_reference(pos, box, pos_only):
    h = H.oracle_halo_setup(pos, box)
    cols = H.concat_cols(H.shell_fixture(300_000, seed=21), H.shell_fixture(300_000, seed=22, octant=False),
                         H.edge_fixture(h.theta_rough, h.phi_rough, seed=5))
    with H.RefRun(cols) as r:
        want = r.halo(pos, box, pos_only=pos_only)
        want_rough = r.last_rough_count
    got, rough = H.oracle_cutout_halo(cols, h, pos_only=pos_only)
    names = H.OUT_COLS_POSONLY if pos_only else H.OUT_COLS
    H.assert_cols_equal(got, want, names)
    assert rough == want_rough


@needs_ref
def test_halo_setup_equals_reference_functions():
    rng = np.random.default_rng(3)
    pts = [(150, 20, 12), (1, 0, 0), (0, 1, 0), (0, 0, 1), (5, 5, 5), (-3, 2, 1)]
    pts += [tuple(rng.uniform(-300, 300, 3)) for _ in range(200)]
    for p in pts:
        h = H.oracle_halo_setup(p, 10.0)
        w = H.ref_rotation_setup(p)
        for nme in ("k", "K", "R", "R_inv"):
            assert np.array(getattr(h, nme), dtype=np.float32).tobytes() == w[nme].tobytes(), (p, nme)
        assert np.float32(h.beta).tobytes() == w["beta"].tobytes(), p
        assert np.float32(h.halo_r).tobytes() == w["halo_r"].tobytes(), p


@needs_ref
def test_small_helpers_equal_reference():
    lib = H.ref(); o = H.oracle()
    for a in np.random.default_rng(1).uniform(0.004, 1.0, 1000).astype(np.float32):
        assert np.float32(lib.ref_aToZ(float(a))).tobytes() == np.float32(o.oracle_aToZ(float(a))).tobytes()
    for z in (0, 0.5, 1, 2, 0.0245, 3.7, 10, 199):
        assert lib.ref_zToStep(z) == o.oracle_zToStep(z, 499, 200.0)
    assert [lib.ref_zToStep(z) for z in (0, 1, 0.5, 2)] == [499, 248, 331, 164]
    for s in range(1, 500, 7):
        assert np.float32(lib.ref_stepToZ(float(s))).tobytes() == np.float32(o.oracle_stepToZ(float(s), 499, 200.0)).tobytes()


@needs_ref
def test_tail_drop_equals_reference():
    """the float rank map of util.cpp:106-108 drops trailing particles once Np > 2^24"""
    lib = H.ref(); o = H.oracle()
    for n_p in (1, 2, 1000, 2 ** 24 - 1, 2 ** 24, 2 ** 24 + 1, 2 ** 24 + 3, 20_000_001, 33_554_433, 100_000_000):
        for ranks in (1, 2, 8):
            # (for ranks > 1 this is the count kept out of ONE rank's block of n_p particles)
            assert lib.ref_scatter_kept(n_p, ranks) == o.oracle_scatter_kept(n_p, ranks), (n_p, ranks)
    assert o.oracle_scatter_kept(100_000_000, 1) == 100_000_000 - 4
    assert o.oracle_scatter_kept(1_000_000_000, 1) == 1_000_000_000 - 32


def test_checkers_carry_glibc_239_acosf_atanf():
    """theta/phi are glibc's (not correctly rounded) acosf/atanf results: the oracle and the reference build link
    glibc 2.39's own objects for the two functions (hidden, from the image's static libm) instead of binding to the
    libm.so.6 of whatever box runs them, so a newer glibc cannot silently change the checker (oracle/Makefile)"""
    import os
    import subprocess
    if not os.path.exists("/usr/lib/x86_64-linux-gnu/libm-2.39.a") and not os.path.isdir(os.path.join(H.ORACLE_DIR, "pin")):
        pytest.skip("no static glibc 2.39 libm on this box and no prebuilt pin")
    libs = [os.path.join(H.ORACLE_DIR, "liboracle.so")]
    if H.ref_available():
        libs.append(os.path.join(H.REF_DIR, "libcutout_ref.so"))
    for lib in libs:
        syms = subprocess.check_output(["nm", lib]).decode()
        assert " t __ieee754_acosf" in syms and " t __atanf" in syms, lib          # defined inside, local
        dyn = subprocess.check_output(["nm", "-D", lib]).decode()
        assert " U acosf" not in dyn and " U atanf" not in dyn, lib                 # not bound to libm.so.6
