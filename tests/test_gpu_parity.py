"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against the oracle
(oracle/liboracle.so, pinned to the real reference by tests/test_oracle_vs_ref.py) on the same seeded inputs.
Bar: byte-exact for every column, the survivor set and its order."""
import os

import numpy as np
import pytest

import harness as H
from __graft_entry__ import load_package

cc = load_package()
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cc.Context(0)
    yield c
    c.close()


@pytest.fixture()
def cells_mode(ctx, request):
    """many halos: run the test with the staged kernels (k2c_sift / k2c_match / k2c_exact), with the one-kernel scan
    (k2_cutout_halo_cells) -- the library picks between them per halo set from the measured bitmap hit rate -- and with
    the particle index of the resident step (k2p_query), which it builds for a step that is cut repeatedly"""
    mode = request.getfixturevalue("cells")
    if mode == "index":  # the particle index of the resident step (k2p_*), built inside the first many-halo cutout
        ctx.set_option("index", "always")
    else:
        ctx.set_option("index", "never")
        ctx.set_option("cells", mode)
    yield
    ran = ctx.stats_ex()["scan_kernels"]
    ctx.set_option("cells", "auto")
    ctx.set_option("index", "auto")
    assert ran == {"staged": "k2c_sift+k2c_match+k2c_exact", "mono": "k2_cutout_halo_cells", "index": "k2p_query"}[mode]


def _same_float_bits(a, b):
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


# ------------------------------------------------------------------------------------------------ arithmetic
def test_device_libm_equals_host_restatement_on_all_floats(ctx):
    """device acosf/atanf restatement == host build of the same header on all 2^32 inputs (the host build is
    checked against glibc on all 2^32 inputs by the CPU suite)"""
    chunk = 1 << 26
    for op in (0, 1):
        for c0 in range(0, 1 << 32, chunk):
            x = np.arange(c0, c0 + chunk, dtype=np.uint64).astype(np.uint32).view(np.float32)
            d = ctx.debug_eval(op, x)
            h = cc.debug_eval_host(op, x)
            assert _same_float_bits(d, h), (op, c0)


def test_device_arcsec_scaling_equals_specification_on_all_floats(ctx):
    """op 6: the device's division-free radians->arcsec scaling vs the specification form (true fp64 division, host)
    on all 2^32 float bit patterns"""
    chunk = 1 << 26
    for c0 in range(0, 1 << 32, chunk):
        x = np.arange(c0, c0 + chunk, dtype=np.uint64).astype(np.uint32).view(np.float32)
        assert _same_float_bits(ctx.debug_eval(6, x), cc.debug_eval_host(6, x)), c0


def test_device_libm_equals_glibc_sample(ctx):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1.0, 1.0, 2_000_000), rng.normal(0, 50, 2_000_000),
                        [0.0, -0.0, 1.0, -1.0, 0.5, -0.5, np.inf, -np.inf, 1e-30, 3e7, 4e7, 0.4375, 0.6875, 1.1875,
                         2.4375]]).astype(np.float32)
    assert _same_float_bits(ctx.debug_eval(0, x), H.oracle_libm("acosf", x))
    assert _same_float_bits(ctx.debug_eval(1, x), H.oracle_libm("atanf", x))


def test_device_theta_phi_equals_oracle(ctx):
    cols = H.concat_cols(H.shell_fixture(1_000_000, seed=3), H.shell_fixture(1_000_000, seed=4, octant=False),
                         H.edge_fixture((306000, 320400), (0, 14400)))
    x, y, z = cols["x"], cols["y"], cols["z"]
    th1, ph1 = H.oracle_theta_phi(x, y, z, guards=False)
    th2, ph2 = H.oracle_theta_phi(x, y, z, guards=True)
    assert _same_float_bits(ctx.debug_eval(2, x, y, z), th1)
    assert _same_float_bits(ctx.debug_eval(3, x, y), ph1)
    assert _same_float_bits(ctx.debug_eval(4, x, y), ph2)
    a = cols["a"]
    want = cc.debug_eval_host(5, a)
    assert ctx.debug_eval(5, a).tobytes() == want.tobytes()


def test_synth_device_equals_host(ctx):
    for kind in (0, 1):
        n = 3_000_001
        ctx.synth(n, seed=99, index_base=12345, kind=kind, box=250.0)
        d = ctx.download(0, n)
        h = cc.synth_host(n, seed=99, index_base=12345, kind=kind, box=250.0)
        for k in cc.IN_COLS:
            assert d[k].tobytes() == h[k].tobytes(), (kind, k)


# ------------------------------------------------------------------------------------------------ use case 1
WINDOWS = [((87, 2), (2, 2)), ((90, 20), (0, 20)), ((45, 1), (45, 1)), ((3, 3), (60, 30)), ((89.9, 0.2), (89.5, 0.6)),
           ((45, 50), (45, 50)), ((0.5, 0.4), (10, 5)), ((120, 10), (20, 5))]


@pytest.mark.parametrize("tc,pc", WINDOWS)
def test_const_equals_oracle(ctx, tc, pc):
    theta_cut, phi_cut = cc.cli_bounds(*tc), cc.cli_bounds(*pc)
    cols = H.concat_cols(H.shell_fixture(700_000, seed=11), H.edge_fixture(theta_cut, phi_cut, octant_only=True),
                         H.shell_fixture(300_123, seed=12, octant=False), H.edge_fixture(theta_cut, phi_cut, seed=9))
    ctx.upload(cols)
    ctx.cutout_const(theta_cut, phi_cut)
    counts, rough = ctx.sync()
    got = ctx.fetch(0, names=cc.OUT_COLS + ("index",))
    want = H.oracle_cutout_const(cols, theta_cut, phi_cut)
    assert int(counts[0]) == want["id"].shape[0]
    H.assert_cols_equal(got, want, what=f"const {tc} {pc}: ")
    # `index` really is the input position of each survivor, in increasing order (stable compaction)
    assert np.array_equal(cols["id"][got["index"]], got["id"]) and np.all(np.diff(got["index"]) > 0)
    st = ctx.stats()
    assert st["survivors"] == int(counts[0]) and st["candidates"] >= st["survivors"]


def test_const_kat_10m(ctx):
    """cfg1 of BASELINE.json: the survey-time known answers of the patched reference"""
    import hashlib
    cols = H.shell_fixture(10_000_000, seed=20261019)
    ctx.upload(cols)
    for (tc, pc, k, h_id, h_th) in (((87, 2), (2, 2), 30889, "4ce6a39d99f49613", "2047f5c587dafe29"),
                                     ((90, 20), (0, 20), 759498, "182f9c350d689719", "64e730ea83c6f615")):
        ctx.cutout_const(cc.cli_bounds(*tc), cc.cli_bounds(*pc))
        counts, _ = ctx.sync()
        got = ctx.fetch(0)
        assert int(counts[0]) == k
        assert hashlib.sha256(got["id"].tobytes()).hexdigest()[:16] == h_id
        assert hashlib.sha256(got["theta"].tobytes()).hexdigest()[:16] == h_th
        want = H.oracle_cutout_const(cols, cc.cli_bounds(*tc), cc.cli_bounds(*pc))
        H.assert_cols_equal(got, want, what="kat: ")


@pytest.mark.parametrize("n", [0, 1, 5, 2047, 2048, 2049, 4096 + 17])
def test_const_ragged_sizes(ctx, n):
    theta_cut, phi_cut = cc.cli_bounds(45, 44), cc.cli_bounds(45, 44)
    cols = H.slice_cols(H.shell_fixture(10_000, seed=2), 0, n)
    ctx.upload(cols)
    ctx.cutout_const(theta_cut, phi_cut)
    counts, _ = ctx.sync()
    want = H.oracle_cutout_const(cols, theta_cut, phi_cut)
    assert int(counts[0]) == want["id"].shape[0]
    H.assert_cols_equal(ctx.fetch(0), want)


def test_const_everything_survives_and_capacity_regrows(ctx):
    """window covering the whole octant: K == N, far beyond the initial output capacity"""
    cols = H.shell_fixture(1_500_000, seed=8)
    theta_cut, phi_cut = np.array([-10.0, 400000.0], np.float32), np.array([-10.0, 400000.0], np.float32)
    ctx.upload(cols)
    ctx.cutout_const(theta_cut, phi_cut)
    counts, _ = ctx.sync()
    want = H.oracle_cutout_const(cols, theta_cut, phi_cut)
    assert int(counts[0]) == want["id"].shape[0] == 1_500_000
    H.assert_cols_equal(ctx.fetch(0), want)


def test_const_special_values_and_global_index(ctx):
    """NaN / Inf / zero / subnormal scale factors (redshift = 1/a - 1 must carry the host's bits, including the
    quieted NaN payload), NaN and Inf coordinates (rejected like the reference), and a shard whose global index
    base is beyond 2^32"""
    cols = H.shell_fixture(400_000, seed=33)
    rng = np.random.default_rng(9)
    a = cols["a"]
    pick = rng.choice(400_000, 40_000, replace=False)
    specials = np.array([np.nan, -np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-45, 1e-39, 3.4e38, 1.0, -1.0, 0.5],
                        dtype=np.float32)
    a[pick] = specials[rng.integers(0, len(specials), pick.shape[0])]
    a[pick[:1000]] = np.array([0x7fa00001, 0xffa12345, 0x7f800001], dtype=np.uint32).view(np.float32)[rng.integers(0, 3, 1000)]
    bad = rng.choice(400_000, 3_000, replace=False)
    for k, col in enumerate(("x", "y", "z")):
        cols[col][bad[k::3]] = np.array([np.nan, np.inf, -np.inf, 0.0], dtype=np.float32)[rng.integers(0, 4, bad[k::3].shape[0])]
    theta_cut, phi_cut = cc.cli_bounds(60, 29), cc.cli_bounds(45, 44)
    base = 5_000_000_000
    ctx.upload(cols, index_base=base)
    ctx.cutout_const(theta_cut, phi_cut)
    counts, _ = ctx.sync()
    got = ctx.fetch(0, names=cc.OUT_COLS + ("index",))
    want = H.oracle_cutout_const(cols, theta_cut, phi_cut)
    assert int(counts[0]) == want["id"].shape[0] > 100_000
    H.assert_cols_equal(got, want, what="special values: ")
    assert np.isnan(got["redshift"]).any() and np.isinf(got["redshift"]).any()
    assert np.array_equal(got["index"] - base, got["id"]) and got["index"].min() >= base


def test_const_tiny_temp_regions_regrow(ctx):
    """all survivors sit in a narrow index range, so a handful of CTA temp regions overflow their first-guess
    capacity: the cutout must notice, grow and re-run"""
    n = 3_000_000
    cols = H.shell_fixture(n, seed=34)
    theta_cut, phi_cut = cc.cli_bounds(87, 2), cc.cli_bounds(2, 2)
    # park everything outside the window, then drop 150k in-window particles into one contiguous block
    th, ph = H.oracle_theta_phi(cols["x"], cols["y"], cols["z"], guards=False)
    inside = (th > theta_cut[0]) & (th < theta_cut[1]) & (ph > phi_cut[0]) & (ph < phi_cut[1])
    for c in ("x", "y", "z"):
        cols[c][inside] = -cols[c][inside]
    rng = np.random.default_rng(35)
    k = 150_000
    t = np.deg2rad(rng.uniform(85.2, 88.8, k)); p = np.deg2rad(rng.uniform(0.2, 3.8, k)); r = rng.uniform(100, 300, k)
    blk = dict(x=r * np.sin(t) * np.cos(p), y=r * np.sin(t) * np.sin(p), z=r * np.cos(t))
    lo = 1_234_567
    for c in ("x", "y", "z"):
        cols[c][lo:lo + k] = blk[c].astype(np.float32)
    with cc.Context(0) as fresh:  # first-guess capacities
        fresh.upload(cols)
        fresh.cutout_const(theta_cut, phi_cut)
        counts, _ = fresh.sync()
        want = H.oracle_cutout_const(cols, theta_cut, phi_cut)
        assert int(counts[0]) == want["id"].shape[0] == k
        H.assert_cols_equal(fresh.fetch(0), want, what="regrow: ")


def test_const_unaligned_device_columns(ctx):
    """attach columns whose device addresses are not 16-byte aligned: the scalar-load kernel variant"""
    import torch
    cols = H.shell_fixture(300_000, seed=13)
    n = 300_000 - 3
    keep, ptrs = [], {}
    for k in cc.IN_COLS:
        t = torch.from_numpy(cols[k]).cuda()
        keep.append(t)
        ptrs[k] = t.data_ptr() + 3 * t.element_size()
    torch.cuda.synchronize()
    sub = H.slice_cols(cols, 3, 300_000)
    theta_cut, phi_cut = cc.cli_bounds(60, 10), cc.cli_bounds(30, 10)
    ctx.attach(n, ptrs)
    ctx.cutout_const(theta_cut, phi_cut)
    ctx.sync()
    H.assert_cols_equal(ctx.fetch(0), H.oracle_cutout_const(sub, theta_cut, phi_cut))
    # the halo kernels on the same unaligned columns (n is not a multiple of the tile either): one halo (plain scan)
    # and nine halos (cell-index scan, which then reads its tiles without the cp.async prefetch)
    rng = np.random.default_rng(4)
    pos = np.vstack([[150, 20, 12], rng.uniform(40, 250, (8, 3))]).astype(np.float32)
    for nh in (1, 9):
        ctx.cutout_halo([cc.halo_setup(p, 90.0).halo for p in pos[:nh]])
        counts, rough = ctx.sync()
        for h in range(nh):
            want, want_rough = H.oracle_cutout_halo(sub, H.oracle_halo_setup(pos[h], 90.0))
            assert int(counts[h]) == want["id"].shape[0] and int(rough[h]) == want_rough
            H.assert_cols_equal(ctx.fetch(h), want, what=f"unaligned, {nh} halos, halo {h}: ")


def test_const_shards_concatenate_to_single_rank_order(ctx):
    """index sharding (one contiguous range per GPU): concatenating the shard outputs in rank order equals the
    single-rank output, so file offsets = exclusive scan of counts"""
    cols = H.shell_fixture(1_000_003, seed=21)
    theta_cut, phi_cut = cc.cli_bounds(70, 15), cc.cli_bounds(40, 15)
    want = H.oracle_cutout_const(cols, theta_cut, phi_cut)
    parts = []
    bounds = [0, 250_000, 500_001, 777_777, 1_000_003]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        ctx.upload(H.slice_cols(cols, lo, hi), index_base=lo)
        ctx.cutout_const(theta_cut, phi_cut)
        ctx.sync()
        parts.append(ctx.fetch(0, names=cc.OUT_COLS + ("index",)))
    got = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    H.assert_cols_equal(got, want)
    assert np.array_equal(got["index"], got["id"])  # shell_fixture ids are the global index


# ------------------------------------------------------------------------------------------------ use case 2
HALOS = [((150, 20, 12), 10.0), ((150, 20, 12), 120.0), ((30, 170, 100), 30.0), ((100, 100, 100), 5.0),
         ((10, 10, 250), 60.0), ((200, 1, 1), 20.0), ((120, -60, 40), 45.0), ((-90, 80, -150), 60.0),
         ((1, 2, 300), 240.0), ((0.5, -0.5, -200), 90.0)]


@pytest.fixture(scope="module")
def halo_cols():
    h = H.oracle_halo_setup((150, 20, 12), 120.0)
    return H.concat_cols(H.shell_fixture(600_000, seed=21), H.shell_fixture(600_000, seed=22, octant=False),
                         H.edge_fixture(h.theta_rough, h.phi_rough, seed=5), H.edge_fixture(h.theta_cut, h.phi_cut, seed=6))


@pytest.mark.parametrize("pos_only", [False, True])
def test_halo_batch_equals_oracle(ctx, halo_cols, pos_only):
    """all halos in ONE pass; each halo's survivors, their (theta_u, index) order, columns and rough count"""
    infos = [cc.halo_setup(p, b) for p, b in HALOS]
    up = dict(halo_cols)
    if pos_only:
        for k in ("vx", "vy", "vz", "rotation", "replication"):
            up[k] = None
    ctx.upload(up)
    ctx.cutout_halo([i.halo for i in infos], pos_only=pos_only)
    counts, rough = ctx.sync()
    names = cc.OUT_COLS_POSONLY if pos_only else cc.OUT_COLS
    total = 0
    for h, (p, b) in enumerate(HALOS):
        want, want_rough = H.oracle_cutout_halo(halo_cols, H.oracle_halo_setup(p, b), pos_only=pos_only)
        assert int(counts[h]) == want["id"].shape[0], (p, b)
        assert int(rough[h]) == want_rough, (p, b)
        H.assert_cols_equal(ctx.fetch(h), want, names, what=f"halo {p} {b}: ")
        total += int(counts[h])
    assert total > 500


def test_halo_sort_keys_and_tail_drop(ctx, halo_cols):
    p, b = (150, 20, 12), 120.0
    info = cc.halo_setup(p, b)
    n = halo_cols["x"].shape[0]
    ctx.upload(halo_cols)
    # n_limit: only a prefix of the step is considered (the reference's Q7 tail drop at Np > 2^24)
    limit = n - 100_000
    ctx.cutout_halo([info.halo], n_limit_global=limit)
    ctx.sync()
    got = ctx.fetch(0, names=cc.OUT_COLS + ("theta_u", "index"))
    want, _ = H.oracle_cutout_halo(H.slice_cols(halo_cols, 0, limit), H.oracle_halo_setup(p, b))
    H.assert_cols_equal(got, want)
    d_th, d_ix = np.diff(got["theta_u"]), np.diff(got["index"])
    assert np.all((d_th > 0) | ((d_th == 0) & (d_ix > 0)))  # processLC.cpp:1214-1221: stable sort on theta_u
    th_u, _ = H.oracle_theta_phi(halo_cols["x"], halo_cols["y"], halo_cols["z"], guards=True)
    assert np.array_equal(th_u[got["index"]].view(np.uint32), got["theta_u"].view(np.uint32))


@pytest.mark.parametrize("cells", ["staged", "mono", "index"])
def test_halo_many_halos_one_pass(ctx, cells_mode, cells):
    """more halos than one launch batch (K2_MAX_HALOS = 1024)"""
    rng = np.random.default_rng(5)
    cols = H.shell_fixture(400_000, seed=31)
    pos = rng.uniform(20, 250, (1100, 3)).astype(np.float32)
    infos = [cc.halo_setup(p, 30.0) for p in pos]
    ctx.upload(cols)
    ctx.cutout_halo([i.halo for i in infos])
    counts, rough = ctx.sync()
    everything, off = ctx.fetch_all(counts)  # one copy per column; halo h = rows off[h]:off[h+1]
    assert int(off[-1]) == int(counts.sum()) == everything["id"].shape[0]
    for h in list(range(0, 1100, 97)) + [1023, 1024, 1099]:
        want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(pos[h], 30.0))
        assert int(counts[h]) == want["id"].shape[0] and int(rough[h]) == want_rough
        H.assert_cols_equal(ctx.fetch(h), want, what=f"halo {h}: ")
        H.assert_cols_equal({k: v[off[h]:off[h + 1]] for k, v in everything.items()}, want, what=f"halo {h} (fetch_all): ")
    with pytest.raises(Exception):
        ctx.fetch_all(counts[:-1])  # wrong total


@pytest.mark.parametrize("cells", ["staged", "mono", "index"])
def test_halo_cell_index_overlapping_halos_and_special_positions(ctx, cells_mode, cells):
    """the cell-index kernel (>= 8 halos): 48 halos whose windows all overlap, so a 256-particle tile owes far more
    (particle, halo) pairs than the warp's pair buffer holds (several passes per tile), plus positions the
    pre-filter cannot judge (zero / tiny / huge / Inf coordinates, all octants) that take the every-halo path"""
    rng = np.random.default_rng(17)
    cols = H.concat_cols(H.shell_fixture(300_000, seed=51), H.shell_fixture(100_000, seed=52, octant=False))
    n = cols["x"].shape[0]
    bad = rng.choice(n, 1_500, replace=False)
    # one special coordinate per particle, none that makes theta_u NaN: the reference orders theta_u with
    # std::stable_sort / lower_bound (processLC.cpp:1214-1221, :1334-1337), which is unspecified once a NaN is in
    # the array (DESIGN.md "NaN theta"), so there is nothing to be equal to
    vals = np.array([np.inf, -np.inf, 0.0, -0.0, 1e-12, -1e-12, 1e-40, 3e38], dtype=np.float32)
    for k, col in enumerate(("x", "y", "z")):
        v = vals if col != "z" else vals[2:]
        cols[col][bad[k::3]] = v[rng.integers(0, len(v), bad[k::3].shape[0])]
    cols["x"][bad[:200]] = np.float32(1e-12)  # tiny x next to ordinary y, z: phi_u close to 90 degrees
    centre = np.array([120.0, 90.0, 100.0])
    pos = (centre + rng.uniform(-4, 4, (48, 3))).astype(np.float32)
    boxes = rng.uniform(40.0, 90.0, 48)
    infos = [cc.halo_setup(p, b) for p, b in zip(pos, boxes)]
    ctx.upload(cols)
    ctx.cutout_halo([i.halo for i in infos])
    counts, rough = ctx.sync()
    if cells == "mono":
        assert ctx.stats()["candidates"] > 20 * int(counts.max())  # many pairs per particle went through the exact path
    for h in range(0, 48, 5):
        want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(pos[h], boxes[h]))
        assert int(counts[h]) == want["id"].shape[0] > 20 and int(rough[h]) == want_rough
        H.assert_cols_equal(ctx.fetch(h), want, what=f"halo {h}: ")


@pytest.mark.parametrize("cells", ["staged", "mono", "index"])
def test_halo_cell_index_window_edges(ctx, cells_mode, cells):
    """the cell-index kernel's second filter decides most pairs without the exact arithmetic: "certainly inside the
    rough window" and "certainly outside the fine window".  Aim particles at every bound of both windows -- the
    rough one in the un-rotated frame, the fine one in the rotated frame (directions built there and carried back
    with R^-1, so float rounding scatters them to both sides of the bound) -- for 9 halos in one pass"""
    rng = np.random.default_rng(23)
    specs = [((150, 20, 12), 40.0), ((30, 170, 100), 25.0), ((100, 100, 100), 90.0), ((10, 10, 250), 60.0),
             ((200, 120, 40), 15.0), ((60, 60, 20), 33.0), ((250, 30, 160), 70.0), ((90, 200, 30), 50.0), ((140, 140, 140), 45.0)]
    infos = [cc.halo_setup(p, b) for p, b in specs]
    parts = [H.shell_fixture(500_000, seed=61)]
    conv = 3.14159265 / 180.0 / 3600.0
    for k, info in enumerate(infos):
        hal = info.halo
        parts.append(H.edge_fixture(hal.theta_rough, hal.phi_rough, seed=100 + k, per_bound=48))
        # fine window: (theta', phi') on each bound +- a few ulps, mid-window in the other angle
        Rinv = np.array(list(info.R_inv), dtype=np.float64).reshape(3, 3)
        tc, pc = np.array(hal.theta_cut, dtype=np.float64), np.array(hal.phi_cut, dtype=np.float64)
        pts = []
        for which, bound in (("t", tc[0]), ("t", tc[1]), ("p", pc[0]), ("p", pc[1])):
            for _ in range(48):
                jit = float(rng.integers(-3, 4)) * float(np.spacing(np.float32(abs(bound) + 1.0)))
                t = (bound + jit if which == "t" else tc.mean() + rng.uniform(-0.45, 0.45) * (tc[1] - tc[0])) * conv
                f = (bound + jit if which == "p" else pc.mean() + rng.uniform(-0.45, 0.45) * (pc[1] - pc[0])) * conv
                r = rng.uniform(60, 400)
                pts.append(Rinv @ (r * np.array([np.sin(t) * np.cos(f), np.sin(t) * np.sin(f), np.cos(t)])))
        pts = np.array(pts, dtype=np.float32)
        e = H.slice_cols(H.shell_fixture(pts.shape[0], seed=200 + k), 0, pts.shape[0])
        e["x"], e["y"], e["z"] = pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy()
        parts.append(e)
    cols = H.concat_cols(*parts)
    cols["id"] = np.arange(cols["x"].shape[0], dtype=np.int64)
    ctx.upload(cols)
    ctx.cutout_halo([i.halo for i in infos])
    counts, rough = ctx.sync()
    near = 0
    for h, (p, b) in enumerate(specs):
        want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(p, b))
        assert int(counts[h]) == want["id"].shape[0] and int(rough[h]) == want_rough, (h, p, b)
        H.assert_cols_equal(ctx.fetch(h), want, what=f"halo {p} {b}: ")
        near += int(np.isin(want["id"], np.arange(500_000, cols["x"].shape[0])).sum())
    assert near > 200  # edge particles did land inside


def test_halo_large_segment(ctx):
    """a 10-degree box: tens of thousands of survivors in one halo -- beyond the single-CTA kernel (K3_SORT_MAX = 8192)
    and beyond one chunk of the general ordering kernel (K3_CHUNK = 16384: three chunks, the last one partial, merged
    by k3_place); a 5.7-degree box in the same pass fills most of one chunk (more than 8 of a thread's 16 key
    registers); a small halo takes a handful of records"""
    cols = H.shell_fixture(2_000_000, seed=77)
    specs = [((150, 120, 90), 600.0), ((150, 20, 12), 30.0), ((60, 200, 80), 340.0)]
    ctx.upload(cols)
    ctx.cutout_halo([cc.halo_setup(p, b).halo for p, b in specs])
    counts, rough = ctx.sync()
    assert int(counts[0]) > 2 * 16384 and 16384 >= int(counts[2]) > 8192 > int(counts[1]) > 0
    for h, (p, b) in enumerate(specs):
        want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(p, b))
        assert int(rough[h]) == want_rough
        H.assert_cols_equal(ctx.fetch(h), want, what=f"halo {p} {b}: ")


def test_halo_huge_segment_regrows_and_merges_hundreds_of_chunks(ctx):
    """a 60-degree box: more than 2^20 survivors in ONE halo -- beyond the initial record capacity (the pass is re-run
    at the exact size) and dozens of 16384-record chunks, each ordered by one CTA and merged by binary searches"""
    cols = H.shell_fixture(5_000_000, seed=91)
    p, b = (100, 100, 100), 3600.0
    fresh = cc.Context(0)  # a context whose buffers still have their initial sizes
    try:
        fresh.upload(cols)
        fresh.cutout_halo([cc.halo_setup(p, b).halo])
        counts, rough = fresh.sync()
        want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(p, b))
        assert int(counts[0]) == want["id"].shape[0] > (1 << 20) and int(rough[0]) == want_rough
        H.assert_cols_equal(fresh.fetch(0), want, what="huge segment: ")
    finally:
        fresh.close()


def test_halo_record_overflow_after_earlier_cutouts(ctx):
    """the record buffers overflow on a context that has ALREADY cut other halo sets: the overflowing pass sees the
    earlier cutouts' records in the not-yet-overwritten slots and must neither follow them nor fault; the automatic
    re-run at the exact size returns the reference's rows (found at 125 M particles x 250 halos, profiles/r02)"""
    cols = H.shell_fixture(4_000_000, seed=92)
    rng = np.random.default_rng(93)
    fresh = cc.Context(0)
    try:
        fresh.upload(cols)
        for box in (200.0, 500.0):  # two passes that fit the initial 2^20 record slots, 12 halos each
            pos = rng.uniform(40, 250, (12, 3)).astype(np.float32)
            fresh.cutout_halo([cc.halo_setup(p, box).halo for p in pos])
            counts, _ = fresh.sync()
            assert 1000 < int(counts.sum()) < (1 << 20)
        pos = rng.uniform(60, 200, (12, 3)).astype(np.float32)
        fresh.cutout_halo([cc.halo_setup(p, 2400.0).halo for p in pos])  # well beyond 2^20 records in total
        counts, rough = fresh.sync()
        assert int(counts.sum()) > (1 << 20)
        everything, off = fresh.fetch_all(counts)
        for h in (0, 5, 11):
            want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(pos[h], 2400.0))
            assert int(rough[h]) == want_rough
            H.assert_cols_equal({c: v[off[h]:off[h + 1]] for c, v in everything.items()}, want, what=f"overflow halo {h}: ")
    finally:
        fresh.close()


def test_halo_empty_inputs(ctx):
    info = cc.halo_setup((150, 20, 12), 10.0)
    ctx.upload(H.slice_cols(H.shell_fixture(10, seed=1), 0, 0))
    ctx.cutout_halo([info.halo])
    counts, rough = ctx.sync()
    assert int(counts[0]) == 0 and int(rough[0]) == 0
    ctx.upload(H.shell_fixture(1000, seed=1))
    ctx.cutout_halo([])
    counts, rough = ctx.sync()
    assert counts.shape[0] == 0


def test_exchange_counts_single_rank(ctx):
    """count exchange enqueued behind the cutout (no communicator = one rank): counts and zero offsets, for both
    use cases, whether or not cc_cutout_sync ran first"""
    cols = H.shell_fixture(500_000, seed=17)
    theta_cut, phi_cut = cc.cli_bounds(60, 10), cc.cli_bounds(30, 10)
    want = H.oracle_cutout_const(cols, theta_cut, phi_cut)["id"].shape[0]
    ctx.upload(cols)
    ctx.cutout_const(theta_cut, phi_cut)
    allc, allr, off = ctx.exchange_counts()
    assert allc.shape == (1, 1) and int(allc[0, 0]) == want and int(off[0]) == 0 and int(allr[0, 0]) == 0
    assert int(ctx.sync()[0][0]) == want
    infos = [cc.halo_setup(p, b) for p, b in HALOS[:4]]
    ctx.cutout_halo([i.halo for i in infos])
    counts, rough = ctx.sync()
    allc, allr, off = ctx.exchange_counts()
    assert np.array_equal(allc[0], counts) and np.array_equal(allr[0], rough) and not off.any()
    allc, allr = allc.copy(), allr.copy()  # the binding reuses its result arrays
    ctx.cutout_halo([i.halo for i in infos])
    allc2, allr2, _ = ctx.exchange_counts()
    assert np.array_equal(allc2, allc) and np.array_equal(allr2, allr)
    H.assert_cols_equal(ctx.fetch(1), H.oracle_cutout_halo(cols, H.oracle_halo_setup(*HALOS[1]))[0])


# ------------------------------------------------------------------------------------------------ streamed
def _pinned(cols):
    out, addrs = {}, []
    for k in cc.IN_COLS:
        a, addr = cc.pinned_empty(cols[k].shape[0], cols[k].dtype)
        a[:] = cols[k]
        out[k] = a
        addrs.append(addr)
    return out, addrs


def test_streamed_equals_oracle(ctx):
    """pinned host -> HBM double-buffered path, several chunks (16 Mi particles per chunk) + a ragged tail"""
    n = (16 << 20) * 2 + 12_345
    cols = cc.synth_host(n, seed=77, kind=0, box=300.0)
    pin, addrs = _pinned(cols)
    try:
        theta_cut, phi_cut = cc.cli_bounds(87, 2), cc.cli_bounds(2, 2)
        ctx.cutout_const_streamed(pin, theta_cut, phi_cut)
        ctx.sync()
        H.assert_cols_equal(ctx.fetch(0), H.oracle_cutout_const(cols, theta_cut, phi_cut), what="streamed const: ")
        info = cc.halo_setup((150, 20, 12), 60.0)
        limit = cc.ref_scatter_kept(n, 1)
        ctx.cutout_halo_streamed(pin, [info.halo], n_limit_global=limit)
        counts, rough = ctx.sync()
        want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup((150, 20, 12), 60.0))
        assert int(rough[0]) == want_rough
        H.assert_cols_equal(ctx.fetch(0), want, what="streamed halo: ")
        # nine halos over all three chunks: the cell-index kernel with a non-zero chunk offset (two halos checked: the
        # oracle sorts all 33 M particles per halo)
        rng = np.random.default_rng(11)
        pos9 = np.vstack([[150, 20, 12], rng.uniform(40, 250, (8, 3))]).astype(np.float32)
        ctx.cutout_halo_streamed(pin, [cc.halo_setup(p, 60.0).halo for p in pos9], n_limit_global=limit)
        counts, rough = ctx.sync()
        for h in (0, 5):
            want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(pos9[h], 60.0))
            assert int(counts[h]) == want["id"].shape[0] and int(rough[h]) == want_rough
            H.assert_cols_equal(ctx.fetch(h), want, what=f"streamed, 9 halos, halo {h}: ")
        # a dozen halos (cell-index kernel) over the first 3M particles, a global index base, everything fetched at
        # once; the survivors' vx..replication are picked on the host from the pinned columns -- and, as a
        # cross-check, by the GPU (CC_DEVICE_GATHER=1)
        n2 = 3_000_000
        sub, pin2 = H.slice_cols(cols, 0, n2), {k: v[:n2] for k, v in pin.items()}
        rng = np.random.default_rng(3)
        pos = rng.uniform(40, 250, (12, 3)).astype(np.float32)
        halos = [cc.halo_setup(p, 90.0).halo for p in pos]
        wants = [H.oracle_cutout_halo(sub, H.oracle_halo_setup(p, 90.0)) for p in pos]
        base = 7_000_000_000
        for device_gather in (False, True):
            if device_gather:
                os.environ["CC_DEVICE_GATHER"] = "1"
            try:
                ctx.cutout_halo_streamed(pin2, halos, n_limit_global=base + cc.ref_scatter_kept(n2, 1), index_base=base)
                counts, rough = ctx.sync()
                everything, off = ctx.fetch_all(counts, names=cc.OUT_COLS + ("index",))
            finally:
                os.environ.pop("CC_DEVICE_GATHER", None)
            assert int(counts.sum()) > 1000
            for h, (want, want_rough) in enumerate(wants):
                got = {k: v[off[h]:off[h + 1]] for k, v in everything.items()}
                assert int(rough[h]) == want_rough
                H.assert_cols_equal(got, want, what=f"streamed halo {h} (device_gather={device_gather}): ")
                assert np.array_equal(got["index"] - base, got["id"])
    finally:
        del pin
        for a in addrs:
            cc.pinned_free(a)


def test_streamed_rejects_pageable(ctx):
    cols = H.shell_fixture(1000, seed=1)
    with pytest.raises(cc.CutoutError, match="pageable"):
        ctx.cutout_const_streamed(cols, cc.cli_bounds(87, 2), cc.cli_bounds(2, 2))


# ------------------------------------------------------------------------------------------------ large
def test_large_resident_const_100m(ctx):
    """BASELINE-size step (100 M particles, device-generated): survivor count and columns equal the oracle run on
    the same (host-regenerated) particles; the tail drop of use case 2 equals the reference's (4 particles)"""
    n = 100_000_000
    ctx.synth(n, seed=20261019, kind=0, box=300.0)
    theta_cut, phi_cut = cc.cli_bounds(87, 2), cc.cli_bounds(2, 2)
    ctx.cutout_const(theta_cut, phi_cut)
    counts, _ = ctx.sync()
    got = ctx.fetch(0, names=cc.OUT_COLS + ("index",))
    host = cc.synth_host(n, seed=20261019, kind=0, box=300.0)
    want = H.oracle_cutout_const(host, theta_cut, phi_cut)
    assert int(counts[0]) == want["id"].shape[0] > 100_000
    H.assert_cols_equal(got, want, what="100M const: ")
    info = cc.halo_setup((150, 20, 12), 10.0)
    ctx.cutout_halo([info.halo], n_limit_global=cc.ref_scatter_kept(n, 1))
    counts, rough = ctx.sync()
    got = ctx.fetch(0, names=cc.OUT_COLS + ("index",))
    assert ctx.stats()["scanned"] == n - 4
    # oracle on the particles that can matter: everything whose exact theta_u lies in the rough annulus
    th_u, _ = H.oracle_theta_phi(host["x"], host["y"], host["z"], guards=True)
    sel = np.nonzero((th_u >= info.halo.theta_rough[0]) & (th_u < info.halo.theta_rough[1]))[0]
    sel = sel[sel < n - 4]
    sub = {k: np.ascontiguousarray(host[k][sel]) for k in cc.IN_COLS}
    want, want_rough = H.oracle_cutout_halo(sub, H.oracle_halo_setup((150, 20, 12), 10.0))
    assert int(rough[0]) == want_rough
    H.assert_cols_equal({k: got[k] for k in cc.OUT_COLS}, want, what="100M halo: ")


def test_device_cols_and_attach_with_torch(ctx):
    """the device-pointer side of the C ABI: attach torch-owned device columns, read the survivor columns straight
    from HBM (cc_cutout_device_cols) -- no host copies made by the library"""
    import ctypes as C
    import torch
    cols = H.shell_fixture(400_000, seed=55)
    keep = {k: torch.from_numpy(cols[k]).cuda() for k in cc.IN_COLS}
    torch.cuda.synchronize()
    ctx.attach(400_000, {k: t.data_ptr() for k, t in keep.items()})
    theta_cut, phi_cut = cc.cli_bounds(60, 10), cc.cli_bounds(30, 10)
    ctx.cutout_const(theta_cut, phi_cut)
    counts, _ = ctx.sync()
    want = H.oracle_cutout_const(cols, theta_cut, phi_cut)
    dev = cc.ColsOut()
    row0, cnt = C.c_int64(-1), C.c_int64(-1)
    cc._check(cc.lib().cc_cutout_device_cols(ctx._h, 0, C.byref(dev), C.byref(row0), C.byref(cnt)))
    assert cnt.value == int(counts[0]) == want["id"].shape[0] and row0.value == 0
    try:
        from cuda import cudart  # cuda-python: a plain cudaMemcpy on the raw device pointers
    except Exception:
        pytest.skip("cuda-python not importable")
    for name, dt in (("theta", np.float32), ("id", np.int64), ("replication", np.int32)):
        host = np.empty(cnt.value, dtype=dt)
        (err,) = cudart.cudaMemcpy(host.ctypes.data, getattr(dev, name), host.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        assert int(err) == 0
        assert host.tobytes() == want[name].tobytes(), name


def _hbm_free_gb():
    try:
        import torch
        return torch.cuda.mem_get_info(0)[0] / 2 ** 30
    except Exception:
        return 1e9


def test_full_size_oracle_parity_1b(ctx):
    """BASELINE north-star size: ONE cutout of a 1 B-particle resident step on one GPU against the oracle run over
    EVERY particle of the step -- the step is regenerated on the host in twenty 50 M-particle index slices
    (cc_synth_host == the device generator, test_synth_device_equals_host), the oracle (no pre-selection, no
    pre-filter: the reference's loop, processLC.cpp:276-310) cuts each slice, the slices are concatenated in index
    order.  Survivor ids, order and all 12 columns must be byte-identical; a particle the device pre-filter wrongly
    rejected would show up as a missing row.  Then index sharding: eight 125 M shards concatenate to the same rows."""
    import concurrent.futures as cf
    import hashlib
    n, nslice = 1_000_000_000, 20
    if _hbm_free_gb() < 70:
        pytest.skip("needs ~60 GB of free HBM")
    seed = 4242
    theta_cut, phi_cut = cc.cli_bounds(87, 2), cc.cli_bounds(2, 2)
    ctx.release()
    ctx.synth(n, seed=seed, kind=0, box=300.0)
    ctx.cutout_const(theta_cut, phi_cut)
    counts, _ = ctx.sync()
    k = int(counts[0])
    assert 1_000_000 < k < 3_000_000
    got = ctx.fetch(0, names=cc.OUT_COLS + ("index",))
    assert np.array_equal(got["index"], got["id"])
    per = n // nslice

    def one(s):  # ctypes calls release the GIL: slices run on several host cores
        host = cc.synth_host(per, seed, index_base=s * per, kind=0, box=300.0)
        return H.oracle_cutout_const(host, theta_cut, phi_cut)

    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        parts = list(ex.map(one, range(nslice)))
    want = {c: np.concatenate([p[c] for p in parts]) for c in cc.OUT_COLS}
    assert want["id"].shape[0] == k
    H.assert_cols_equal(got, want, what="1B const vs oracle on every particle: ")
    full_sig = hashlib.sha256(got["id"].tobytes()).hexdigest()
    ctx.release()
    h = hashlib.sha256()
    total = 0
    for r in range(8):
        lo, hi = cc.shard_range(n, 8, r)
        ctx.synth(hi - lo, seed=seed, index_base=lo, kind=0, box=300.0)
        ctx.cutout_const(theta_cut, phi_cut)
        kk = int(ctx.sync()[0][0])
        part = ctx.fetch(0, names=("id", "index"))
        assert np.array_equal(part["id"], part["index"])
        h.update(part["id"].tobytes())
        total += kk
    assert total == k and h.hexdigest() == full_sig
    ctx.release()


def test_many_halos_at_size_sample_parity(ctx):
    """BASELINE config 5 at size: 1000 halos (four boxLength groups, as four processLC-style passes) over a 125 M-particle
    resident shard; a sample of 20 halos is compared with the oracle run on everything inside a loose float64 annulus
    around the halo's rough theta window (selected with numpy, independently of the device pre-filter and cell index)."""
    n = 125_000_000
    if _hbm_free_gb() < 20:
        pytest.skip("needs ~12 GB of free HBM")
    seed = 777
    ctx.release()
    ctx.synth(n, seed=seed, kind=0, box=300.0)
    host = cc.synth_host(n, seed, kind=0, box=300.0)
    z64 = host["z"].astype(np.float64)
    r = np.sqrt(host["x"].astype(np.float64) ** 2 + host["y"].astype(np.float64) ** 2 + z64 ** 2)
    th64 = np.degrees(np.arccos(np.clip(z64 / r, -1, 1))) * 3600.0
    del z64, r
    rng = np.random.default_rng(1000)
    pos = rng.uniform(30.0, 280.0, (1000, 3)).astype(np.float32)
    limit = cc.ref_scatter_kept(n, 1)
    checked = 0
    per_group, all_halos = [], []
    ctx.set_option("index", "never")  # the four passes: the cell walkers (every pass streams the 125 M particles)
    for g, box in enumerate((5.0, 10.0, 20.0, 30.0)):
        hp = pos[g * 250:(g + 1) * 250]
        infos = [cc.halo_setup(p, box) for p in hp]
        all_halos += [i.halo for i in infos]
        ctx.cutout_halo([i.halo for i in infos], n_limit_global=limit)
        counts, rough = ctx.sync()
        assert ctx.stats_ex()["nan_theta"] == 0 and ctx.stats_ex()["tail_dropped"] == n - limit
        everything, off = ctx.fetch_all(counts)
        per_group.append((counts.copy(), rough.copy(), everything))
        for h in (3, 77, 128, 201, 249):
            hh = infos[h].halo
            sel = np.nonzero((th64 >= hh.theta_rough[0] - 30.0) & (th64 < hh.theta_rough[1] + 30.0))[0]
            sel = sel[sel < limit]
            sub = {c: np.ascontiguousarray(host[c][sel]) for c in cc.IN_COLS}
            want, want_rough = H.oracle_cutout_halo(sub, H.oracle_halo_setup(hp[h], box))
            assert int(rough[h]) == want_rough and int(counts[h]) == want["id"].shape[0]
            H.assert_cols_equal({c: v[off[h]:off[h + 1]] for c, v in everything.items()}, want, what=f"box {box} halo {h}: ")
            checked += 1
    assert checked == 20
    # the same 1000 (halo, box length) cutouts in ONE pass (host/processLC.cpp processLC_boxes; bench cfg5_shared): every
    # halo's counts and rows are those of the four separate passes, bit for bit
    for index in ("never", "always"):  # one pass: cell walker, then the particle index (built inside this cutout)
        ctx.set_option("index", index)
        ctx.cutout_halo(all_halos, n_limit_global=limit)
        counts, rough = ctx.sync()
        st = ctx.stats_ex()
        assert st["index_resident"] == (index == "always") and (st["index_build_ms"] > 0) == (index == "always")
        assert (st["scan_kernels"] == "k2p_query") == (index == "always")
        assert st["particles_read"] == limit if index == "never" else 0 < st["particles_read"] < limit // 4
    ctx.set_option("index", "auto")
    everything, _ = ctx.fetch_all(counts)
    assert np.array_equal(counts, np.concatenate([p[0] for p in per_group]))
    assert np.array_equal(rough, np.concatenate([p[1] for p in per_group]))
    for c in everything:
        assert np.array_equal(everything[c].view(np.uint8), np.concatenate([p[2][c] for p in per_group]).view(np.uint8)), c
    ctx.release()


def test_nan_theta_is_counted_once_per_particle(ctx):
    """positions whose un-rotated theta is NaN (NaN coordinate, z = +-Inf, the origin, x = y = 0 with z^2 underflowing)
    are outside the parity claim for use case 2 (the reference's stable_sort / lower_bound on a NaN key is
    unspecified); the device counts them, once per particle whatever the number of halos, in both scan kernels"""
    cols = H.shell_fixture(300_000, seed=77)
    bad = np.array([5, 70_001, 123_456, 200_000, 299_999])
    cols["x"][bad[0]] = np.nan
    cols["z"][bad[1]] = np.inf
    cols["x"][bad[2]] = cols["y"][bad[2]] = cols["z"][bad[2]] = 0.0
    cols["x"][bad[3]] = cols["y"][bad[3]] = 0.0
    cols["z"][bad[3]] = np.float32(1e-30)
    cols["y"][bad[4]] = np.nan
    cols["z"][1000] = -np.inf        # also NaN (Inf/Inf)
    cols["x"][2000] = np.inf         # theta = acos(0) = 90 deg: fine
    ctx.upload(cols)
    rng = np.random.default_rng(3)
    for nh in (1, 12):
        pos = rng.uniform(40, 250, (nh, 3)).astype(np.float32)
        ctx.cutout_halo([cc.halo_setup(p, 30.0).halo for p in pos])
        ctx.sync()
        assert ctx.stats_ex()["nan_theta"] == 6, (nh, ctx.stats_ex())
    ctx.cutout_const(cc.cli_bounds(87, 2), cc.cli_bounds(2, 2))
    ctx.sync()
    assert ctx.stats_ex()["nan_theta"] == 0  # use case 1 has no sort: NaN positions simply fail the window test


def test_payload_policy_and_gather_variants(ctx):
    """the interleaved gather payload is built lazily (never for a sparse cutout, inside the second dense cutout on a
    resident step) and every gather variant -- SoA columns, payload, payload + bulk stores -- returns the same bytes"""
    cols = H.shell_fixture(600_000, seed=91)
    dense_t, dense_p = cc.cli_bounds(60, 25), cc.cli_bounds(40, 35)
    want = H.oracle_cutout_const(cols, dense_t, dense_p)
    assert want["id"].shape[0] > 150_000
    info = cc.halo_setup((150, 20, 12), 60.0)
    want_h, _ = H.oracle_cutout_halo(cols, H.oracle_halo_setup((150, 20, 12), 60.0))
    try:
        ctx.set_option("payload", "auto")
        ctx.upload(cols)
        ctx.cutout_halo([info.halo])
        ctx.sync()
        assert not ctx.stats_ex()["payload_resident"]         # sparse: never
        H.assert_cols_equal(ctx.fetch(0), want_h, what="halo, SoA gather: ")
        ctx.cutout_const(dense_t, dense_p)
        ctx.sync()
        assert not ctx.stats_ex()["payload_resident"]         # first dense cutout: gathers from the columns
        H.assert_cols_equal(ctx.fetch(0), want, what="const, SoA gather: ")
        ctx.cutout_const(dense_t, dense_p)
        ctx.sync()
        st = ctx.stats_ex()
        assert st["payload_resident"] and st["payload_build_ms"] > 0   # second dense cutout: built inside this call
        H.assert_cols_equal(ctx.fetch(0), want, what="const, payload gather: ")
        for mode in ("bulk", "ldst", "auto"):
            ctx.set_option("gather", mode)
            ctx.cutout_const(dense_t, dense_p)
            ctx.sync()
            assert ctx.stats_ex()["payload_build_ms"] == 0
            H.assert_cols_equal(ctx.fetch(0), want, what=f"const, gather={mode}: ")
        ctx.cutout_halo([info.halo])
        ctx.sync()
        H.assert_cols_equal(ctx.fetch(0), want_h, what="halo, payload gather: ")
        ctx.set_option("payload", "never")
        ctx.upload(cols)
        for mode in ("bulk", "ldst"):
            ctx.set_option("gather", mode)
            ctx.cutout_const(dense_t, dense_p)
            ctx.sync()
            ctx.cutout_const(dense_t, dense_p)
            ctx.sync()
            assert not ctx.stats_ex()["payload_resident"]
            H.assert_cols_equal(ctx.fetch(0), want, what=f"const, payload=never gather={mode}: ")
        with pytest.raises(cc.CutoutError):
            ctx.set_option("gather", "sideways")
        with pytest.raises(cc.CutoutError):
            ctx.set_option("no-such-key", "x")
    finally:
        ctx.set_option("payload", "auto")
        ctx.set_option("gather", "auto")


def test_alternating_halo_sets_reuse_their_device_tables(ctx):
    """a batch job alternates between a few halo sets (one per boxLength): each keeps its device-resident parameters and
    cell index; results do not depend on the order of use, also after more sets than the cache holds"""
    cols = H.shell_fixture(300_000, seed=12)
    ctx.upload(cols)
    rng = np.random.default_rng(9)
    sets = []
    for s_ in range(11):  # the cache holds 8
        nh = 9 if s_ % 2 == 0 else 2
        pos = rng.uniform(40, 250, (nh, 3)).astype(np.float32)
        box = float(rng.uniform(20, 90))
        sets.append((pos, box, [cc.halo_setup(p, box).halo for p in pos]))
    wants = {}
    for rnd in range(2):
        for s_ in list(range(11)) + [0, 5, 0, 10, 3]:
            pos, box, halos = sets[s_]
            ctx.cutout_halo(halos)
            counts, rough = ctx.sync()
            everything, off = ctx.fetch_all(counts)
            for h in (0, len(halos) - 1):
                if (s_, h) not in wants:
                    wants[(s_, h)] = H.oracle_cutout_halo(cols, H.oracle_halo_setup(pos[h], box))
                w, wr = wants[(s_, h)]
                assert int(rough[h]) == wr
                H.assert_cols_equal({c: v[off[h]:off[h + 1]] for c, v in everything.items()}, w, what=f"set {s_} halo {h}: ")


def test_resident_step_slots(ctx):
    """a context keeps several lightcone steps resident (cc_step_select) and cuts any of them without another
    host -> device copy -- what lets consecutive processLC calls (one boxLength each) skip read + upload"""
    steps = [H.shell_fixture(200_000 + 1000 * i, seed=60 + i) for i in range(3)]
    tc, pc = cc.cli_bounds(80, 8), cc.cli_bounds(20, 8)
    info = cc.halo_setup((150, 20, 12), 90.0)
    fresh = cc.Context(0)
    try:
        free0, total = fresh.device_mem()
        assert 0 < free0 <= total
        for i, cols in enumerate(steps):
            fresh.select(i + 1)
            assert fresh.step_size() == -1          # an unused slot is empty
            fresh.upload(cols, index_base=1000 * i)
        moved = fresh.bytes_uploaded()
        assert moved == sum(c["x"].shape[0] for c in steps) * 44
        for rnd in range(2):
            for i in (2, 0, 1, 0):
                fresh.select(i + 1)
                assert fresh.step_size() == steps[i]["x"].shape[0]
                fresh.cutout_const(tc, pc)
                fresh.sync()
                got = fresh.fetch(0, names=cc.OUT_COLS + ("index",))
                want = H.oracle_cutout_const(steps[i], tc, pc)
                H.assert_cols_equal(got, want, what=f"slot {i + 1} const: ")
                assert np.array_equal(got["index"] - 1000 * i, got["id"])
                fresh.cutout_halo([info.halo])
                fresh.sync()
                want_h, _ = H.oracle_cutout_halo(steps[i], H.oracle_halo_setup((150, 20, 12), 90.0))
                H.assert_cols_equal(fresh.fetch(0), want_h, what=f"slot {i + 1} halo: ")
        assert fresh.bytes_uploaded() == moved      # nothing was uploaded again
        fresh.select(2)
        fresh.release()
        assert fresh.step_size() == -1
        fresh.select(1)
        assert fresh.step_size() == steps[0]["x"].shape[0]
        fresh.select(0)
        with pytest.raises(cc.CutoutError):
            fresh.cutout_const(tc, pc)              # slot 0 holds nothing
    finally:
        fresh.close()


def test_particle_index_policy_and_slots():
    """the particle index of a resident step (k2p_*): policy auto builds it inside the SECOND many-halo cutout on a
    resident step of >= 4 Mi particles, later many-halo cutouts read only the cells under the halos' windows, and the
    cutouts are the same bit for bit; cutouts of a few halos keep scanning; the index belongs to the step: another
    slot has none, coming back finds it, a new step in the slot drops it"""
    n = 5_000_000
    rng = np.random.default_rng(99)
    pos = rng.uniform(30.0, 280.0, (40, 3)).astype(np.float32)
    boxes = rng.uniform(5.0, 40.0, 40)
    halos = [cc.halo_setup(p, b).halo for p, b in zip(pos, boxes)]
    fresh = cc.Context(0)

    def cut(hl, limit=-1):
        fresh.cutout_halo(hl, n_limit_global=limit)
        counts, rough = fresh.sync()
        ev, _ = fresh.fetch_all(counts)
        return counts.copy(), rough.copy(), ev, fresh.stats_ex()

    def same(a, b):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        for c in a[2]:
            assert np.array_equal(a[2][c].view(np.uint8), b[2][c].view(np.uint8)), c

    try:
        fresh.select(1)
        fresh.synth(n, seed=5, kind=0, box=300.0)
        res = [cut(halos) for _ in range(3)]
        assert int(res[0][0].sum()) > 1000
        for k, r in enumerate(res):
            st = r[3]
            assert st["index_resident"] == (k >= 1) and (st["index_build_ms"] > 0) == (k == 1), (k, st)
            assert (st["scan_kernels"] == "k2p_query") == (k >= 1), (k, st)
            assert st["particles_read"] == n if k == 0 else 0 < st["particles_read"] < n // 4
            same(r, res[0])
        limit = n - 1_234_567  # the reference's tail drop applies to indexed rows too
        fresh.set_option("index", "never")
        want = cut(halos, limit)
        fresh.set_option("index", "auto")
        got = cut(halos, limit)
        assert got[3]["scan_kernels"] == "k2p_query" and want[3]["scan_kernels"] != "k2p_query"
        assert int(want[0].sum()) < int(res[0][0].sum())
        same(got, want)
        assert cut(halos[:2])[3]["scan_kernels"] == "k2_cutout_halo"  # below 8 halos: no cell index, plain scan
        fresh.select(2)
        fresh.synth(n, seed=6, kind=0, box=300.0)
        other = cut(halos)
        assert not other[3]["index_resident"] and other[3]["scan_kernels"] != "k2p_query"
        fresh.select(1)
        back = cut(halos)
        assert back[3]["index_resident"] and back[3]["index_build_ms"] == 0 and back[3]["scan_kernels"] == "k2p_query"
        same(back, res[0])
        fresh.release()
        fresh.synth(n, seed=7, kind=0, box=300.0)  # a new step in the slot: its index is gone
        assert not cut(halos)[3]["index_resident"]
    finally:
        fresh.close()


def test_small_cutout_kernel_sizes_and_timing_switch(ctx):
    """the single-CTA ordering + gather kernel (k3_small) orders by counting up to 1024 records and with the bitonic
    network beyond; it hands the counts to the host itself and leaves the device state clean for the next cutout.
    Cutouts of growing size, each run twice (the second run takes the kernel), several halos, then a size that makes it
    refuse (> 8192) and fall back; per-kernel timing is off unless asked for"""
    cols = H.shell_fixture(1_500_000, seed=64)
    ctx.upload(cols)
    pos = np.array([[150, 20, 12], [100, 100, 100], [30, 170, 100]], dtype=np.float32)
    seen = []
    for box in (40.0, 90.0, 130.0, 170.0, 260.0, 420.0):
        infos = [cc.halo_setup(p, box) for p in pos]
        wants = [H.oracle_cutout_halo(cols, H.oracle_halo_setup(p, box)) for p in pos]
        for rep in range(2):
            ctx.cutout_halo([i.halo for i in infos])
            counts, rough = ctx.sync()
            everything, off = ctx.fetch_all(counts)
            for h, (w, wr) in enumerate(wants):
                assert int(rough[h]) == wr
                H.assert_cols_equal({c: v[off[h]:off[h + 1]] for c, v in everything.items()}, w, what=f"box {box} rep {rep} halo {h}: ")
        seen.append(int(counts.sum()))
    assert seen[0] < 1024 < seen[2] and any(1024 < k <= 8192 for k in seen) and seen[-1] > 8192, seen
    assert ctx.last_scan_ms() < 0  # timing events are not recorded by default
    try:
        ctx.set_option("timing", "on")
        ctx.cutout_halo([cc.halo_setup(pos[0], 40.0).halo])
        ctx.sync()
        assert 0 < ctx.last_scan_ms() <= ctx.last_total_ms()
    finally:
        ctx.set_option("timing", "off")
