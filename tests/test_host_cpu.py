"""CPU-only checks of the product's host side: the C-ABI library loads and exports every declared symbol, the
host set-up functions reproduce the oracle (and through it the reference) bit for bit, the fdlibm restatement
the kernels execute equals the host glibc on every float, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import harness as H
from __graft_entry__ import REPO, load_package

cc = load_package()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "cosmo_cutout_b200.h")).read()
    declared = set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"cc_ctx"}
    assert declared == set(cc.SYMBOLS), declared ^ set(cc.SYMBOLS)
    L = cc.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert b"sm_100a" in L.cc_version()


def test_header_is_plain_c_and_the_integration_stub_links(tmp_path):
    """include/cosmo_cutout_b200.h compiles as C11 with -pedantic, and the call sequence of INTEGRATION.md section 3
    (tests/native/abi_stub.c) links against the product library; without a GPU the stub sees cc_create fail loudly"""
    exe = str(tmp_path / "abi_stub")
    libdir = os.path.join(REPO, "cosmo-cutout_b200", "lib")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(REPO, "include"),
                           os.path.join(REPO, "tests", "native", "abi_stub.c"), "-o", exe,
                           "-L" + libdir, "-lcosmo_cutout_b200", "-Wl,-rpath," + libdir])
    p = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert p.returncode == 0 and "theta bounds 306000.0 320400.0" in p.stdout, p.stdout
    if cc.lib().cc_device_count() == 0:
        assert "cc_create without a GPU: rc=" in p.stdout and "rc=0" not in p.stdout


def test_no_cpu_fallback_without_gpu():
    if cc.lib().cc_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cc.CutoutError, match="no CUDA device|no CPU fallback"):
        cc.Context(0)


def test_product_does_not_link_oracle():
    out = subprocess.check_output(["ldd", cc.LIB_PATH]).decode()
    assert "oracle" not in out and "cutout_ref" not in out
    for root, _, files in os.walk(os.path.join(REPO, "cosmo-cutout_b200")):
        if "build" in root or "__pycache__" in root:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py", "Makefile")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle/" not in txt and "cutout_oracle" not in txt, (root, f)


def test_halo_setup_equals_oracle():
    rng = np.random.default_rng(17)
    pts = [(150, 20, 12), (1, 0, 0), (0, 1, 0), (0, 0, 1), (5, 5, 5), (-3, 2, 1), (200, 1, 1), (10, 10, 250)]
    pts += [tuple(rng.uniform(-300, 300, 3)) for _ in range(300)]
    for p in pts:
        for box in (5.0, 10.0, 30.0, 120.0):
            a = cc.halo_setup(p, box)
            o = H.oracle_halo_setup(p, box)
            for nme in ("halo_r", "half_box_rad", "beta"):
                assert np.float32(getattr(a, nme)).tobytes() == np.float32(getattr(o, nme)).tobytes(), (p, box, nme)
            for nme in ("k", "K", "R", "R_inv", "corners_rot", "corners_sph"):
                assert bytes(getattr(a, nme)) == bytes(getattr(o, nme)), (p, box, nme)
            for nme in ("theta_cut", "phi_cut", "theta_rough", "phi_rough"):
                assert bytes(getattr(a.halo, nme)) == bytes(getattr(o, nme)), (p, box, nme)
            assert bytes(a.halo.R) == bytes(o.R)


def test_cli_bounds_and_tail_drop_equal_oracle():
    for c, h in ((87, 2), (2, 2), (90, 20), (0, 20), (45.5, 0.25), (89.9, 0.2), (3, 3)):
        assert cc.cli_bounds(c, h).tobytes() == H.oracle_cli_bounds(c, h).tobytes()
    o = H.oracle()
    for n_p in (0, 1, 2, 1000, 2 ** 24 - 1, 2 ** 24, 2 ** 24 + 1, 2 ** 24 + 3, 20_000_001, 33_554_433, 100_000_000,
                1_000_000_000):
        for ranks in (1, 2, 8):
            assert cc.ref_scatter_kept(n_p, ranks) == o.oracle_scatter_kept(n_p, ranks), (n_p, ranks)
    assert cc.ref_scatter_kept(100_000_000, 1) == 100_000_000 - 4


def test_restated_libm_equals_glibc_everywhere(tmp_path):
    """all 2^32 float inputs, ~15 s on 8 cores: the device code executes exactly this operation sequence"""
    exe = str(tmp_path / "check_libm")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-ffp-contract=off", "-fopenmp",
                           "-I" + os.path.join(REPO, "cosmo-cutout_b200", "csrc"), "-x", "c++",
                           os.path.join(REPO, "tests", "native", "check_libm_restate.cpp"), "-o", exe, "-lm"])
    out = subprocess.check_output([exe, "1"]).decode()
    assert "acosf mismatches=0 atanf mismatches=0 checked=4294967296" in out, out


def test_division_free_arcsec_scaling_is_exact_everywhere(tmp_path):
    """the device replaces `/ PI` by reciprocal-multiply + one fma correction: proven equal to the true fp64 division
    for every finite non-zero float (all 2^32 bit patterns, ~3 s)"""
    exe = str(tmp_path / "check_arcsec")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-ffp-contract=off", "-fopenmp", "-mfma",
                           os.path.join(REPO, "tests", "native", "check_arcsec_fast.cpp"), "-o", exe, "-lm"])
    out = subprocess.check_output([exe]).decode()
    assert "quotient_mismatches=0 float_mismatches=0" in out and "checked=4278190078" in out, out


def test_conservative_windows_hold_against_exact_arithmetic(tmp_path):
    """prefilter.h: the outer windows never reject what the exact test accepts and the inner window of the cell-index
    kernel's second filter never accepts what the exact rough test rejects -- 4000 windows x 600 points aimed at the
    bounds, each with the worst-case error of the device's approximations applied both ways"""
    exe = str(tmp_path / "check_windows")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-ffp-contract=off",
                           "-I" + os.path.join(REPO, "cosmo-cutout_b200", "csrc"), "-x", "c++",
                           os.path.join(REPO, "tests", "native", "check_window_filters.cpp"), "-o", exe, "-lm"])
    out = subprocess.check_output([exe]).decode()
    assert "violations=0" in out and "windows=4000" in out, out
    decided = int(out.split("inner_decided=")[1].split()[0])
    assert decided > 100_000, out  # the inner window is not vacuous


@pytest.fixture(scope="module")
def gio_dump(tmp_path_factory):
    """the product's column reader compiled into a stand-alone dump tool (pinned allocation replaced by malloc)"""
    exe = str(tmp_path_factory.mktemp("gio") / "gio_dump")
    host = os.path.join(REPO, "cosmo-cutout_b200", "host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + host, "-I" + os.path.join(REPO, "include"),
                           os.path.join(REPO, "tests", "native", "gio_dump.cpp"), os.path.join(host, "column_reader.cpp"),
                           "-o", exe, "-lpthread"])
    return exe


GIO_VARIANTS = [dict(nranks=1), dict(nranks=4), dict(nranks=5, block_headers=True), dict(nranks=3, big_endian=True),
                dict(nranks=6, partitions=3), dict(nranks=7, partitions=2, big_endian=True, block_headers=True),
                dict(nranks=4, old_headers=True)]


@pytest.mark.parametrize("variant", GIO_VARIANTS, ids=lambda v: "-".join(f"{k}{int(x)}" for k, x in v.items()))
def test_column_reader_reads_genericio_format(gio_dump, tmp_path, variant):
    """host/column_reader.cpp, GenericIO back-end: files written by an independent writer (harness.write_gio_step)
    -- one file or partitioned, either byte order, with and without block headers, old header sizes -- come back as
    the same columns, in global rank order, as the flat column store gives"""
    cols = H.shell_fixture(20_011, seed=8)
    header = H.write_gio_step(str(tmp_path), 487, cols, seed=3, **variant)
    for pos_only in (False, True):
        out = tmp_path / ("dump_p" if pos_only else "dump")
        out.mkdir()
        msg = subprocess.check_output([gio_dump, header, str(out)] + (["posOnly"] if pos_only else [])).decode()
        assert f"n={cols['x'].shape[0]}" in msg, msg
        names = ("x", "y", "z", "a", "id") if pos_only else cc.IN_COLS
        for k in names:
            got = np.fromfile(str(out / (k + ".bin")), dtype=cols[k].dtype)
            assert got.tobytes() == np.ascontiguousarray(cols[k]).tobytes(), k


def test_column_reader_flat_store_and_errors(gio_dump, tmp_path):
    cols = H.shell_fixture(1000, seed=9)
    header = H.write_flat_step(str(tmp_path), 475, cols)
    out = tmp_path / "dump"
    out.mkdir()
    assert "n=1000" in subprocess.check_output([gio_dump, header, str(out)]).decode()
    assert np.fromfile(str(out / "id.bin"), dtype=np.int64).tobytes() == cols["id"].tobytes()
    # a GenericIO file with a variable of the wrong width, and a truncated one, are refused with a message
    bad = dict(cols)
    bad["id"] = cols["id"].astype(np.int32)
    h2 = H.write_gio_step(str(tmp_path / "w"), 487, bad, nranks=2)
    blob = open(h2, "rb").read()
    # patch the declared size of "id" from 8 to 4 (the harness writes i8 data; declare it 4 bytes wide)
    at = blob.index(b"id".ljust(256, b"\0")) + 256 + 8
    open(h2, "wb").write(blob[:at] + np.array([4], dtype="<u8").tobytes() + blob[at + 8:])
    p = subprocess.run([gio_dump, h2, str(out)], stdout=subprocess.PIPE, env=dict(os.environ, CC_GIO_CRC="off"))
    assert p.returncode != 0 and b"4-byte elements" in p.stdout  # (the patched header no longer matches its CRC)
    h3 = H.write_gio_step(str(tmp_path / "t"), 487, cols, nranks=3)
    open(h3, "r+b").truncate(os.path.getsize(h3) - 5000)
    p = subprocess.run([gio_dump, h3, str(out)], stdout=subprocess.PIPE)
    assert p.returncode != 0 and b"short read" in p.stdout


def test_column_reader_verifies_crc64_and_refuses_compressed_blocks(gio_dump, tmp_path):
    """GenericIO closes its header and every variable block with a CRC64 (CRC-64/XZ: the catalogue's check value pins
    the test writer's implementation); the reader verifies both, names what is corrupt, and can be told to go on.  A
    file written through a compression filter is refused when its header is parsed, with a way out."""
    assert (~H.crc64_xz(b"123456789")) & 0xFFFFFFFFFFFFFFFF == 0x995DC9BBDF1939FA
    assert H.crc64_xz(b"123456789" + H.crc64_closing_bytes(b"123456789")) == 0
    cols = H.shell_fixture(5_003, seed=10)
    out = tmp_path / "dump"
    out.mkdir()

    def run(header, **env):
        e = dict(os.environ)
        e.update(env)
        return subprocess.run([gio_dump, header, str(out)], stdout=subprocess.PIPE, env=e)

    # a flipped bit in the data of one variable
    h = H.write_gio_step(str(tmp_path / "a"), 487, cols, nranks=3)
    blob = bytearray(open(h, "rb").read())
    blob[len(blob) // 2] ^= 0x10
    open(h, "wb").write(bytes(blob))
    p = run(h)
    assert p.returncode != 0 and b"CRC64 mismatch in variable" in p.stdout, p.stdout
    p = run(h, CC_GIO_CRC="warn")
    assert p.returncode == 0 and b"WARNING" in p.stdout and b"n=5003" in p.stdout
    p = run(h, CC_GIO_CRC="off")
    assert p.returncode == 0 and b"WARNING" not in p.stdout
    # a wrong CRC behind one variable (data intact), in a partitioned file with block headers
    h = H.write_gio_step(str(tmp_path / "b"), 487, cols, nranks=4, partitions=2, block_headers=True, bad_crc="vy")
    p = run(h)
    assert p.returncode != 0 and b"CRC64 mismatch in variable vy" in p.stdout, p.stdout
    # a damaged header
    h = H.write_gio_step(str(tmp_path / "c"), 487, cols, nranks=2)
    blob = bytearray(open(h, "rb").read())
    blob[8 + 8 * 13] ^= 0x01  # PhysOrigin: not a field the reader uses, but it is under the header's CRC
    open(h, "wb").write(bytes(blob))
    p = run(h)
    assert p.returncode != 0 and b"CRC64 mismatch in the header" in p.stdout, p.stdout
    # compressed blocks
    h = H.write_gio_step(str(tmp_path / "d"), 487, cols, nranks=2, block_headers=True, filter_name="BLOSC")
    p = run(h)
    assert p.returncode != 0 and b"'BLOSC' filter" in p.stdout and b"GENERICIO_COMPRESS=0" in p.stdout, p.stdout


def test_host_build_of_device_arithmetic_equals_oracle():
    cols = H.concat_cols(H.shell_fixture(200_000, seed=3), H.shell_fixture(200_000, seed=4, octant=False),
                         H.edge_fixture((306000, 320400), (0, 14400)))
    x, y, z = cols["x"], cols["y"], cols["z"]
    th1, ph1 = H.oracle_theta_phi(x, y, z, guards=False)
    th2, ph2 = H.oracle_theta_phi(x, y, z, guards=True)

    def same(a, b):
        return np.array_equal(a.view(np.uint32)[~np.isnan(a)], b.view(np.uint32)[~np.isnan(b)]) and \
            np.array_equal(np.isnan(a), np.isnan(b))
    assert same(cc.debug_eval_host(2, x, y, z), th1)
    assert same(cc.debug_eval_host(3, x, y), ph1)
    assert same(cc.debug_eval_host(4, x, y), ph2)
    a = np.ascontiguousarray(cols["a"][:2000])
    want = np.array([H.oracle().oracle_aToZ(float(v)) for v in a], dtype=np.float32)
    assert cc.debug_eval_host(5, a).tobytes() == want.tobytes()


def test_synth_host_is_deterministic_and_sliceable():
    a = cc.synth_host(10_000, seed=42, index_base=0, kind=0, box=300.0)
    b = cc.synth_host(4_000, seed=42, index_base=3_000, kind=0, box=300.0)
    for k in cc.IN_COLS:
        assert a[k][3_000:7_000].tobytes() == b[k].tobytes(), k
    assert (a["x"] > 0).all() and (a["y"] > 0).all() and (a["z"] > 0).all() and (a["x"] <= 300).all()
    c = cc.synth_host(10_000, seed=42, index_base=0, kind=1, box=300.0)
    assert (c["x"] < 0).any() and (np.abs(c["x"]) < 300).all()
    assert np.array_equal(a["id"], np.arange(10_000))
    assert a["rotation"].min() >= 0 and a["rotation"].max() <= 5 and a["replication"].max() < 2 ** 20


def test_entry_points_reject_bad_arguments_without_touching_a_gpu():
    """error conventions of the C ABI: negative codes + a message, no crash, nothing thrown"""
    L = cc.lib()
    null = C.c_void_p()
    two = (C.c_float * 2)(1.0, 2.0)
    for call in (lambda: L.cc_cutout_const(null, two, two),
                 lambda: L.cc_cutout_halo(null, 1, null, 0, -1),
                 lambda: L.cc_cutout_sync(null, null, null),
                 lambda: L.cc_step_upload(null, 10, null, 0),
                 lambda: L.cc_exchange_counts(null, null, null, null),
                 lambda: L.cc_cutout_fetch(null, 0, 0, 0, null),
                 lambda: L.cc_comm_peer_attach(null, null),
                 lambda: L.cc_halo_setup(null, C.c_float(10.0), null)):
        assert call() < 0
    assert L.cc_last_error() is not None
    assert L.cc_step_size(null) == -1 and L.cc_launch_count(null) == 0
    out = C.c_void_p()
    assert L.cc_create(-1, C.byref(out)) < 0 and not out.value


def test_binding_fails_loudly_when_the_library_is_missing(monkeypatch):
    import importlib
    monkeypatch.setattr(cc, "_lib", None)
    monkeypatch.setattr(cc, "LIB_PATH", os.path.join(REPO, "cosmo-cutout_b200", "lib", "does_not_exist.so"))
    with pytest.raises(cc.CutoutError, match="missing"):
        cc.lib()


def test_numpy_replica_of_the_synthetic_lightcone_equals_the_library():
    """bench.py's reference arm makes its columns with tests/harness.py:synth_numpy so that it never loads the product
    library; the replica must be the same generator, bit for bit, at any global index"""
    for kind, base in ((0, 0), (0, 987_654_321_012), (1, 3_999_999_000)):
        a = cc.synth_host(50_021, 20261019, index_base=base, kind=kind, box=300.0)
        b = H.synth_numpy(50_021, 20261019, index_base=base, kind=kind, box=300.0)
        for k in cc.IN_COLS:
            assert a[k].dtype == b[k].dtype and a[k].tobytes() == b[k].tobytes(), (kind, base, k)
