"""Shared test plumbing: ctypes bindings of the oracle (oracle/liboracle.so), of the real reference
(oracle/_ref/libcutout_ref.so, when built) and synthetic lightcone fixtures.

Nothing here is product code.  The product is reached only through cosmo_cutout_b200 (C-ABI).
"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

F32_COLS = ("x", "y", "z", "vx", "vy", "vz", "a")
IN_COLS = F32_COLS + ("id", "rotation", "replication")
OUT_COLS = ("x", "y", "z", "vx", "vy", "vz", "redshift", "theta", "phi", "id", "rotation", "replication")
OUT_COLS_POSONLY = ("x", "y", "z", "redshift", "theta", "phi", "id")
COL_DTYPES = dict(x=np.float32, y=np.float32, z=np.float32, vx=np.float32, vy=np.float32, vz=np.float32,
                  a=np.float32, redshift=np.float32, theta=np.float32, phi=np.float32, id=np.int64,
                  rotation=np.int32, replication=np.int32)


# --------------------------------------------------------------------------------------------
# fixtures
# --------------------------------------------------------------------------------------------
def shell_fixture(n, seed=20261019, rmin=100.0, rmax=300.0, octant=True):
    """Uniform-in-solid-angle shell, the generator of SURVEY.md App. B.2 (numpy PCG64).
    octant=False spreads phi over (-pi, pi) and cos(theta) over (-1, 1) for use-case-2 aliasing tests."""
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    ph = rng.random(n) * np.pi / 2
    ct = u
    r = rng.uniform(rmin, rmax, n)
    if not octant:
        ct = 2 * ct - 1
        ph = 4 * ph - np.pi
    st = np.sqrt(1 - ct * ct)
    cols = dict(
        x=(r * st * np.cos(ph)).astype(np.float32),
        y=(r * st * np.sin(ph)).astype(np.float32),
        z=(r * ct).astype(np.float32),
    )
    for name in ("vx", "vy", "vz"):
        cols[name] = rng.normal(0, 300, n).astype(np.float32)
    cols["a"] = rng.uniform(0.5, 1, n).astype(np.float32)
    cols["id"] = np.arange(n, dtype=np.int64)
    cols["rotation"] = rng.integers(0, 6, n).astype(np.int32)
    cols["replication"] = rng.integers(0, 2 ** 20, n).astype(np.int32)
    return cols


def synth_numpy(n, seed, index_base=0, kind=0, box=300.0):
    """numpy replica of the product's counter-based synthetic lightcone (kernels.cuh synth_one / cc_synth_host), so
    that the reference arm of bench.py can make its columns without loading the product library;
    tests/test_host_cpu.py asserts it equals cc_synth_host bit for bit."""
    g = np.arange(index_base, index_base + n, dtype=np.uint64)

    def mix(col):
        with np.errstate(over="ignore"):
            z = np.uint64(seed) + np.uint64(0x9e3779b97f4a7c15) * (g + np.uint64(1)) + \
                np.uint64(0xd1b54a32d192ed03) * np.uint64(col + 1)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
            return z ^ (z >> np.uint64(31))

    def unit(col):  # (2k+1) * 2^-24, k in [0, 2^23): exact in float32
        k = (mix(col) >> np.uint64(41)).astype(np.uint32)
        return (np.uint32(2) * k + np.uint32(1)).astype(np.float32) * np.float32(5.9604644775390625e-08)

    f = np.float32
    u = [unit(c) for c in range(7)]
    if kind == 0:
        x, y, z = (u[i] * f(box) for i in range(3))
    else:
        x, y, z = ((f(2.0) * u[i] - f(1.0)) * f(box) for i in range(3))
    hh = mix(7)
    return dict(x=x.astype(f), y=y.astype(f), z=z.astype(f),
                vx=((u[3] - f(0.5)) * f(600.0)).astype(f), vy=((u[4] - f(0.5)) * f(600.0)).astype(f),
                vz=((u[5] - f(0.5)) * f(600.0)).astype(f), a=(f(0.5) + f(0.5) * u[6]).astype(f),
                id=g.astype(np.int64), rotation=((hh >> np.uint64(33)) % np.uint64(6)).astype(np.int32),
                replication=(hh & np.uint64(0xfffff)).astype(np.int32))


def f32_neighbours(v, k=2):
    """the 2k+1 floats around v (by ulp steps)"""
    v = np.float32(v)
    out = [v]
    lo = hi = v
    for _ in range(k):
        lo = np.nextafter(lo, np.float32(-np.inf), dtype=np.float32)
        hi = np.nextafter(hi, np.float32(np.inf), dtype=np.float32)
        out += [lo, hi]
    return np.array(sorted(out), dtype=np.float32)


def edge_fixture(theta_cut, phi_cut, seed=7, per_bound=64, octant_only=False):
    """Particles aimed at the window boundary: for each of the four bounds, directions whose exact
    angle sits within a few float ulps of the bound (several radii each, so the float rounding of x,y,z
    scatters them to both sides), plus the degenerate cases the reference guards or trips over:
    x == 0 with y>0 / y<0, on the z axis, on octant faces, negative octants."""
    rng = np.random.default_rng(seed)
    conv = 3.14159265 / 180.0 / 3600.0
    xs, ys, zs = [], [], []
    t_mid = 0.5 * (float(theta_cut[0]) + float(theta_cut[1]))
    p_mid = 0.5 * (float(phi_cut[0]) + float(phi_cut[1]))
    for which, bound in (("t", theta_cut[0]), ("t", theta_cut[1]), ("p", phi_cut[0]), ("p", phi_cut[1])):
        for _ in range(per_bound):
            jitter = rng.integers(-3, 4) * np.spacing(np.float32(abs(bound) + 1.0))
            if which == "t":
                t = (float(bound) + float(jitter)) * conv
                p = (p_mid + rng.uniform(-0.4, 0.4) * (float(phi_cut[1]) - float(phi_cut[0]))) * conv
            else:
                p = (float(bound) + float(jitter)) * conv
                t = (t_mid + rng.uniform(-0.4, 0.4) * (float(theta_cut[1]) - float(theta_cut[0]))) * conv
            r = rng.uniform(50, 4000)
            xs.append(r * np.sin(t) * np.cos(p)); ys.append(r * np.sin(t) * np.sin(p)); zs.append(r * np.cos(t))
    # degenerate directions
    deg = [(0.0, 5.0, 1.0), (0.0, -5.0, 1.0), (0.0, 5.0, -1.0), (3.0, 0.0, 1.0), (3.0, 4.0, 0.0),
           (1e-30, 1.0, 1.0), (1.0, 1e-30, 1.0), (1.0, 1.0, 1e-30), (1e-20, 1e-20, 1e-20),
           (1e18, 1e18, 1e18), (3e19, 1.0, 2.0), (-1.0, -2.0, 3.0), (-1.0, 2.0, 3.0), (1.0, -2.0, 3.0),
           (1.0, 2.0, -3.0), (100.0, 100.0, 100.0), (1.0, 1.0, 1.0e4), (1.0, 1.0e4, 1.0)]
    if not octant_only:
        deg += [(0.0, 0.0, 7.0), (0.0, 0.0, -7.0)]  # theta exactly 0 / pi; phi = atan(0/0) = NaN
    for (a, b, c) in deg:
        xs.append(a); ys.append(b); zs.append(c)
    n = len(xs)
    cols = dict(x=np.array(xs, dtype=np.float32), y=np.array(ys, dtype=np.float32), z=np.array(zs, dtype=np.float32))
    for name in ("vx", "vy", "vz"):
        cols[name] = rng.normal(0, 300, n).astype(np.float32)
    cols["a"] = rng.uniform(0.3, 1, n).astype(np.float32)
    cols["id"] = (np.arange(n, dtype=np.int64) + (1 << 40))
    cols["rotation"] = rng.integers(0, 6, n).astype(np.int32)
    cols["replication"] = rng.integers(0, 2 ** 20, n).astype(np.int32)
    return cols


def concat_cols(*parts):
    return {k: np.concatenate([p[k] for p in parts]) for k in IN_COLS}


def slice_cols(cols, lo, hi):
    return {k: np.ascontiguousarray(cols[k][lo:hi]) for k in IN_COLS}


def write_flat_step(root, step, cols, prefix="lc", names=IN_COLS):
    """The flat column layout both readers accept: <root>/<prefix><step>/lc_intrp_output.<step> (empty
    header) and one raw little-endian file per variable '<header>#<var>'."""
    d = os.path.join(root, f"{prefix}{step}")
    os.makedirs(d, exist_ok=True)
    header = os.path.join(d, f"lc_intrp_output.{step}")
    open(header, "wb").close()
    for nme in names:
        np.ascontiguousarray(cols[nme]).tofile(header + "#" + nme)
    return header


def write_gio_step(root, step, cols, prefix="lc", names=IN_COLS, nranks=3, big_endian=False, block_headers=False,
                   partitions=0, old_headers=False, seed=0, bad_crc=None, filter_name=None):
    """A lightcone step in the GenericIO on-disk format (written here from the format description, independently of
    the reader under test): global header, variable headers, one rank header per writing rank, optional block
    headers, the 8 closing CRC64 bytes after the header and after every data block (crc64_closing_bytes; `bad_crc`
    writes a wrong one behind the named variable's first block, `filter_name` marks every block of the second
    variable as written through that compression filter).  The particles are cut into `nranks` contiguous pieces of uneven size (one of them empty when nranks > 2).
    partitions > 0: the main file only carries the "$rank"/"$partition" map and the data goes to '<header>#<p>',
    ranks dealt round-robin to partitions and stored there in DESCENDING rank order, so a reader has to restore the
    global rank order.  old_headers: rank headers without GlobalRank, variable headers without ElementSize, global
    header without the Blocks fields (the format's size fields make such files readable)."""
    rng = np.random.default_rng(seed)
    n = cols["x"].shape[0]
    d = os.path.join(root, f"{prefix}{step}")
    os.makedirs(d, exist_ok=True)
    header = os.path.join(d, f"lc_intrp_output.{step}")
    E = ">" if big_endian else "<"
    cuts = np.sort(rng.integers(0, n + 1, nranks - 1)) if nranks > 1 else np.array([], dtype=np.int64)
    if nranks > 2:
        cuts[1] = cuts[0]  # an empty rank
    bounds = np.concatenate([[0], cuts, [n]]).astype(np.int64)

    def u64(*v):
        return np.array(v, dtype=E + "u8").tobytes()

    def f64(*v):
        return np.array(v, dtype=E + "f8").tobytes()

    def gio_bytes(ranks, variables):
        """ranks: list of (global_rank, {name: array}); variables: list of (name, numpy dtype string without order)"""
        nvars = len(variables)
        vars_size = 256 + 16 + (0 if old_headers else 8)
        ranks_size = 40 + (0 if old_headers else 8)
        use_blocks = block_headers and not old_headers
        global_size = 8 + 8 * (18 if old_headers else 20)
        blocks_size = 48 if use_blocks else 0
        vars_start = global_size
        ranks_start = vars_start + nvars * vars_size
        blocks_start = ranks_start + len(ranks) * ranks_size if use_blocks else 0
        header_size = ranks_start + len(ranks) * ranks_size + len(ranks) * nvars * blocks_size + 8
        # data layout
        off = header_size
        rank_start, block_start, payload, bad_done = [], [], [], []
        for (_gr, data) in ranks:
            rank_start.append(off)
            for (name, dt) in variables:
                a = np.ascontiguousarray(data[name]).astype(E + dt)
                if use_blocks:  # block headers allow (and the writer of the library leaves) gaps: add one
                    pad = b"\xAB" * 16
                    payload.append(pad)
                    off += len(pad)
                block_start.append((off, a.nbytes))
                payload.append(a.tobytes())
                closing = crc64_closing_bytes(a.tobytes())
                if bad_crc == name and a.nbytes:
                    closing = bytes([closing[0] ^ 0x40]) + closing[1:]
                    bad_done.append(1)
                payload.append(closing if not (bad_crc == name and len(bad_done) > 1) else crc64_closing_bytes(a.tobytes()))
                off += a.nbytes + 8
        total = sum(data[variables[0][0]].shape[0] for (_gr, data) in ranks)
        out = [b"HACC01B\0" if big_endian else b"HACC01L\0"]
        out.append(u64(header_size, total, 1, 1, 1, nvars, vars_size, vars_start, len(ranks), ranks_size, ranks_start, global_size))
        out.append(f64(0, 0, 0, 256, 256, 256))
        if not old_headers:
            out.append(u64(blocks_size, blocks_start))
        for (name, dt) in variables:
            size = np.dtype(dt).itemsize
            flags = (1 if dt[0] == "f" else 0) | (2 if dt[0] in "fi" else 0)
            out.append(name.encode().ljust(256, b"\0") + u64(flags, size) + (b"" if old_headers else u64(size)))
        for k, (gr, data) in enumerate(ranks):
            ne = data[variables[0][0]].shape[0]
            out.append(u64(gr % 4, gr // 4, 0, ne, rank_start[k]) + (b"" if old_headers else u64(gr)))
        if use_blocks:
            for k, (st, sz) in enumerate(block_start):
                filt = (filter_name or "").encode().ljust(8, b"\0") if (filter_name and k % nvars == 1) else b"\0" * 8
                out.append(filt + b"\0" * 24 + u64(st, sz))
        out.append(crc64_closing_bytes(b"".join(out)))
        blob = b"".join(out)
        assert len(blob) == header_size, (len(blob), header_size)
        return blob + b"".join(payload)

    variables = [(nme, {"id": "i8", "rotation": "i4", "replication": "i4"}.get(nme, "f4")) for nme in names]
    ranks = [(r, {nme: cols[nme][bounds[r]:bounds[r + 1]] for nme in names}) for r in range(nranks)]
    if partitions <= 0:
        open(header, "wb").write(gio_bytes(ranks, variables))
    else:
        part_of = np.arange(nranks, dtype=np.int32) % partitions
        open(header, "wb").write(gio_bytes([(0, {"$rank": np.arange(nranks, dtype=np.int32), "$partition": part_of})],
                                           [("$rank", "i4"), ("$partition", "i4")]))
        for p_ in range(partitions):
            mine = [ranks[r] for r in range(nranks) if part_of[r] == p_][::-1]
            open(header + f"#{p_}", "wb").write(gio_bytes(mine, variables))
    return header


_CRC64_TABLE = None


def crc64_xz(data, reg=0xFFFFFFFFFFFFFFFF):
    """CRC-64/XZ register after `data` (start ~0; the CRC is ~register): GenericIO's checksum.  Written here from the
    catalogue parameters (poly 0x42F0E1EBA9EA3693 reflected, init ~0, xorout ~0, check 0x995DC9BBDF1939FA),
    independently of the reader under test; byte-at-a-time, vectorised over a block with numpy."""
    global _CRC64_TABLE
    if _CRC64_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0xC96C5795D7870F42 if c & 1 else c >> 1
            t.append(c)
        _CRC64_TABLE = t
    t = _CRC64_TABLE
    for b in bytes(data):
        reg = t[(reg ^ b) & 0xFF] ^ (reg >> 8)
    return reg


def crc64_closing_bytes(data):
    """the 8 bytes GenericIO's crc64_invert appends: with them the CRC of data + bytes is ~0 (register 0)"""
    return int(crc64_xz(data)).to_bytes(8, "little")


def read_cutout_dir(d, step, names=OUT_COLS):
    out = {}
    for nme in names:
        p = os.path.join(d, f"{nme}.{step}.bin")
        out[nme] = np.fromfile(p, dtype=COL_DTYPES[nme])
    return out


def assert_cols_equal(got, want, names=OUT_COLS, what=""):
    """byte-for-byte equality of every column (NaNs compare by bit pattern)"""
    for nme in names:
        g = np.ascontiguousarray(got[nme]); w = np.ascontiguousarray(want[nme])
        assert g.dtype == w.dtype, f"{what}{nme}: dtype {g.dtype} != {w.dtype}"
        assert g.shape == w.shape, f"{what}{nme}: {g.shape[0]} survivors, expected {w.shape[0]}"
        if g.tobytes() != w.tobytes():
            bad = np.nonzero(g.view(np.uint8).reshape(g.shape[0], -1) != w.view(np.uint8).reshape(w.shape[0], -1))[0]
            i = int(bad[0])
            raise AssertionError(f"{what}{nme}: first mismatch at row {i}: got {g[i]!r} want {w[i]!r} "
                                 f"({len(set(bad.tolist()))} rows differ)")


# --------------------------------------------------------------------------------------------
# oracle (C restatement)
# --------------------------------------------------------------------------------------------
class _ColsIn(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, C.c_void_p) for k in IN_COLS]


class _ColsOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in OUT_COLS]


class OracleHalo(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("halo_r", C.c_float), ("half_box_rad", C.c_float),
                ("theta_cut", C.c_float * 2), ("phi_cut", C.c_float * 2), ("k", C.c_float * 3),
                ("beta", C.c_float), ("K", C.c_float * 9), ("R", C.c_float * 9), ("R_inv", C.c_float * 9),
                ("corners_rot", C.c_float * 12), ("corners_sph", C.c_float * 8),
                ("theta_rough", C.c_float * 2), ("phi_rough", C.c_float * 2)]


_oracle = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            build_oracle()
        lib = C.CDLL(so)
        lib.oracle_cutout_const.restype = C.c_int64
        lib.oracle_cutout_halo.restype = C.c_int64
        lib.oracle_scatter_kept.restype = C.c_int64
        lib.oracle_scatter_kept.argtypes = [C.c_int64, C.c_int]
        lib.oracle_aToZ.restype = C.c_float
        lib.oracle_aToZ.argtypes = [C.c_float]
        lib.oracle_zToStep.restype = C.c_int
        lib.oracle_zToStep.argtypes = [C.c_float, C.c_int, C.c_float]
        lib.oracle_stepToZ.restype = C.c_float
        lib.oracle_stepToZ.argtypes = [C.c_float, C.c_int, C.c_float]
        lib.oracle_cli_bounds.argtypes = [C.c_float, C.c_float, C.c_void_p]
        lib.oracle_halo_setup.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        lib.oracle_libc_version.restype = C.c_char_p
        _oracle = lib
    return _oracle


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _cols_in(cols, n=None):
    ci = _ColsIn()
    ci.n = int(n if n is not None else cols["x"].shape[0])
    for k in IN_COLS:
        a = cols.get(k)
        if a is not None:
            assert a.flags["C_CONTIGUOUS"] and a.dtype == COL_DTYPES[k], k
        setattr(ci, k, _ptr(a))
    return ci


def _alloc_out(n):
    arrs = {k: np.empty(max(n, 1), dtype=COL_DTYPES[k]) for k in OUT_COLS}
    co = _ColsOut()
    for k in OUT_COLS:
        setattr(co, k, _ptr(arrs[k]))
    return arrs, co


def oracle_cli_bounds(center_deg, half_deg):
    out = np.zeros(2, dtype=np.float32)
    oracle().oracle_cli_bounds(C.c_float(center_deg), C.c_float(half_deg), _ptr(out))
    return out


def oracle_theta_phi(x, y, z, guards):
    n = x.shape[0]
    th = np.empty(n, dtype=np.float32); ph = np.empty(n, dtype=np.float32)
    fn = oracle().oracle_theta_phi_uc2 if guards else oracle().oracle_theta_phi_uc1
    fn(C.c_int64(n), _ptr(x), _ptr(y), _ptr(z), _ptr(th), _ptr(ph))
    return th, ph


def oracle_cutout_const(cols, theta_cut, phi_cut):
    n = cols["x"].shape[0]
    ci = _cols_in(cols)
    arrs, co = _alloc_out(n)
    tc = np.asarray(theta_cut, dtype=np.float32); pc = np.asarray(phi_cut, dtype=np.float32)
    k = oracle().oracle_cutout_const(C.byref(ci), _ptr(tc), _ptr(pc), C.byref(co))
    return {nme: arrs[nme][:k].copy() for nme in OUT_COLS}


def oracle_halo_setup(pos, box_length):
    h = OracleHalo()
    p = np.asarray(pos, dtype=np.float32)
    oracle().oracle_halo_setup(_ptr(p), C.c_float(box_length), C.byref(h))
    return h


def oracle_cutout_halo(cols, halo, pos_only=False, ref_numranks=1):
    n = cols["x"].shape[0]
    ci = _cols_in(cols)
    arrs, co = _alloc_out(n)
    rough = C.c_int64(0)
    k = oracle().oracle_cutout_halo(C.byref(ci), C.byref(halo), C.c_int(int(pos_only)), C.c_int(ref_numranks),
                                   C.byref(co), C.byref(rough))
    names = OUT_COLS_POSONLY if pos_only else OUT_COLS
    return {nme: arrs[nme][:k].copy() for nme in names}, int(rough.value)


def oracle_libm(fn, arr):
    out = np.empty_like(arr)
    getattr(oracle(), f"oracle_{fn}_array")(C.c_int64(arr.shape[0]), _ptr(arr), _ptr(out))
    return out


# --------------------------------------------------------------------------------------------
# real reference (oracle/_ref), optional
# --------------------------------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "libcutout_ref.so"))


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(os.path.join(REF_DIR, "libcutout_ref.so"))
        lib.ref_aToZ.restype = C.c_float; lib.ref_aToZ.argtypes = [C.c_float]
        lib.ref_zToStep.restype = C.c_int; lib.ref_zToStep.argtypes = [C.c_float]
        lib.ref_stepToZ.restype = C.c_float; lib.ref_stepToZ.argtypes = [C.c_float]
        lib.ref_scatter_kept.restype = C.c_longlong; lib.ref_scatter_kept.argtypes = [C.c_longlong, C.c_int]
        lib.shim_trace_name.restype = C.c_char_p
        lib.shim_trace_time.restype = C.c_double
        lib.shim_trace_value.restype = C.c_longlong
        lib.gio_shim_register.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
        _ref = lib
    return _ref


def _strs(lst):
    arr = (C.c_char_p * len(lst))(*[s.encode() for s in lst])
    return arr


class RefRun:
    """Run the real reference's processLC on in-memory columns; outputs land in a temp dir."""

    def __init__(self, cols, step=487, names=IN_COLS):
        self.lib = ref()
        self.step = step
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        self.tmp = tempfile.mkdtemp(prefix="ccref_", dir=base)
        self.in_dir = os.path.join(self.tmp, "in") + "/"
        self.out_dir = os.path.join(self.tmp, "out") + "/"
        os.makedirs(os.path.join(self.in_dir, f"lc{step}"))
        os.makedirs(self.out_dir)
        header = os.path.join(self.in_dir, f"lc{step}", f"lc_intrp_output.{step}")
        open(header, "wb").close()
        self._keep = []
        self.lib.gio_shim_clear()
        for nme in names:
            a = np.ascontiguousarray(cols[nme])
            self._keep.append(a)
            # the reference builds the path as dir + prefix + step + "/" + file (processLC.cpp:82,91)
            key = f"{self.in_dir}lc{step}/lc_intrp_output.{step}#{nme}".encode()
            self.lib.gio_shim_register(key, _ptr(a), a.nbytes)

    def trace(self):
        n = self.lib.shim_trace_count()
        return [(self.lib.shim_trace_name(i).decode(), self.lib.shim_trace_time(i)) for i in range(n)]

    def const(self, theta_cut, phi_cut):
        self.lib.shim_trace_reset()
        rc = self.lib.ref_processLC_const(self.in_dir.encode(), self.out_dir.encode(), _strs([str(self.step)]), 1,
                                          C.c_float(theta_cut[0]), C.c_float(theta_cut[1]),
                                          C.c_float(phi_cut[0]), C.c_float(phi_cut[1]), 1)
        assert rc == 0, rc
        return read_cutout_dir(os.path.join(self.out_dir, f"lcCutout{self.step}"), self.step)

    def const_loop_seconds(self):
        """duration of processLC.cpp:276-310 in the last const() run: last MPI_File_open -> next MPI_Barrier"""
        tr = self.trace()
        last_open = max(i for i, (nme, _) in enumerate(tr) if nme == "MPI_File_open")
        nxt = next(i for i in range(last_open + 1, len(tr)) if tr[i][0] == "MPI_Barrier")
        return tr[nxt][1] - tr[last_open][1]

    def halo(self, pos, box_length, pos_only=False):
        self.lib.shim_trace_reset()
        p = np.asarray(pos, dtype=np.float32)
        props = np.full(6, -1.0, dtype=np.float32)
        sub = os.path.join(self.out_dir, f"lcCutout{self.step}")
        if os.path.isdir(sub):
            shutil.rmtree(sub)  # never re-enter an existing dir: SURVEY.md App. A Q12
        rc = self.lib.ref_processLC_halo(self.in_dir.encode(), _strs([self.out_dir]), 1, _strs([str(self.step)]), 1,
                                         _ptr(p), _ptr(props), 6, C.c_float(box_length), 0, 0, 0, int(pos_only),
                                         1, 0, 1)
        assert rc == 0, rc
        names = OUT_COLS_POSONLY if pos_only else OUT_COLS
        # the Allgathers that carry int counts (processLC.cpp:1569-1572): cutout_size, then rough_cutout_size;
        # the two 8-byte Np gathers of the redistribution (:979, :1108) come earlier in the trace
        n = self.lib.shim_trace_count()
        gathers = [self.lib.shim_trace_value(i) for i in range(n)
                   if self.lib.shim_trace_name(i) == b"MPI_Allgather"]
        self.last_rough_count = int(gathers[-1])
        self.last_count = int(gathers[-2])
        self.last_np_after_redistribution = int(gathers[1])
        return read_cutout_dir(sub, self.step, names)

    def close(self):
        self.lib.gio_shim_clear()
        shutil.rmtree(self.tmp, ignore_errors=True)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def ref_rotation_setup(pos):
    lib = ref()
    a = np.asarray(pos, dtype=np.float32)
    r = np.float32(np.sqrt(np.float32(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])))
    b = np.array([r, 0, 0], dtype=np.float32)
    k = np.zeros(3, np.float32); beta = C.c_float(0); K = np.zeros(9, np.float32)
    R = np.zeros(9, np.float32); Ri = np.zeros(9, np.float32)
    lib.ref_rotation_setup(_ptr(a), _ptr(b), _ptr(k), C.byref(beta), _ptr(K), _ptr(R), _ptr(Ri))
    return dict(k=k, beta=np.float32(beta.value), K=K, R=R, R_inv=Ri, halo_r=r)
