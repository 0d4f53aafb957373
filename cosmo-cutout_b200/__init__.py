"""cosmo-cutout_b200 -- ctypes binding of the C ABI (include/cosmo_cutout_b200.h).

The product is the C++/CUDA library `lib/libcosmo_cutout_b200.so` and the `bin/lc_cutout` CLI; this module only
lets Python (tests, bench.py) call the same C entry points a C++ host would.  It performs no arithmetic of its
own and has no fallback: if the shared library is missing or no B200 is present the calls raise.

The directory name carries a hyphen (it mirrors the reference repo's name), so import it through
`__graft_entry__.load_package()` which registers it as module `cosmo_cutout_b200`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcosmo_cutout_b200.so")
CLI_PATH = os.path.join(_HERE, "bin", "lc_cutout")

IN_COLS = ("x", "y", "z", "vx", "vy", "vz", "a", "id", "rotation", "replication")
OUT_COLS = ("x", "y", "z", "vx", "vy", "vz", "redshift", "theta", "phi", "id", "rotation", "replication")
OUT_COLS_POSONLY = ("x", "y", "z", "redshift", "theta", "phi", "id")
OUT_COLS_ALL = OUT_COLS + ("theta_u", "index")
COL_DTYPES = dict(x=np.float32, y=np.float32, z=np.float32, vx=np.float32, vy=np.float32, vz=np.float32,
                  a=np.float32, redshift=np.float32, theta=np.float32, phi=np.float32, id=np.int64,
                  rotation=np.int32, replication=np.int32, theta_u=np.float32, index=np.int64)


class CutoutError(RuntimeError):
    pass


class ColsIn(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in IN_COLS]


class ColsOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in OUT_COLS_ALL]


class Halo(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("theta_cut", C.c_float * 2), ("phi_cut", C.c_float * 2),
                ("theta_rough", C.c_float * 2), ("phi_rough", C.c_float * 2)]


class HaloInfo(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("halo_r", C.c_float), ("half_box_rad", C.c_float), ("k", C.c_float * 3),
                ("beta", C.c_float), ("K", C.c_float * 9), ("R", C.c_float * 9), ("R_inv", C.c_float * 9),
                ("corners_rot", C.c_float * 12), ("corners_sph", C.c_float * 8), ("halo", Halo)]


# every symbol include/cosmo_cutout_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = (
    "cc_version", "cc_last_error", "cc_device_count", "cc_create", "cc_destroy", "cc_stream",
    "cc_halo_setup", "cc_cli_bounds", "cc_ref_scatter_kept",
    "cc_step_upload", "cc_step_attach", "cc_step_synth", "cc_synth_host", "cc_step_download", "cc_step_size",
    "cc_step_release", "cc_step_select", "cc_device_mem", "cc_bytes_uploaded",
    "cc_cutout_const", "cc_cutout_halo", "cc_cutout_sync", "cc_cutout_fetch", "cc_cutout_fetch_all", "cc_cutout_device_cols",
    "cc_cutout_const_streamed", "cc_cutout_halo_streamed", "cc_host_alloc", "cc_host_free", "cc_host_register",
    "cc_host_unregister",
    "cc_comm_unique_id", "cc_comm_init", "cc_comm_init_all", "cc_exchange_counts", "cc_comm_rank", "cc_comm_size",
    "cc_comm_peer_handle", "cc_comm_peer_attach",
    "cc_shard_range", "cc_merge_order",
    "cc_last_scan_ms", "cc_last_total_ms", "cc_launch_count", "cc_last_stats", "cc_last_stats_ex", "cc_set_option",
    "cc_debug_eval", "cc_debug_eval_host",
)

_lib = None


def lib():
    """the loaded shared library; raises if it has not been built"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CutoutError(f"{LIB_PATH} is missing: run `make -C cosmo-cutout_b200` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.cc_version.restype = C.c_char_p
        L.cc_last_error.restype = C.c_char_p
        L.cc_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.cc_destroy.argtypes = [C.c_void_p]
        L.cc_stream.restype = C.c_void_p
        L.cc_stream.argtypes = [C.c_void_p]
        L.cc_halo_setup.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.cc_cli_bounds.argtypes = [C.c_float, C.c_float, C.c_void_p]
        L.cc_cli_bounds.restype = None
        L.cc_ref_scatter_kept.restype = C.c_int64
        L.cc_ref_scatter_kept.argtypes = [C.c_int64, C.c_int]
        L.cc_step_upload.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.cc_step_attach.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.cc_step_synth.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_float]
        L.cc_synth_host.argtypes = [C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_float] + [C.c_void_p] * 10
        L.cc_step_download.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 10
        L.cc_step_size.restype = C.c_int64
        L.cc_step_size.argtypes = [C.c_void_p]
        L.cc_step_release.argtypes = [C.c_void_p]
        L.cc_step_select.argtypes = [C.c_void_p, C.c_int]
        L.cc_device_mem.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.cc_bytes_uploaded.restype = C.c_int64
        L.cc_bytes_uploaded.argtypes = [C.c_void_p]
        L.cc_cutout_const.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.cc_cutout_halo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64]
        L.cc_cutout_sync.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.cc_cutout_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.cc_cutout_fetch_all.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.cc_cutout_device_cols.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cc_cutout_const_streamed.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.cc_cutout_halo_streamed.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                              C.c_int, C.c_int64]
        L.cc_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.cc_host_free.argtypes = [C.c_void_p]
        L.cc_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.cc_host_unregister.argtypes = [C.c_void_p]
        L.cc_comm_unique_id.argtypes = [C.c_void_p]
        L.cc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.cc_comm_init_all.argtypes = [C.c_void_p, C.c_int]
        L.cc_exchange_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cc_comm_peer_handle.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.cc_comm_peer_attach.argtypes = [C.c_void_p, C.c_void_p]
        L.cc_comm_rank.argtypes = [C.c_void_p]
        L.cc_comm_size.argtypes = [C.c_void_p]
        L.cc_shard_range.restype = None
        L.cc_shard_range.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.cc_merge_order.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cc_last_scan_ms.restype = C.c_double
        L.cc_last_scan_ms.argtypes = [C.c_void_p]
        L.cc_last_total_ms.restype = C.c_double
        L.cc_last_total_ms.argtypes = [C.c_void_p]
        L.cc_launch_count.restype = C.c_int64
        L.cc_launch_count.argtypes = [C.c_void_p]
        L.cc_last_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.cc_last_stats_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.cc_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.cc_debug_eval.argtypes = [C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 4
        L.cc_debug_eval_host.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 4
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise CutoutError(f"cosmo_cutout_b200 error {rc}: {lib().cc_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _cols_in_struct(cols, ptr_of=_ptr):
    s = ColsIn()
    for k in IN_COLS:
        a = cols.get(k)
        if a is not None and isinstance(a, np.ndarray):
            assert a.flags["C_CONTIGUOUS"] and a.dtype == COL_DTYPES[k], k
        setattr(s, k, ptr_of(a) if a is not None else None)
    return s


# ---------------------------------------------------------------------------------- host helpers
def halo_setup(pos, box_length):
    """cc_halo_setup: per-halo bounds and rotation (processLC.cpp:512-733) -> HaloInfo"""
    info = HaloInfo()
    p = np.asarray(pos, dtype=np.float32)
    _check(lib().cc_halo_setup(_ptr(p), C.c_float(box_length), C.byref(info)))
    return info


def cli_bounds(center_deg, half_deg):
    out = np.zeros(2, dtype=np.float32)
    lib().cc_cli_bounds(C.c_float(center_deg), C.c_float(half_deg), _ptr(out))
    return out


def ref_scatter_kept(n, numranks=1):
    return int(lib().cc_ref_scatter_kept(int(n), int(numranks)))


def shard_range(n, nranks, rank):
    lo, hi = C.c_int64(0), C.c_int64(0)
    lib().cc_shard_range(int(n), int(nranks), int(rank), C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def merge_order(theta_u_parts, index_parts):
    """cc_merge_order: merged row order of per-rank (theta_u, index)-sorted survivor lists -> (src_rank, src_row)"""
    n = len(theta_u_parts)
    counts = np.array([t.shape[0] for t in theta_u_parts], dtype=np.int64)
    tp = (C.c_void_p * n)(*[t.ctypes.data for t in theta_u_parts])
    ip = (C.c_void_p * n)(*[t.ctypes.data for t in index_parts])
    tot = int(counts.sum())
    src_rank = np.zeros(tot, dtype=np.int32)
    src_row = np.zeros(tot, dtype=np.int64)
    _check(lib().cc_merge_order(n, _ptr(counts), tp, ip, _ptr(src_rank), _ptr(src_row)))
    return src_rank, src_row


def synth_host(n, seed, index_base=0, kind=0, box=300.0):
    cols = {k: np.empty(n, dtype=COL_DTYPES[k]) for k in IN_COLS}
    _check(lib().cc_synth_host(int(n), int(seed), int(index_base), int(kind), C.c_float(box),
                               *[_ptr(cols[k]) for k in IN_COLS]))
    return cols


def debug_eval_host(op, in0, in1=None, in2=None):
    out = np.empty_like(in0)
    _check(lib().cc_debug_eval_host(int(op), in0.shape[0], _ptr(in0), _ptr(in1), _ptr(in2), _ptr(out)))
    return out


def pinned_empty(n, dtype):
    """numpy array over cudaHostAlloc'ed (portable, mapped) memory; keep the returned array alive, free with
    pinned_free(arr)"""
    dtype = np.dtype(dtype)
    p = C.c_void_p()
    nbytes = max(int(n) * dtype.itemsize, 16)
    _check(lib().cc_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n))
    return arr, p.value


def pinned_free(addr):
    _check(lib().cc_host_free(C.c_void_p(addr)))


# ---------------------------------------------------------------------------------- context
class Context:
    """one GPU: resident lightcone step + cutouts (C ABI cc_ctx)"""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().cc_create(int(device), C.byref(self._h)))
        self.device = device
        self._keep = None
        self.nhalos = 0
        self.pos_only = False
        self.kind = None
        self._halo_cache = None
        self._count_bufs = None
        self._xchg_bufs = None

    def close(self):
        if self._h:
            lib().cc_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # step residency
    def upload(self, cols, index_base=0):
        n = int(cols["x"].shape[0])
        s = _cols_in_struct(cols)
        _check(lib().cc_step_upload(self._h, n, C.byref(s), int(index_base)))

    def attach(self, n, dev_ptrs, index_base=0):
        """dev_ptrs: dict name -> integer device address (e.g. torch tensor .data_ptr())"""
        s = ColsIn()
        for k in IN_COLS:
            setattr(s, k, dev_ptrs.get(k))
        _check(lib().cc_step_attach(self._h, int(n), C.byref(s), int(index_base)))

    def synth(self, n, seed, index_base=0, kind=0, box=300.0):
        _check(lib().cc_step_synth(self._h, int(n), int(seed), int(index_base), int(kind), C.c_float(box)))

    def download(self, first, count, names=IN_COLS):
        cols = {k: np.empty(count, dtype=COL_DTYPES[k]) for k in names}
        _check(lib().cc_step_download(self._h, int(first), int(count), *[_ptr(cols.get(k)) for k in IN_COLS]))
        return cols

    def step_size(self):
        return int(lib().cc_step_size(self._h))

    def release(self):
        _check(lib().cc_step_release(self._h))

    def select(self, slot):
        """cc_step_select: make resident-step slot `slot` the current one"""
        _check(lib().cc_step_select(self._h, int(slot)))

    def device_mem(self):
        f, t = C.c_int64(0), C.c_int64(0)
        _check(lib().cc_device_mem(self._h, C.byref(f), C.byref(t)))
        return int(f.value), int(t.value)

    def bytes_uploaded(self):
        return int(lib().cc_bytes_uploaded(self._h))

    # cutouts
    def cutout_const(self, theta_cut, phi_cut):
        tc = np.asarray(theta_cut, dtype=np.float32)
        pc = np.asarray(phi_cut, dtype=np.float32)
        _check(lib().cc_cutout_const(self._h, _ptr(tc), _ptr(pc)))
        self.nhalos, self.pos_only, self.kind = 1, False, "const"

    def _halo_array(self, halos):
        """ctypes array of the halo structs; the same list object passed again reuses the marshalled array (the
        structs are copied on marshalling: mutate-and-reuse callers must pass a new list)"""
        if self._halo_cache is None or self._halo_cache[0] is not halos or len(halos) != self._halo_cache[2]:
            self._halo_cache = (halos, (Halo * len(halos))(*halos), len(halos))
        return self._halo_cache[1]

    def cutout_halo(self, halos, pos_only=False, n_limit_global=-1):
        arr = self._halo_array(halos)
        _check(lib().cc_cutout_halo(self._h, len(halos), arr, int(bool(pos_only)), int(n_limit_global)))
        self.nhalos, self.pos_only, self.kind = len(halos), bool(pos_only), "halo"

    def cutout_const_streamed(self, cols, theta_cut, phi_cut, index_base=0):
        tc = np.asarray(theta_cut, dtype=np.float32)
        pc = np.asarray(phi_cut, dtype=np.float32)
        s = _cols_in_struct(cols)
        self._keep = cols
        _check(lib().cc_cutout_const_streamed(self._h, int(cols["x"].shape[0]), C.byref(s), int(index_base),
                                              _ptr(tc), _ptr(pc)))
        self.nhalos, self.pos_only, self.kind = 1, False, "const"

    def cutout_halo_streamed(self, cols, halos, pos_only=False, n_limit_global=-1, index_base=0):
        arr = self._halo_array(halos)
        s = _cols_in_struct(cols)
        self._keep = cols
        _check(lib().cc_cutout_halo_streamed(self._h, int(cols["x"].shape[0]), C.byref(s), int(index_base),
                                             len(halos), arr, int(bool(pos_only)), int(n_limit_global)))
        self.nhalos, self.pos_only, self.kind = len(halos), bool(pos_only), "halo"

    def sync(self):
        n = max(self.nhalos, 1)
        if self._count_bufs is None or self._count_bufs[0].shape[0] != n:
            c, r = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
            self._count_bufs = (c, r, _ptr(c), _ptr(r))
        c, r, pc, pr = self._count_bufs
        _check(lib().cc_cutout_sync(self._h, pc, pr))
        return c[:self.nhalos].copy(), r[:self.nhalos].copy()

    def fetch(self, halo=0, names=None, count=None):
        if names is None:
            names = OUT_COLS_POSONLY if self.pos_only else OUT_COLS
        if count is None:
            count = int(self.sync()[0][halo])
        out = {k: np.empty(count, dtype=COL_DTYPES[k]) for k in names}
        s = ColsOut()
        for k in OUT_COLS_ALL:
            setattr(s, k, _ptr(out.get(k)))
        _check(lib().cc_cutout_fetch(self._h, int(halo), 0, int(count), C.byref(s)))
        return out

    def fetch_all(self, counts, names=None, out=None):
        """all halos with one copy per column: returns (columns, row offsets [H+1]); halo h = rows off[h]:off[h+1].
        `out` may hold preallocated (e.g. pinned) arrays of at least sum(counts) rows per column"""
        if names is None:
            names = OUT_COLS_POSONLY if self.pos_only else OUT_COLS
        off = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        total = int(off[-1])
        if out is None:
            out = {k: np.empty(total, dtype=COL_DTYPES[k]) for k in names}
        s = ColsOut()
        for k in OUT_COLS_ALL:
            setattr(s, k, _ptr(out.get(k)) if k in names else None)
        _check(lib().cc_cutout_fetch_all(self._h, total, C.byref(s)))
        return {k: out[k][:total] for k in names}, off

    def exchange_counts(self):
        """cc_exchange_counts -> (counts[rank, halo], rough[rank, halo], offsets[halo]); the arrays are reused by the next
        call with the same shape (copy them to keep them)"""
        r = max(self.nhalos, 1)
        cache = self._xchg_bufs
        if cache is None or cache[0] != r:
            n = self.comm_size()
            allc, allr, off = np.zeros(n * r, dtype=np.int64), np.zeros(n * r, dtype=np.int64), np.zeros(r, dtype=np.int64)
            cache = self._xchg_bufs = (r, allc.reshape(n, r), allr.reshape(n, r), off, _ptr(allc), _ptr(allr), _ptr(off))
        rc = lib().cc_exchange_counts(self._h, cache[4], cache[5], cache[6])
        if rc:
            _check(rc)
        return cache[1], cache[2], cache[3]

    # communicator
    def comm_init(self, nranks, rank, id_bytes):
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        _check(lib().cc_comm_init(self._h, int(nranks), int(rank), buf))
        self._xchg_bufs = None

    def comm_peer_handle(self, nranks, rank):
        """allocate this GPU's count-exchange mailbox and return its 64-byte IPC handle"""
        buf = C.create_string_buffer(64)
        _check(lib().cc_comm_peer_handle(self._h, int(nranks), int(rank), buf))
        self._xchg_bufs = None
        return bytes(buf.raw)

    def comm_peer_attach(self, handles):
        """handles: the 64-byte handles of all ranks, in rank order"""
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob))
        _check(lib().cc_comm_peer_attach(self._h, buf))

    def comm_rank(self):
        return int(lib().cc_comm_rank(self._h))

    def comm_size(self):
        return int(lib().cc_comm_size(self._h))

    # measurement
    def last_scan_ms(self):
        return float(lib().cc_last_scan_ms(self._h))

    def last_total_ms(self):
        return float(lib().cc_last_total_ms(self._h))

    def launch_count(self):
        return int(lib().cc_launch_count(self._h))

    def stats(self):
        out = np.zeros(4, dtype=np.int64)
        _check(lib().cc_last_stats(self._h, _ptr(out)))
        return dict(scanned=int(out[0]), candidates=int(out[1]), survivors=int(out[2]), rough=int(out[3]))

    def stats_ex(self):
        out = np.zeros(12, dtype=np.int64)
        _check(lib().cc_last_stats_ex(self._h, _ptr(out), 12))
        return dict(scanned=int(out[0]), candidates=int(out[1]), survivors=int(out[2]), rough=int(out[3]),
                    nan_theta=int(out[4]), payload_resident=bool(out[5]), payload_build_ms=out[6] / 1000.0,
                    tail_dropped=int(out[7]),
                    scan_kernels=("k1s_scan", "k2_cutout_halo", "k2c_sift+k2c_match+k2c_exact", "k2_cutout_halo_cells",
                                  "k2p_query")[int(out[8])],
                    particles_read=int(out[9]), index_resident=bool(out[10]), index_build_ms=out[11] / 1000.0)

    def set_option(self, key, value):
        """cc_set_option: payload|index=auto|always|never, exchange=auto|nccl|peer|eager, cells=auto|staged|mono,
        timing=on|off, gather=auto|ldst|bulk"""
        _check(lib().cc_set_option(self._h, key.encode(), value.encode()))

    def stream(self):
        return lib().cc_stream(self._h)

    def debug_eval(self, op, in0, in1=None, in2=None):
        out = np.empty_like(in0)
        _check(lib().cc_debug_eval(self._h, int(op), in0.shape[0], _ptr(in0), _ptr(in1), _ptr(in2), _ptr(out)))
        return out


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().cc_comm_unique_id(buf))
    return bytes(buf.raw)
