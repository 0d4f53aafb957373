import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, ctypes as C
from __graft_entry__ import load_package
cc = load_package()
n = 100_000_000
host = {}
for k in cc.IN_COLS:
    a, addr = cc.pinned_empty(n, cc.COL_DTYPES[k]); host[k] = a
cc._check(cc.lib().cc_synth_host(n, 20261019, 0, 0, C.c_float(300.0), *[host[k].ctypes.data_as(C.c_void_p) for k in cc.IN_COLS]))
rng = np.random.default_rng(1000)
pos = rng.uniform(30.0, 280.0, (1000, 3)).astype(np.float32)
halos = [cc.halo_setup(tuple(float(v) for v in p), 15.0).halo for p in pos]
ctx = cc.Context(0)
for pos_only in (False, True):
    for it in range(4):
        t0 = time.perf_counter()
        ctx.cutout_halo_streamed(host, halos, pos_only=pos_only, n_limit_global=-1, index_base=0)
        t1 = time.perf_counter()
        counts, _ = ctx.sync()
        t2 = time.perf_counter()
        out, off = ctx.fetch_all(counts)
        t3 = time.perf_counter()
        print(f"pos_only={pos_only} enqueue {1e3*(t1-t0):.1f} ms, sync {1e3*(t2-t1):.1f} ms, fetch {1e3*(t3-t2):.1f} ms; device scan {ctx.last_scan_ms():.1f} total {ctx.last_total_ms():.1f} ms; survivors {int(counts.sum())}")
