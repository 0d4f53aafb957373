"""Does creating an NCCL communicator slow k2_cutout_halo_cells down?  One process, two GPUs."""
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from __graft_entry__ import load_package
cc = load_package()
from cuda import cudart

def limits(tag):
    out = []
    for name in ("cudaLimitStackSize", "cudaLimitMallocHeapSize", "cudaLimitPrintfFifoSize", "cudaLimitMaxL2FetchGranularity", "cudaLimitPersistingL2CacheSize"):
        err, v = cudart.cudaDeviceGetLimit(getattr(cudart.cudaLimit, name))
        out.append((name, v))
    err, cfg = cudart.cudaDeviceGetCacheConfig()
    print(tag, out, "cache", cfg, flush=True)

rng = np.random.default_rng(1000)
pos = rng.uniform(30.0, 280.0, (1000, 3)).astype(np.float32)
halos = [cc.halo_setup(tuple(float(v) for v in p), 15.0).halo for p in pos]
ctx0, ctx1 = cc.Context(0), cc.Context(1)
ctx0.synth(125_000_000, seed=20261019, index_base=0, kind=0, box=300.0)
def run(tag):
    t = []
    for _ in range(6):
        ctx0.cutout_halo(halos)
        ctx0.sync()
        t.append(ctx0.last_scan_ms())
    print(tag, "scan ms", [round(x, 3) for x in t], flush=True)
cudart.cudaSetDevice(0)
limits("before")
run("before comm")
arr = (C.c_void_p * 2)(ctx0._h, ctx1._h)
cc._check(cc.lib().cc_comm_init_all(arr, 2))
cudart.cudaSetDevice(0)
limits("after")
run("after comm_init_all")
