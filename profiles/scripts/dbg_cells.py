import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, harness as H
from __graft_entry__ import load_package
cc = load_package()
rng = np.random.default_rng(17)
cols = H.concat_cols(H.shell_fixture(300_000, seed=51), H.shell_fixture(100_000, seed=52, octant=False))
n = cols["x"].shape[0]
bad = rng.choice(n, 1_500, replace=False)
vals = np.array([np.inf, -np.inf, 0.0, -0.0, 1e-12, -1e-12, 1e-40, 3e38], dtype=np.float32)
for k, col in enumerate(("x", "y", "z")):
    v = vals if col != "z" else vals[2:]
    cols[col][bad[k::3]] = v[rng.integers(0, len(v), bad[k::3].shape[0])]
cols["x"][bad[:200]] = np.float32(1e-12)
cols["id"] = np.arange(n, dtype=np.int64)
centre = np.array([120.0, 90.0, 100.0])
pos = (centre + rng.uniform(-4, 4, (48, 3))).astype(np.float32)
boxes = rng.uniform(40.0, 90.0, 48)
infos = [cc.halo_setup(p, b) for p, b in zip(pos, boxes)]
ctx = cc.Context(0)
ctx.upload(cols)
ctx.cutout_halo([i.halo for i in infos])
counts, rough = ctx.sync()
print("bins" if not os.environ.get("CC_NO_BINS") else "nobins", ctx.stats())
for h in range(0, 48, 5):
    want, want_rough = H.oracle_cutout_halo(cols, H.oracle_halo_setup(pos[h], boxes[h]))
    got = ctx.fetch(h)
    print(h, int(counts[h]), want["id"].shape[0], int(rough[h]), want_rough)
    extra = np.setdiff1d(got["id"], want["id"]); miss = np.setdiff1d(want["id"], got["id"])
    u, c = np.unique(got["id"], return_counts=True)
    print("  extra", extra[:10], "missing", miss[:10], "dups", u[c > 1][:10])
    for i in list(extra[:5]) + list(miss[:5]):
        print("   ", i, cols["x"][i], cols["y"][i], cols["z"][i])
