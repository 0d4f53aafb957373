import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from __graft_entry__ import load_package
cc = load_package()
ctx = cc.Context(0)
n = 100_000_000
ctx.synth(n, seed=20261019, index_base=0, kind=0, box=300.0)
for name, th, ph in (("cfg4", (90, 20), (0, 20)), ("const", (87, 2), (2, 2)), ("mid", (60, 10), (45, 10))):
    ctx.cutout_const(cc.cli_bounds(*th), cc.cli_bounds(*ph)); ctx.sync()
    ctx.cutout_const(cc.cli_bounds(*th), cc.cli_bounds(*ph)); ctx.sync()
    st = ctx.stats()
    print(name, st, "cand/surv = %.3f" % (st["candidates"] / max(1, st["survivors"])), "scan ms %.3f total %.3f" % (ctx.last_scan_ms(), ctx.last_total_ms()), flush=True)
