"""per-100-line instruction / stall-sample distribution of a kernel from an ncu report (source page, SASS)
usage: python profiles/scripts/src_blocks.py report.ncu-rep units_per_launch [top]"""
import csv, subprocess, sys, io
rep, units = sys.argv[1], float(sys.argv[2])
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, ie, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ip = hdr.index("Predicated-On Thread Instructions Executed")
data = [(int(r[ie]), int(r[ism] or 0), r[isrc], int(r[ip])) for r in rows[2:] if len(r) > ie]
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("lines", len(data), "warp instr", tot, "samples", ts, "instr/unit %.1f" % (tot / units))
for b in range(0, len(data), 100):
    blk = data[b:b + 100]
    e = sum(d[0] for d in blk); s = sum(d[1] for d in blk)
    if e > tot * 0.003:
        print(b, f"{e/1e6:7.1f}M instr {e/tot*100:5.1f}%  samples {s/ts*100:5.1f}%  per-unit {e/units:6.1f}")
for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][1])[:ntop]):
    print(i, data[i][0], data[i][3], data[i][1], data[i][2][:90])
