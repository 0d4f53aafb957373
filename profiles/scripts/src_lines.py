"""per-CUDA-source-line instruction / stall-sample distribution of one launch of an ncu report
usage: python profiles/scripts/src_lines.py report.ncu-rep <launch-skip> [top]"""
import csv, subprocess, sys, io, collections, os
rep, skip = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
per = collections.defaultdict(lambda: [0, 0, ""])
cur, fname, hdr = None, "", None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        fname = os.path.basename(r[1]); cur = None
        continue
    if "# Samples" in r:
        hdr = r
        iline, isamp, iexe, isrc = hdr.index("Line No"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
        continue
    if hdr is None or len(r) <= iexe:
        continue
    if r[iline]:
        try:
            cur = (fname, int(r[iline])); per[cur][2] = r[isrc]
        except ValueError:
            cur = None
        continue
    if cur is None or r[iexe] in ("-", "") or r[isamp] in ("-", ""):
        continue
    per[cur][0] += int(r[iexe]); per[cur][1] += int(r[isamp])
tot_e = sum(v[0] for v in per.values()) or 1; tot_s = sum(v[1] for v in per.values()) or 1
print("warp instr", tot_e, "samples", tot_s)
byfile = collections.defaultdict(lambda: [0, 0])
for (f, _), v in per.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, v in byfile.items():
    print(f"  {f}: instr {v[0]/tot_e*100:5.1f}%  samples {v[1]/tot_s*100:5.1f}%")
for k in sorted(sorted(per, key=lambda k: -per[k][1])[:ntop]):
    e, s, src = per[k]
    print(f"{k[0][:14]:14s}{k[1]:5d} instr {e/tot_e*100:5.1f}%  samples {s/tot_s*100:5.1f}%  {src.strip()[:100]}")
