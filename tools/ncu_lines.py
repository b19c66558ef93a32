"""Aggregate an Nsight Compute source page (ncu -i REP --page source --csv --print-source sass,cuda) per CUDA source line:
executed warp instructions and stall samples.  usage: python tools/ncu_lines.py REP.ncu-rep [top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
rows = rows[:hi + 1] + [r for r in rows[hi + 1:]]
ix = {n: i for i, n in enumerate(hdr)}
# the csv has two "Source" columns: cuda source line text (col 1) then SASS (col 3)
inst_i = ix["Instructions Executed"]; samp_i = ix["# Samples"]
agg = collections.OrderedDict()
cur_file = ""
tot_i = tot_s = 0
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    try:
        n = int(r[inst_i] or 0); s = int(r[samp_i] or 0)
    except ValueError:
        continue
    key = (cur_file + ":" + r[0], r[1].strip()[:100])
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += n; a[1] += s; a[2] += 1
    tot_i += n; tot_s += s
print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
for (ln, src), (n, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% stall  sass {k:4d}  L{ln}: {src}")
