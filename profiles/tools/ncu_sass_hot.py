#!/usr/bin/env python3
"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv`: stall samples by reason, by opcode, and the top instructions."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[start - 1][:2])
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[start + 1:] if len(r) == len(hdr)]
S = "Warp Stall Sampling (All Samples)"
f = lambda r, k: float(r[ix[k]] or 0)
tot = sum(f(r, S) for r in body)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter({k: sum(f(r, k) for r in body) for k in reasons})
print("samples", int(tot), " instr(warp)", int(sum(f(r, "Instructions Executed") for r in body)))
print("by reason:", [(k[6:], round(100 * v / tot, 1)) for k, v in agg.most_common(9)])
byop = collections.Counter()
cnt = collections.Counter()
for r in body:
    t = r[ix["Source"]].split()
    op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
    byop[op] += f(r, S)
    cnt[op] += f(r, "Instructions Executed")
print("by opcode (stall %, executed %):", [(k, round(100 * v / tot, 1), round(100 * cnt[k] / sum(cnt.values()), 1)) for k, v in byop.most_common(12)])
for r in sorted(body, key=lambda r: -f(r, S))[:ntop]:
    rs = sorted(((f(r, k), k[6:]) for k in reasons), reverse=True)[:2]
    print("%5s %-72s %5.2f%%  %s" % (r[ix["Address"]][-5:], r[ix["Source"]][:72], 100 * f(r, S) / tot, " ".join("%s=%d" % (k, v) for v, k in rs)))
