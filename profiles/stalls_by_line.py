"""Warp-stall samples of one kernel of an .ncu-rep, summed per CUDA source line (ncu --page source --print-source cuda,sass).
Usage: python profiles/stalls_by_line.py <report.ncu-rep> <kernel regex> [top N]   (the report needs --import-source on and -lineinfo)"""
import collections
import csv
import io
import subprocess
import sys


def I(v):
    try:
        return int(v)
    except (TypeError, ValueError):
        return 0


rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur, hdr, agg = None, None, []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = {}
        for i, h in enumerate(hdr):
            if i < len(r) and h not in d:
                d[h] = r[i]
        agg.append((cur, int(r[0]), d))
stalls = [h for h in dict.fromkeys(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(I(d["# Samples"]) for _, _, d in agg)
tot_inst = sum(I(d["Instructions Executed"]) for _, _, d in agg)
by_reason = collections.Counter()
by_file = collections.Counter()
for f, _, d in agg:
    by_file[f] += I(d["# Samples"])
    for h in stalls:
        by_reason[h[6:]] += I(d.get(h))
sr = sum(by_reason.values())
print(f"kernel {kern}: {tot} warp samples, {tot_inst} warp instructions")
print("stall reasons: " + ", ".join(f"{k} {100 * v / sr:.1f}%" for k, v in by_reason.most_common(9)))
print("by file: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in by_file.most_common(8)))
agg.sort(key=lambda x: -I(x[2]["# Samples"]))
for f, ln, d in agg[:top]:
    s = I(d["# Samples"])
    st = sorted(((I(d.get(h)), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{f}:{ln:<5d} {100 * s / tot:5.1f}% of samples, {100 * I(d['Instructions Executed']) / tot_inst:4.1f}% of instructions  "
          + ", ".join(f"{b} {a}" for a, b in st if a) + "   | " + d["Source"].strip()[:100])
