"""Turns the raw captures of profiles/capture.sh (in gpurun_out/) into the small tracked summaries under profiles/.
Usage: python profiles/summarize.py <tag>"""
import csv
import json
import os
import re
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
SRC, DST = "gpurun_out", "profiles"
OURS = ("k_gate_staged", "k_gate_v4", "k_gate_v2", "k_gate", "k_pool_gather", "k_widen16", "k_fit_heavy_tp", "k_fit_heavy", "k_fit", "k_post", "k_pack", "k_qtab", "k_dfma", "k_scan_total", "k_qgrid", "DeviceScan")


def short(name):
    for k in OURS:
        if re.search(r"\b%s\b" % k, name):
            return k
    return None


def rows(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.DictReader(lines))


# launch list -> per-kernel launches / total ms / share (our kernels only)
p = os.path.join(SRC, "launches_%s.csv" % TAG)
if os.path.exists(p):
    agg = {}
    for r in rows(p):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        if not k:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(DST, "launch_summary_%s.csv" % TAG), "w") as f:
        f.write("kernel,launches,total_ms,share\n")
        for k, (c, ms) in agg.items():
            f.write("%s,%d,%.3f,%.4f\n" % (k, c, ms, ms / tot))
    subprocess.run("cp %s %s" % (p, os.path.join(DST, "launches_%s.csv" % TAG)), shell=True, check=True)

# DRAM traffic per launch of the roofline kernels
p = os.path.join(SRC, "traffic_%s.csv" % TAG)
if os.path.exists(p):
    per = {}
    for r in rows(p):
        k = short(r["Kernel Name"])
        if not k:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            per.setdefault(k, {}).setdefault(r["ID"], [0.0, 0.0])[0] += v
        elif r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
            per.setdefault(k, {}).setdefault(r["ID"], [0.0, 0.0])[1] += v
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of `python bench.py --steps 2 --warmup 1 --no-cpu "
                     "--no-e2e` (C2, one launch = one chunk of 262144 positions)", "kernels": {}}
    for k, d in per.items():
        vals = list(d.values())
        out["kernels"][k] = {"launches": len(vals), "dram_bytes_per_launch": sum(v[0] for v in vals) / len(vals),
                             "ms_per_launch_under_ncu": sum(v[1] for v in vals) / len(vals)}
    json.dump(out, open(os.path.join(DST, "traffic_%s.json" % TAG), "w"), indent=1)

# full-set report -> details text + compact json
rep = os.path.join(SRC, "prof_%s_full.ncu-rep" % TAG)
if os.path.exists(rep):
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(DST, "ncu_full_details_%s.txt" % TAG), "w").write(det)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    keep = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")
    summ = []
    for r in rr[2:]:
        d = dict(zip(hdr, r))
        e = {"kernel": short(d.get("Kernel Name", "")) or d.get("Kernel Name", "")[:40]}
        for k in keep:
            if k in d:
                e[k] = "%s %s" % (d[k], units[hdr.index(k)])
        summ.append(e)
    json.dump(summ, open(os.path.join(DST, "ncu_full_summary_%s.json" % TAG), "w"), indent=1)
for f in ("bench_%s_full.json" % TAG, "bench_%s_short.json" % TAG):
    if os.path.exists(os.path.join(SRC, f)):
        subprocess.run("cp %s %s" % (os.path.join(SRC, f), os.path.join(DST, f)), shell=True, check=True)
print("summaries written to", DST)
