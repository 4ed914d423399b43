#!/usr/bin/env python
"""bench.py — site-alleles fitted / s (robust NB + p/q-values) on B200, with the CPU path beside it.

One "step" = one pass of the hot path (gate -> robust NB fit -> p/q/GQ, Holm power q-values, status,
strand bias, genotypes) over one synthetic readcount batch of the BASELINE.json configuration
configs[1]: 100 samples x 1 Mb of targets at ~100X (C2).  Region-sharded weak scaling: every rank
owns its own 1 Mb region (different seed), no data-path collective.

    value   device-resident: the count planes are already in HBM when the timed region starts
    e2e     through the public API with HOST (pinned) buffers: H2D of the planes and D2H of the
            results inside the timed region
    roofline      the fit kernel (k_fit) against the measured FP64 FMA peak of the same GPU
    roofline_gate the streaming gate kernel (k_gate) against the measured HBM copy bandwidth
    cpu_baseline  the CPU oracle (C restatement of the R reference; R itself is not installable
                  here) on a bounded sample of the same batch, all host cores

`--impl reference` times that CPU oracle alone (the reference arm of this tier).
"""
import argparse
import json
import shutil
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "site-alleles fitted/sec (robust NB + p/q-values)"
UNIT = "site-alleles/s"
REF_SAMPLE_POSITIONS = 8192  # positions of the batch the CPU arm samples (numpy generator, same bytes in both arms)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--positions", type=int, default=0, help="positions per rank per step (0 = the config's P)")
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline sample duration")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--rows-only", action="store_true",
                    help="end-to-end leg asks for the emitted rows only (per-unit output pointers NULL: no per-unit D2H)")
    ap.add_argument("--no-text", action="store_true", help="skip the text-table leg (loader and drop-in driver timings)")
    ap.add_argument("--text-tile", type=int, default=48, help="text-table leg: the 16000 distinct synthetic positions are "
                    "tiled this many times (48: 768000 positions x 100 samples, 2.1 GB of text)")
    ap.add_argument("--p-germline", type=float, default=1e-4)
    ap.add_argument("--dump-cycles", default="")
    ap.add_argument("--gpus-list", default="", help="--scaling strong: comma-separated device counts to run on the SAME "
                    "host batch, one JSON line each (e.g. 1,2,4,8); default: --gpus only")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank its own region (torchrun, one process per GPU); strong: ONE workload over "
                         "--gpus devices from ONE host process (ns_multi_call_snv), run without torchrun")
    return ap.parse_args()


# ---- synthetic batch on the GPU (same model as needlestack_b200.synth, torch RNG) ------------
def generate_torch(cfg, P, seed, device, p_germline=1e-4):
    import torch
    n = cfg.n
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f32 = torch.float32

    def rnd(*shape):
        return torch.rand(*shape, generator=g, device=device, dtype=f32)

    def gamma(shape_t):
        return torch._standard_gamma(shape_t, generator=g)

    scale = torch.exp(0.25 * torch.randn(n, generator=g, device=device, dtype=f32))
    mean_dp = cfg.depth * scale[None, :]
    lam = gamma(torch.full((P, n), 20.0, device=device, dtype=f32)) * (mean_dp / 20.0)
    DP = torch.poisson(lam, generator=g)
    del lam
    ref = torch.randint(0, 4, (P,), generator=g, device=device)
    room = DP.clone()
    ys = []
    var_pos = torch.nonzero(rnd(P) < 1e-3).flatten()
    var_k = torch.randint(0, 3, (len(var_pos),), generator=g, device=device)
    germ_pos = torch.nonzero(rnd(P) < p_germline).flatten()
    germ_k = torch.randint(0, 3, (len(germ_pos),), generator=g, device=device)
    for k in range(3):
        e = 10.0 ** (cfg.e_lo + (cfg.e_hi - cfg.e_lo) * rnd(P, 1))
        sg = 10.0 ** (cfg.s_lo + (cfg.s_hi - cfg.s_lo) * rnd(P, 1))
        shape = (1.0 / sg).expand(P, n).contiguous()
        lam_k = gamma(shape) * (e * DP * sg)
        y = torch.poisson(lam_k, generator=g)
        del lam_k, shape
        # planted variants in a few samples, common germline sites in ~30 %
        vp = var_pos[var_k == k]
        if len(vp):
            carriers = rnd(len(vp), n) < (1.5 / n)
            forced = torch.randint(0, n, (len(vp),), generator=g, device=device)
            carriers[torch.arange(len(vp), device=device), forced] = True
            if cfg.af_log:
                af = 10.0 ** (cfg.af_lo + (cfg.af_hi - cfg.af_lo) * rnd(len(vp), n))
            else:
                af = cfg.af_lo + (cfg.af_hi - cfg.af_lo) * rnd(len(vp), n)
            add = torch.binomial(DP[vp], af * carriers, generator=g)
            y[vp] += add
        gp = germ_pos[germ_k == k]
        if len(gp):
            carriers = rnd(len(gp), n) < 0.3
            af = torch.where(rnd(len(gp), n) < 0.5, 0.5, 1.0)
            y[gp] += torch.binomial(DP[gp], af * carriers, generator=g)
        y = torch.minimum(y, room)
        room -= y
        ys.append(y)
    counts = torch.zeros((P, 9, n), device=device, dtype=torch.int32)
    counts[:, 0, :] = DP.to(torch.int32)
    pidx = torch.arange(P, device=device)
    alts = torch.tensor([[b for b in range(4) if b != r] for r in range(4)], device=device)
    half = torch.full((1,), 0.5, device=device, dtype=f32)
    for k in range(3):
        b = alts[ref, k]
        fwd = torch.binomial(ys[k], half.expand_as(ys[k]), generator=g)
        counts[pidx, 1 + b, :] = fwd.to(torch.int32)
        counts[pidx, 5 + b, :] = (ys[k] - fwd).to(torch.int32)
    rf = torch.binomial(room, half.expand_as(room), generator=g)
    counts[pidx, 1 + ref, :] = rf.to(torch.int32)
    counts[pidx, 5 + ref, :] = (room - rf).to(torch.int32)
    return ref.to(torch.uint8).contiguous(), counts.contiguous()


# ---- clocks during the timed region ---------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ---- the CPU arm --------------------------------------------------------------------------
def cpu_oracle_rate(ref_np, counts_np, n_threads, target_s, prm_kw):
    """fitted site-alleles / s of the CPU oracle on the first positions of the batch (bounded sample)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    orc.lib()
    prm = orc.call_params(**prm_kw)

    def run(Ps):
        c = np.ascontiguousarray(counts_np[:Ps].transpose(1, 0, 2))
        t = time.perf_counter()
        O = orc.call_snv_batch(ref_np[:Ps], c, prm, n_threads=n_threads)
        dt = time.perf_counter() - t
        return dt, int((O["gated"] == 0).sum()), O

    Pmax = len(ref_np)
    Ps = min(Pmax, 64 * n_threads)
    dt, nf, O = run(Ps)
    if dt < target_s / 3 and Ps < Pmax:
        Ps2 = int(min(Pmax, max(Ps, Ps * target_s / max(dt, 1e-3))))
        Ps2 -= Ps2 % n_threads
        if Ps2 > Ps:
            Ps = Ps2
            dt, nf, O = run(Ps)
    return dict(value=nf / dt, fitted=nf, seconds=dt, positions=Ps, oracle=O)


def text_leg(cfg, device, positions=16000, tile=12):
    """The path either side of the kernels, timed on a C2-shaped TEXT table ("text-table parsing is timed
    separately"): the columnar loader alone (ns_table_parse) and the whole drop-in driver - table text in, VCF text out
    (block reader -> loader -> ns_call_snv -> C++ VCF writer, three stages overlapped).  The table holds `positions`
    distinct synthetic positions written once and tiled `tile` times (writing 3.3 KB lines from Python is the slow part)."""
    import io
    import tempfile
    from needlestack_b200 import driver, readcounts, synth
    n = cfg.n
    ref, counts = synth.generate(cfg, P=positions, seed=5)
    cwd = os.getcwd()
    d = tempfile.mkdtemp()
    try:
        os.chdir(d)
        synth.write_table("t.tsv", ref, counts)
        with open("names.txt", "w") as f:
            f.write("".join(f"bam{i} S{i + 1}\n" for i in range(n)))
        raw = open("t.tsv", "rb").read()
        nl = raw.index(b"\n") + 1
        text = raw[:nl] + raw[nl:] * tile
        with open("tiled.tsv", "wb") as f:  # the driver reads a FILE on its STDIN (mmap-ed, parsed in place)
            f.write(text)
        P = positions * tile
        best = None
        for _ in range(3):
            t = time.perf_counter()
            T = readcounts.load_table(text, n)
            dt = time.perf_counter() - t
            T.close()
            best = dt if best is None else min(best, dt)
        runs = []
        nrec = 0
        for _ in range(3):
            with open("tiled.tsv", "rb") as f:
                t = time.perf_counter()
                driver.main(["--out_file=out.vcf", f"--device={device}"], stdin=f)
                runs.append(time.perf_counter() - t)
        with open("out.vcf", "rb") as f:
            nrec = sum(1 for ln in f if not ln.startswith(b"#"))
        dt = min(runs[1:])
        return {"table": f"{cfg.name}-shaped text, {n} samples x {P} positions ({positions} distinct, tiled x{tile}), "
                         f"{len(text) / 1e9:.2f} GB", "loader_gbs": len(text) / 1e9 / best, "loader_s": best,
                "loader_threads": os.cpu_count(), "driver_s": dt, "driver_gbs": len(text) / 1e9 / dt,
                "driver_site_alleles_per_s": 3 * P / dt, "vcf_records": nrec,
                "driver": "python -m needlestack_b200.driver < table (mmap-ed input parsed in place, C++ loader, ns_call_snv, "
                          "C++ VCF writer; reader / GPU / writer stages overlapped)"}
    finally:
        os.chdir(cwd)
        import shutil
        shutil.rmtree(d, ignore_errors=True)


def parity_check(caller, ns, ref_np, counts_np, O, Ps):
    """The CUDA path (through the C ABI, host buffers) against the oracle's outputs on the first Ps positions of
    the BENCHMARKED batch, outside every timed region: ties the headline number to a correctness statement."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    R = caller.call_snv(ref_np[:Ps], counts_np[:Ps], prm=ns.make_params(emit_mode=ns.NS_EMIT_ALL))
    d = {k: R.dense(k) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "pooled")}
    d["gt"] = R.dense("gt", fill=0)
    d["status"] = R.dense("status", fill=0)
    for k in ("sigma", "slope", "max_gq", "flags", "n_outer", "n_irls", "n_obj"):
        d[k] = np.array(R[k])
    try:
        rep = parity.compare(O, d, label="bench batch")
        bad = 0
    except AssertionError as e:
        rep = {"error": str(e)[:600]}
        bad = -1
    keep = ("n_units", "n_fitted", "calls", "calls_compared", "nonconverged", "nonconverged_compared", "gate_mismatch",
            "gt_mismatch", "status_mismatch", "maxit_mismatch", "iter_mismatch", "irls_mismatch", "obj_mismatch",
            "near_threshold", "sigma_abs_rule_only", "sigma_within_tol_only", "max_rel_sigma", "max_abs_sigma",
            "max_rel_slope", "max_d_gq", "max_rel_p", "error")
    out = {k: rep[k] for k in keep if k in rep}
    out["units"] = out.pop("n_units", 3 * Ps)
    out["fitted"] = out.pop("n_fitted", None)
    out["mismatches"] = bad if bad else int(rep["gate_mismatch"] + rep["gt_mismatch"] + rep["status_mismatch"] +
                                             rep["maxit_mismatch"])
    out["positions"] = int(Ps)
    out["tolerance"] = "flags/GT/STATUS bit-exact; sigma, slope rel 1e-6 (sigma abs 2e-8 / 1e-6 counted); p, q, Phred 1e-6"
    return out


def strong_main(a, cfg, P, n, config, torch, ns):
    """ONE workload (P positions) over a.gpus devices from this one host process: the planes live in page-locked host
    memory, ns_multi_call_snv hands chunks of positions to one host thread per device and gathers the rows in
    genomic order.  Every step is a whole call (H2D, kernels, D2H, gather inside the timed region)."""
    nd = torch.cuda.device_count()
    want = max([int(x) for x in a.gpus_list.split(",") if x] or [a.gpus])
    if want > nd:
        raise SystemExit(f"bench.py --scaling strong: {want} devices asked for, only {nd} visible")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    h_counts = ns.PinnedArray((P, 9, n), np.int32)
    h_ref = ns.PinnedArray((P,), np.uint8)
    # the batch is generated on device 0 in slices (a 10 Mb x 400 sample batch does not fit next to torch's temporaries)
    slice_p = max(1, min(P, int(2.0e9 // (36 * n))))
    Ph = min(P, REF_SAMPLE_POSITIONS)
    for k, p0 in enumerate(range(0, P, slice_p)):
        pc = min(slice_p, P - p0)
        d_ref, d_counts = generate_torch(cfg, pc, cfg.seed + 7919 * k, dev, a.p_germline)
        torch.from_numpy(h_counts.array[p0:p0 + pc]).copy_(d_counts)
        torch.from_numpy(h_ref.array[p0:p0 + pc]).copy_(d_ref)
        del d_ref, d_counts
    from needlestack_b200 import synth
    h_ref.array[:Ph], h_counts.array[:Ph] = synth.generate(cfg, P=Ph, seed=cfg.seed, p_germline=a.p_germline)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    prm = ns.make_params()
    counts_list = [int(x) for x in a.gpus_list.split(",") if x] or [a.gpus]
    base_ms = None
    for ng in counts_list:
        line = strong_one(a, cfg, P, n, dict(config), ns, prm, h_ref, h_counts, ng, check=(ng == counts_list[-1]))
        if base_ms is None:
            base_ms = (ng, line["ms_per_step"])
        line["speedup_vs_first"] = {"n_gpus_first": base_ms[0], "speedup": base_ms[1] / line["ms_per_step"]}
        print(json.dumps(line), flush=True)
    return 0


def strong_one(a, cfg, P, n, config, ns, prm, h_ref, h_counts, ng, check=True):
    m = ns.MultiCaller(list(range(ng)), reuse_outputs=True)
    for _ in range(max(a.warmup, 1)):
        R = m.call_snv(h_ref.array, h_counts.array, prm=prm)
    samplers = [ClockSampler(i) for i in range(ng)]
    for sp in samplers:
        sp.start()
    fitted = 0
    per_dev = [dict(chunks=0, wall_ms=0.0, device_ms=0.0, gate_ms=0.0, fit_ms=0.0, pool_tail_ms=0.0) for _ in range(ng)]
    launches = 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        R = m.call_snv(h_ref.array, h_counts.array, prm=prm)
        fitted += R.n_fitted
        for d, t in zip(per_dev, m.timings()):
            d["chunks"] += t["chunks"]
            d["wall_ms"] += t["wall_ms"]
            d["device_ms"] += t["total_ms"]
            d["gate_ms"] += t["gate_ms"]
            d["fit_ms"] += t["fit_ms"]
            d["pool_tail_ms"] += t["pool_tail_ms"]
            launches += t["kernel_launches"]
    ms = (time.perf_counter() - t0) * 1e3
    clocks = [sp.stop() for sp in samplers]
    for d in per_dev:
        for k in d:
            d[k] = d[k] / a.steps
    U = 3 * P
    d2h = U * (8 + 8 + 8 + 4 + 4 + 4) + R.n_emitted * (n * 58 + 8 * 8 + 8)
    value = fitted / (ms * 1e-3)
    # single-context result of the same call: the gather must not change anything
    same = None
    if ng > 1 and check:
        with ns.Caller(0) as c1:
            R1 = c1.call_snv(h_ref.array, h_counts.array, prm=prm)
        same = bool(R1.n_emitted == R.n_emitted and R1.n_fitted == R.n_fitted and np.array_equal(R1.emit_unit, R.emit_unit)
                    and np.array_equal(R1.gq, R.gq, equal_nan=True) and np.array_equal(R1.sigma, R.sigma, equal_nan=True))
    config = dict(config, parallelism=f"one host process, {ng} devices, queue of position chunks (ns_multi_call_snv)",
                  positions_total=P,
                  workload=f"{cfg.name}: {n} samples x {P} positions IN TOTAL at ~{cfg.depth:g}X, SNV alleles (3 per position), "
                           f"one table split over {ng} GPU(s)")
    config.pop("positions_per_gpu", None)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ng, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": {"sm_mhz": min((c["sm_mhz"] or 0) for c in clocks), "sm_max_mhz": max((c["sm_max_mhz"] or 0) for c in clocks),
                       "reasons": sorted(set(r for c in clocks for r in c["reasons"]))},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(h_counts.nbytes + h_ref.nbytes),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms / a.steps,
                    "api": "needlestack_b200.MultiCaller.call_snv -> ns_multi_call_snv (C ABI), host planes in, host results out"},
            "value_note": "one workload from host memory: `value` is the end-to-end rate (there is no device-resident form of a "
                          "multi-device call); timed by the host clock around the synchronous calls",
            "gpu_launches": int(launches), "per_device": per_dev,
            "detail": {"fitted_per_step": R.n_fitted, "emitted_per_step": R.n_emitted,
                       "identical_to_single_device_call": same,
                       "site_alleles_processed_per_s": U * a.steps / (ms * 1e-3)}}
    m.close()
    return line


def main():
    a = parse()
    from needlestack_b200 import synth
    cfg = synth.CONFIGS[a.config]
    if a.samples:
        cfg = synth.Config(**{**cfg.__dict__, "n": a.samples})
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    P = a.positions or cfg.P
    n = cfg.n
    workload = f"{a.config}: {n} samples x {P} positions/GPU at ~{cfg.depth:g}X, SNV alleles (3 per position)"
    config = {"workload": workload, "samples": n, "positions_per_gpu": P, "depth": cfg.depth,
              "parallelism": f"region-sharded x{world}", "l2": "inputs (%.2f GB/GPU) larger than L2" %
              (P * 9 * n * 4 / 1e9), "params": "needlestack.r defaults (min_coverage 30, min_reads 3, GQ 50, SOR 100)"}
    ncores = os.cpu_count() or 1

    if a.impl == "reference":
        # the reference arm of this tier: the CPU restatement of the R path (R is not installable
        # here), all host cores, bounded sample of the same workload per step
        if rank != 0:
            return 0
        Ps = min(P, REF_SAMPLE_POSITIONS)
        ref_np, counts_np = synth.generate(cfg, P=Ps, seed=cfg.seed, p_germline=a.p_germline)
        rates = []
        t_all = time.perf_counter()
        per = max(2.0, min(20.0, 120.0 / max(1, a.steps + a.warmup)))
        res = None
        for i in range(a.warmup + a.steps):
            res = cpu_oracle_rate(ref_np, counts_np, ncores, per, {})
            if i >= a.warmup:
                rates.append(res)
        fitted = sum(r["fitted"] for r in rates)
        secs = sum(r["seconds"] for r in rates)
        val = fitted / secs
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / max(1, a.steps),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port",
                                 "sample": f"{res['positions']} positions ({res['fitted']} fitted site-alleles) "
                                           f"per step = the first positions of the b200 arm's rank-0 batch, C oracle (restatement of glm_rob_nb.r/needlestack.r; R "
                                           f"absent), {ncores} pthreads",
                                 "rscript_on_path": shutil.which("Rscript")},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import needlestack_b200 as ns
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    if a.scaling == "strong":
        if rank != 0:
            return 0
        return strong_main(a, cfg, P, n, config, torch, ns)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # stdout carries the one JSON line: NCCL's log (version banner, rank/ring lines at NCCL_DEBUG=INFO) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def allgather(x):
        if dist is None:
            return [x]
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    d_ref, d_counts = generate_torch(cfg, P, cfg.seed + 1000 * rank, dev, a.p_germline)
    # the first positions of every rank's batch come from the numpy generator: for rank 0 they are exactly the sample the
    # reference arm (--impl reference) and the cpu_baseline / parity_check legs run the CPU oracle on
    Ph = min(P, REF_SAMPLE_POSITIONS)
    h_ref0, h_counts0 = synth.generate(cfg, P=Ph, seed=cfg.seed + 1000 * rank, p_germline=a.p_germline)
    d_ref[:Ph] = torch.from_numpy(h_ref0).to(dev)
    d_counts[:Ph] = torch.from_numpy(h_counts0).to(dev)
    del h_ref0, h_counts0
    torch.cuda.synchronize()
    caller = ns.Caller(local_rank, reuse_outputs=True)  # results land in the Caller's pinned buffers
    caller.set_stream(torch.cuda.current_stream().cuda_stream)
    prm = ns.make_params()
    fp64_peak = caller.fp64_peak_tflops()

    def step_device():
        return caller.call_snv_device(d_ref.data_ptr(), d_counts.data_ptr(), P, n, prm, per_unit=not a.rows_only)

    # ---- device-resident leg ----
    for _ in range(a.warmup):
        R = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tm_sum = {}
    launches = 0
    fitted = 0
    seeds = lattice = quad = 0
    e0.record()
    for _ in range(a.steps):
        R = step_device()
        tm = caller.timings()
        for k, v in tm.items():
            tm_sum[k] = tm_sum.get(k, 0) + v
        fitted += R.n_fitted
        seeds += R.n_seed
        lattice += R.n_lattice
        quad += R.n_quad
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if a.dump_cycles and rank == 0:
        np.savez_compressed(a.dump_cycles, cycles=R.unit_cycles, iters=R.iters, flags=R.flags, sigma=R.sigma,
                            slope=R.slope)
    clocks = sampler.stop()
    ms_max = allmax(ms)
    per_rank_ms = allgather(ms / a.steps)
    fitted_all = allsum(fitted)
    launches = int(tm_sum.get("kernel_launches", 0))
    value = fitted_all / (ms_max * 1e-3)
    units_all = allsum(3 * P * a.steps)
    n_emitted = R.n_emitted

    # ---- rooflines (rank 0's kernels) ----
    fit_s = tm_sum["fit_ms"] * 1e-3
    fseeds, flattice = tm_sum["fast_seed"], tm_sum["fast_lattice"]
    flops = 400.0 * fseeds + 32.0 * flattice
    ach = flops / fit_s / 1e12 if fit_s > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (burst)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    gate_s = tm_sum["gate_ms"] * 1e-3
    gate_bytes = a.steps * (P * (36.0 * n + 1) + 3 * P * 44.0)
    gate_ach = gate_bytes / gate_s / 1e9 if gate_s > 0 else 0.0
    # DRAM bytes per launch: NOT measured in this run (that needs ncu) - read from the committed ncu capture of this same
    # workload (profiles/capture.sh step 3) and labelled as such
    traffic = {}
    traffic_src = None
    if a.config == "C2" and P == cfg.P and n == cfg.n:
        try:
            tj = sorted((f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("traffic_r")),
                        key=lambda f: os.path.getmtime(os.path.join(ROOT, "profiles", f)))[-1]
            for k, v in json.load(open(os.path.join(ROOT, "profiles", tj)))["kernels"].items():
                traffic[k] = v["dram_bytes_per_launch"]
            traffic["k_gate"] = traffic.get("k_gate_v4", traffic.get("k_gate"))
            traffic_src = f"committed ncu capture profiles/{tj} (same command, earlier run), not this run"
        except Exception:
            traffic = {}
    roofline = {"kernel": "k_fit", "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach / fp64_peak if fp64_peak else None, "traffic": traffic.get("k_fit"), "traffic_source": traffic_src,
                "peak_datasheet": 37.2, "peak_vs_datasheet": fp64_peak / 37.2 if fp64_peak else None,
                "peak_source": "DFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
                "flops_model": "400*N_seed + 32*N_lattice (+ 2000*N_quad, zero in this kernel) (SURVEY.md 8d); "
                               "counters from the kernel",
                "n_seed": fseeds, "n_lattice": flattice, "all_kernels": {"n_seed": seeds, "n_lattice": lattice,
                                                                       "n_quad": quad},
                "share_of_step": tm_sum["fit_ms"] / ms if ms else None}
    roofline_gate = {"kernel": "k_gate_v4 / k_gate_v2 / k_gate (by sample count)", "bound": "hbm", "achieved": gate_ach, "peak": hbm_peak, "unit": "GB/s",
                     "frac": gate_ach / hbm_peak, "traffic": traffic.get("k_gate"), "traffic_source": traffic_src,
                     "peak_source": hbm_src,
                     "algorithmic_bytes_per_launch": gate_bytes / (a.steps * -(-P // 262144)),
                     "launches_per_step": -(-P // 262144),
                     "bytes_model": "36*n+1 B read per position + 44 B written per unit",
                     "share_of_step": tm_sum["gate_ms"] / ms if ms else None}

    # ---- end-to-end leg: host (pinned) planes in, host results out ----
    e2e = None
    if not a.no_e2e:
        h_counts = ns.PinnedArray((P, 9, n), np.int32)
        h_ref = ns.PinnedArray((P,), np.uint8)
        torch.cuda.synchronize()
        torch.from_numpy(h_counts.array).copy_(d_counts)
        torch.from_numpy(h_ref.array).copy_(d_ref)
        torch.cuda.synchronize()

        def step_host():
            return caller.call_snv(h_ref.array, h_counts.array, prm=prm, per_unit=not a.rows_only)

        for _ in range(min(a.warmup, 3)):
            Rh = step_host()
        barrier()
        fitted_h = 0
        e0.record()
        for _ in range(a.steps):
            Rh = step_host()
            fitted_h += Rh.n_fitted
        e1.record()
        tm_h = caller.timings()
        barrier()
        ms_h_own = e0.elapsed_time(e1)
        ms_h = allmax(ms_h_own)
        per_rank_ms_e2e = allgather(ms_h_own / a.steps)
        U = 3 * P
        d2h = (0 if a.rows_only else U * (8 + 8 + 8 + 4 + 4 + 4 + 8)) + Rh.n_emitted * (n * 58 + 8 * 8 + 8 + 28)
        e2e = {"value": allsum(fitted_h) / (ms_h * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h_counts.nbytes + h_ref.nbytes), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_h / a.steps, "per_rank_ms": per_rank_ms_e2e,
               "kernel_ms_last_call_rank0": {k: v for k, v in tm_h.items() if k.endswith("_ms")},
               "api": "needlestack_b200.Caller.call_snv -> ns_call_snv (C ABI)" + (", emitted rows only" if a.rows_only else "")}
        assert Rh.n_fitted == R.n_fitted and Rh.n_emitted == R.n_emitted
        ref_np = h_ref.array[: min(P, REF_SAMPLE_POSITIONS)].copy()
        counts_np = h_counts.array[: min(P, REF_SAMPLE_POSITIONS)].copy()
        h_counts.free()  # at most one set of page-locked planes per rank at a time
        if int(d_counts.max().item()) < 65536:
            # informational: the same call with the 16-bit wire format (ns_call_snv_u16); the headline e2e above is the
            # reference-facing int32 form
            h16 = ns.PinnedArray((P, 9, n), np.uint16)
            torch.from_numpy(h16.array.view(np.int16)).copy_(d_counts.to(torch.int16))
            torch.cuda.synchronize()
            for _ in range(2):
                R16 = caller.call_snv(h_ref.array, h16.array, prm=prm, per_unit=not a.rows_only)
            barrier()
            e0.record()
            for _ in range(a.steps):
                R16 = caller.call_snv(h_ref.array, h16.array, prm=prm, per_unit=not a.rows_only)
            e1.record()
            barrier()
            ms16 = allmax(e0.elapsed_time(e1))
            assert R16.n_fitted == R.n_fitted and R16.n_emitted == R.n_emitted
            e2e["u16_wire_format"] = {"value": allsum(R16.n_fitted * a.steps) / (ms16 * 1e-3), "ms_per_step": ms16 / a.steps,
                                      "h2d_bytes_per_step": int(h16.nbytes + h_ref.nbytes)}
            h16.free()
    else:
        ref_np = d_ref[: min(P, REF_SAMPLE_POSITIONS)].cpu().numpy()
        counts_np = d_counts[: min(P, REF_SAMPLE_POSITIONS)].cpu().numpy()

    cpu = None
    parity_rep = None
    if rank == 0 and world == 1 and not a.no_cpu:
        r = cpu_oracle_rate(ref_np, counts_np, ncores, a.cpu_seconds, {})
        cpu = {"value": r["value"], "unit": UNIT, "cores": ncores, "kind": "port",
               "sample": f"first {r['positions']} positions of the same batch ({r['fitted']} fitted site-alleles, "
                         f"{r['seconds']:.1f} s), C oracle (restatement of glm_rob_nb.r/needlestack.r; R is not "
                         f"installed on this image), {ncores} pthreads"}
        parity_rep = parity_check(caller, ns, ref_np, counts_np, r["oracle"], r["positions"])
    text = None
    if rank == 0 and world == 1 and not a.no_text and not a.no_cpu:
        try:
            text = text_leg(cfg, local_rank, tile=max(a.text_tile, 1))
        except Exception as e:  # the headline does not depend on this leg
            text = {"error": repr(e)[:300]}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_gate": roofline_gate,
                "cpu_baseline": cpu, "parity_check": parity_rep,
                "per_rank_ms": per_rank_ms,
                "detail": {"site_alleles_processed_per_s": units_all / (ms_max * 1e-3), "text": text,
                           "pool": {"units_per_step_rank0": tm_sum.get("pool_units", 0) / a.steps,
                                    "tail_ms_per_step_rank0": tm_sum.get("pool_tail_ms", 0) / a.steps,
                                    "chunks_per_step_rank0": tm_sum.get("chunks", 0) / a.steps},
                           "gate_pass_fraction": fitted_all / units_all, "fitted_per_step_rank0": R.n_fitted,
                           "emitted_per_step_rank0": n_emitted,
                           "heavy_units_per_step_rank0": tm_sum.get("n_heavy", 0) / a.steps,
                           "heavy_units_not_predicted_per_step_rank0": tm_sum.get("n_deferred", 0) / a.steps,
                           "kernel_ms_per_step_rank0": {k: v / a.steps for k, v in tm_sum.items()
                                                        if k.endswith("_ms")}}}
        print(json.dumps(line))
    caller.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
