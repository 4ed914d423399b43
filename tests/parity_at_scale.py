"""Larger parity sweep (manual: `python tests/parity_at_scale.py [out.json]` on a GPU box): the CUDA path against
the CPU oracle on every BASELINE shape at sizes the oracle finishes in about a minute each.  Writes one JSON
report (profiles/parity_r1.json)."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402
import parity  # noqa: E402
import needlestack_b200 as ns  # noqa: E402
from needlestack_b200 import synth  # noqa: E402

# (config, positions, caller parameters, tumour/normal pairs?, p_variant, p_germline)
CASES = [("C1", 40000, {}, False, 1e-3, 1e-4), ("C1", 8000, {}, False, 0.03, 0.005),
         ("C2", 16000, {}, False, 1e-3, 1e-4), ("C2", 3000, {}, False, 0.03, 0.003),
         ("C2", 4000, {"extra_rob": 1}, False, 0.01, 0.01),
         ("C3", 48, {}, False, 1e-3, 1e-4), ("C3", 32, {}, False, 0.3, 0.0),
         ("C4", 1500, {}, True, 1e-3, 1e-4), ("C4", 600, {}, True, 0.1, 0.05),
         ("C5", 300000, {}, False, 1e-3, 1e-4),
         # the throughput regime of the cluster kernel (one CTA per unit, two-phase evaluation), forced on a small batch
         ("C3", 48, {"_env": {"NS_HEAVY_CLUSTER": "1", "NS_HEAVY_THREADS": "640"}}, False, 1e-3, 1e-4)]


def main(out_path):
    rep_all = []
    threads = os.cpu_count() or 1
    with ns.Caller(0) as c:
        for cfgname, P, kw, tn, pv, pg in CASES:
            kw = dict(kw)
            env = kw.pop("_env", {})
            os.environ.update(env)
            cfg = synth.CONFIGS[cfgname]
            ref, counts = synth.generate(cfg, P=P, seed=cfg.seed + 7, p_variant=pv, p_germline=pg)
            pairs = synth.tn_pairs(cfg) if tn else None
            t = time.time()
            O = orc.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)),
                                   orc.call_params(pairs=pairs, **kw), n_threads=threads)
            t_or = time.time() - t
            t = time.time()
            R = c.call_snv(ref, counts, prm=ns.make_params(pairs=pairs, emit_mode=ns.NS_EMIT_ALL, **kw))
            t_gpu = time.time() - t
            d = {k: R.dense(k) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "pooled")}
            d["gt"] = R.dense("gt", fill=0)
            d["status"] = R.dense("status", fill=0)
            for k in ("sigma", "slope", "max_gq", "flags", "n_outer", "n_irls", "n_obj"):
                d[k] = R[k]
            try:
                rep = parity.compare(O, d, label=cfgname)
                rep["ok"] = True
            except AssertionError as e:
                rep = dict(ok=False, error=str(e)[:2000])
            rep.pop("bad", None)
            for k in env:
                os.environ.pop(k, None)
            rep.update(config=cfgname, positions=P, p_variant=pv, p_germline=pg, params=kw, env=env, tn_pairs=tn, oracle_seconds=round(t_or, 2),
                       oracle_threads=threads, gpu_seconds=round(t_gpu, 3), calls=int(O["gate2"].sum()),
                       heavy_units=int(c.timings()["n_heavy"]))
            rep = {k: (float(v) if isinstance(v, (np.floating,)) else v) for k, v in rep.items()}
            print(json.dumps(rep))
            rep_all.append(rep)
    tot = {k: int(sum(r.get(k, 0) for r in rep_all)) for k in ("n_units", "n_fitted", "calls", "calls_compared", "tn_calls",
                                                              "nonconverged", "nonconverged_compared", "calls_nonconverged",
                                                              "gt_mismatch", "status_mismatch", "gate_mismatch",
                                                              "maxit_mismatch", "iter_mismatch", "irls_mismatch",
                                                              "obj_mismatch", "near_threshold", "sigma_abs_rule_only",
                                                              "sigma_within_tol_only")}
    tot["all_ok"] = all(r.get("ok") for r in rep_all)
    print(json.dumps({"total": tot}))
    json.dump({"cases": rep_all, "total": tot}, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "parity_report.json")
