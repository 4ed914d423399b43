"""Regenerates tests/golden/pyref_units.npz: >= 1000 glmrob.nb() units (inputs + the outputs of the INDEPENDENT
scipy/numpy transcription tests/pyref.py).  tests/test_pyref_vs_oracle.py checks the C oracle against these stored
outputs (seconds) and re-runs pyref live on a few of them.  pyref shares no numerics with oracle/orc_math.c (boost
negative binomial, scipy's QUADPACK, scipy digamma), so agreement far below the parity tolerance is the closest
thing to a pin of the oracle that exists without R (SURVEY.md 8c, item v).

    python tests/golden/make_pyref_fixture.py [n_procs]        (about 10 minutes on 8 cores)

Unit classes (selected with the C oracle's own flags, then re-computed by pyref with no knowledge of them):
plain converged fits of the C1 / C2 / C5 shapes, units that stop on maxit = 30, units on the spline + integrate()
branch, extra-robust units (samples removed before the fit), reference-inverted units (the regression runs on
the REF counts, needlestack.r:330-337)."""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
import parity  # noqa: E402
import pyref  # noqa: E402
from needlestack_b200 import synth  # noqa: E402

# batches to draw units from: (config, positions, seed, p_variant, p_germline, p_refswap, extra_rob, n override,
# depth override, {class: how many units to keep})
BATCHES = [("C1", 12000, 501, 0.02, 0.01, 0.01, 0, None, None, dict(plain=400, maxit=60, spline=30, ref_inv=30)),
           ("C1", 3000, 502, 0.02, 0.03, 0.0, 1, None, None, dict(extra_rob=60, plain=60)),
           ("C2", 1200, 503, 0.02, 0.02, 0.02, 0, None, None, dict(plain=150, maxit=25, spline=8, ref_inv=8)),
           ("C5", 16000, 504, 0.01, 0.01, 0.01, 0, None, None, dict(plain=150, maxit=30, spline=10, ref_inv=10)),
           ("C1", 40, 505, 0.3, 0.0, 0.0, 0, 24, 4000.0, dict(spline=30)),
           ("C2", 300, 506, 0.02, 0.05, 0.03, 1, 40, None, dict(extra_rob=40, plain=40, ref_inv=10)),
           ("C1", 1500, 507, 0.01, 0.0, 0.5, 0, None, None, dict(ref_inv=40)),
           ("C2", 200, 508, 0.01, 0.0, 0.5, 0, None, None, dict(ref_inv=12))]
CLASSES = ("extra_rob", "maxit", "plain", "ref_inv", "spline")


def classify(O, u):
    """a unit is filed under the first class that applies"""
    if O["ref_inv"][u]:
        return "ref_inv"
    if O["n_quad"][u] > 0:
        return "spline"
    if O["extra_rob"][u]:
        return "extra_rob"
    if O["maxit_hit"][u]:
        return "maxit"
    return "plain"


def collect():
    units = []
    have = {k: 0 for k in CLASSES}
    for cfgname, P, seed, pv, pg, prs, er, n_over, d_over, quota in BATCHES:
        cfg = synth.CONFIGS[cfgname]
        if n_over or d_over:
            cfg = synth.Config(**{**cfg.__dict__, "n": n_over or cfg.n, "depth": d_over or cfg.depth})
        ref, counts = synth.generate(cfg, P=P, seed=seed, p_variant=pv, p_germline=pg, p_refswap=prs)
        O = orc.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), orc.call_params(extra_rob=er),
                               n_threads=os.cpu_count() or 1)
        got = {k: 0 for k in CLASSES}
        for u in np.where(O["gated"] == 0)[0]:
            cls = classify(O, u)
            if got[cls] >= quota.get(cls, 0):
                continue
            dp, vp, vm, rp, rm = parity.snv_unit_rows(ref, counts, int(u))
            y = (rp + rm) if O["ref_inv"][u] else (vp + vm)
            units.append(dict(y=y.astype(np.int32), x=dp.astype(np.int32), extra_rob=er, cls=cls, batch=cfgname))
            got[cls] += 1
            have[cls] += 1
    return units, have


def run_one(t):
    y, x, er = t
    R = pyref.glmrob_nb(y.astype(float), x.astype(float), min_coverage=30, min_reads=3, extra_rob=bool(er))
    return R["sigma"], R["slope"], R["n_outer"], int(R["extra_rob"]), R["GQ"], R["pvalues"]


def main():
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    units, have = collect()
    print("units", len(units), have, flush=True)
    with mp.Pool(nproc) as pool:
        res = pool.map(run_one, [(u["y"], u["x"], u["extra_rob"]) for u in units], chunksize=1)
    off = np.concatenate([[0], np.cumsum([len(u["y"]) for u in units])]).astype(np.int64)
    classes = list(CLASSES)
    np.savez_compressed(os.path.join(HERE, "pyref_units.npz"), off=off,
                        y=np.concatenate([u["y"] for u in units]), x=np.concatenate([u["x"] for u in units]),
                        extra_rob_arg=np.array([u["extra_rob"] for u in units], dtype=np.int8),
                        cls=np.array([classes.index(u["cls"]) for u in units], dtype=np.int8),
                        cls_names=np.array(classes),
                        sigma=np.array([r[0] for r in res]), slope=np.array([r[1] for r in res]),
                        n_outer=np.array([r[2] for r in res], dtype=np.int32),
                        extra_rob_out=np.array([r[3] for r in res], dtype=np.int8),
                        gq=np.concatenate([r[4] for r in res]), pvalue=np.concatenate([r[5] for r in res]))
    print("written", len(units), "units")


if __name__ == "__main__":
    main()
