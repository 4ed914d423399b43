"""Regenerates tests/golden/*.npz: seeded inputs of each BASELINE.json shape (small P) and the CPU
oracle's outputs on them.  The reference ships no golden vectors and R is not installed here
(SURVEY.md section 4), so these pin the ORACLE (itself cross-checked against an independent
scipy transcription, tests/pyref.py, and mpmath/QUADPACK known answers) — parity unpinned w.r.t. R.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
from needlestack_b200 import synth  # noqa: E402

CASES = {
    # name: (config, P, seed, p_variant, p_germline, caller kwargs, tn?)
    "c1_default": ("C1", 400, 101, 0.02, 0.005, {}, False),
    "c2_default": ("C2", 120, 102, 0.03, 0.01, {}, False),
    "c1_extra_rob": ("C1", 400, 103, 0.02, 0.03, {"extra_rob": 1}, False),
    "c1_all_snvs_afmin": ("C1", 150, 104, 0.03, 0.01, {"output_all_SNVs": 1, "afmin_power": 0.05}, False),
    "c4_tn_pairs": ("C4", 20, 105, 0.1, 0.05, {}, True),
    "c3_deep": ("C3", 4, 106, 0.25, 0.0, {}, False),
}
KEYS = ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "sigma", "slope", "gated",
        "extra_rob", "ref_inv", "gate1", "gate2", "max_gq", "n_outer", "maxit_hit", "quad_error", "pooled")


def make(name):
    cfgname, P, seed, pv, pg, kw, tn = CASES[name]
    cfg = synth.CONFIGS[cfgname]
    ref, counts = synth.generate(cfg, P=P, seed=seed, p_variant=pv, p_germline=pg)
    if name == "c1_default":
        ref[::37] = 4  # a few 'N' reference bases: needlestack.r:319 skips them
    pairs = synth.tn_pairs(cfg) if tn else None
    O = orc.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), orc.call_params(pairs=pairs, **kw),
                           n_threads=8)
    out = {k: O[k] for k in KEYS}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), ref=ref, counts=counts, **out)
    print(name, counts.shape, "fitted", int((O["gated"] == 0).sum()), "gate2", int(O["gate2"].sum()))


if __name__ == "__main__":
    for nm in CASES:
        make(nm)
