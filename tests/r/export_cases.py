"""Writes the golden fixtures' fitted units as plain text for tests/r/crosscheck.R (run where R is installed):
one line per unit: case, unit, n, sigma, slope, then y[1..n], x[1..n], GQ[1..n] as the oracle / CUDA path report them.
Usage: python tests/r/export_cases.py > cases.tsv"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import golden_util  # noqa: E402

BASE_ROWS = {0: (1, 5), 1: (2, 6), 2: (3, 7), 3: (4, 8)}  # A, T, C, G: forward / reverse planes

for name in ("c1_default", "c2_default", "c3_deep"):
    g = golden_util.load(name)
    kw, _ = golden_util.case_params(name)
    if kw:
        continue  # glmrob.nb is called with needlestack.r's defaults below
    ref, counts = g["ref"], g["counts"]
    n = counts.shape[2]
    done = 0
    for u in np.nonzero(g["gated"] == 0)[0]:
        p, k = divmod(int(u), 3)
        alts = [b for b in range(4) if b != ref[p]]
        b = int(ref[p]) if g["ref_inv"][u] else alts[k]
        y = counts[p, BASE_ROWS[b][0]] + counts[p, BASE_ROWS[b][1]]
        x = counts[p, 0]
        f = [name, str(u), str(n), repr(float(g["sigma"][u])), repr(float(g["slope"][u]))]
        f += [str(int(v)) for v in y] + [str(int(v)) for v in x] + [repr(float(v)) for v in g["gq"][u]]
        print("\t".join(f))
        done += 1
        if done >= 200:
            break
