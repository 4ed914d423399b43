"""The C oracle against an independent scipy/numpy transcription of glmrob.nb() (tests/pyref.py):
two restatements with disjoint numerics (Loader+Lentz+own QUADPACK port vs boost+scipy's QUADPACK)
must agree far below the parity tolerance — the closest thing to a pin without R."""
import numpy as np
import pytest

import golden_util
import orc
import parity
import pyref


@pytest.mark.parametrize("name,units", [("c1_default", 8), ("c2_default", 5), ("c3_deep", 1)])
def test_independent_restatements_agree(name, units):
    G = golden_util.load(name)
    done = 0
    for u in np.where((G["gated"] == 0) & (G["maxit_hit"] == 0) & (G["ref_inv"] == 0))[0]:
        dp, vp, vm, rp, rm = parity.snv_unit_rows(G["ref"], G["counts"], int(u))
        y, x = (vp + vm).astype(float), dp.astype(float)
        R = pyref.glmrob_nb(y, x, min_coverage=30, min_reads=3, extra_rob=False)
        assert abs(R["sigma"] - G["sigma"][u]) <= max(1e-8 * G["sigma"][u], 2e-8)
        assert abs(R["slope"] - G["slope"][u]) <= 1e-8 * G["slope"][u]
        assert np.max(np.abs(R["GQ"] - G["gq"][u]) / np.maximum(1, G["gq"][u])) <= 1e-7
        assert R["n_outer"] == G["n_outer"][u]
        done += 1
        if done >= units:
            break
    assert done == units


def test_gate_and_extra_robust_agree():
    rng = np.random.default_rng(9)
    for _ in range(40):
        n = int(rng.integers(11, 40))
        x = rng.integers(0, 120, size=n).astype(float)
        y = np.minimum(rng.poisson(0.01 * x) + (rng.random(n) < 0.2) * rng.integers(0, 60, size=n), x)
        O = orc.glmrob_nb(y, x, min_coverage=30, min_reads=3, extra_rob=1)
        R = pyref.glmrob_nb(y, x, min_coverage=30, min_reads=3, extra_rob=True)
        assert bool(O["info"].gated) == bool(np.isnan(R["sigma"]))
        assert bool(O["extra_rob"]) == bool(R["extra_rob"])
