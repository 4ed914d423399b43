"""The C oracle against an independent scipy/numpy transcription of glmrob.nb() (tests/pyref.py):
two restatements with disjoint numerics (Loader+Lentz+own QUADPACK port vs boost+scipy's QUADPACK)
must agree far below the parity tolerance — the closest thing to a pin without R.

tests/golden/pyref_units.npz holds 1121 units (inputs + pyref's outputs; tests/golden/make_pyref_fixture.py, ~10
CPU-minutes) of five classes: plain converged fits of the C1 / C2 / C5 shapes, units that stop on maxit = 30, units on
the spline + integrate() branch, extra-robust units, reference-inverted units.  The oracle is run on ALL of them
here; pyref itself is re-run on a few cheap ones to show that the fixture is what the committed script produces."""
import os

import numpy as np
import pytest

import golden_util
import orc
import pyref

TOL_SIGMA_REL, TOL_SIGMA_ABS, TOL_SLOPE, TOL_PHRED = 1e-8, 2e-8, 1e-8, 1e-7


@pytest.fixture(scope="module")
def units():
    F = np.load(os.path.join(golden_util.GOLDEN_DIR, "pyref_units.npz"))
    return {k: F[k] for k in F.files}


def _unit(F, i):
    a, b = F["off"][i], F["off"][i + 1]
    return F["y"][a:b].astype(float), F["x"][a:b].astype(float), int(F["extra_rob_arg"][i]), slice(a, b)


def test_fixture_covers_every_class(units):
    F = units
    names = list(F["cls_names"])
    cnt = {nm: int((F["cls"] == k).sum()) for k, nm in enumerate(names)}
    assert len(F["off"]) - 1 >= 1000
    assert cnt["spline"] >= 50 and cnt["maxit"] >= 50 and cnt["extra_rob"] >= 50 and cnt["ref_inv"] >= 50
    assert int((F["n_outer"] == 30).sum()) >= 50  # the non-converged class really stops on maxit


def test_oracle_agrees_with_independent_transcription_on_all_units(units):
    F = units
    U = len(F["off"]) - 1
    worst = dict(sigma=0.0, slope=0.0, phred=0.0)
    for i in range(U):
        y, x, er, sl = _unit(F, i)
        O = orc.glmrob_nb(y, x, min_coverage=30, min_reads=3, extra_rob=er)
        assert not O["info"].gated, i
        ds = abs(O["sigma"] - F["sigma"][i])
        assert ds <= max(TOL_SIGMA_REL * F["sigma"][i], TOL_SIGMA_ABS), (i, O["sigma"], F["sigma"][i])
        dl = abs(O["slope"] - F["slope"][i]) / F["slope"][i]
        assert dl <= TOL_SLOPE, (i, O["slope"], F["slope"][i])
        assert O["info"].n_outer == F["n_outer"][i], i                       # same trajectory, not just the same end
        assert bool(O["extra_rob"]) == bool(F["extra_rob_out"][i]), i
        assert bool(O["info"].maxit_hit) == bool(F["n_outer"][i] >= 30), i
        g = F["gq"][sl]
        dg = np.max(np.abs(O["GQ"] - g) / np.maximum(1, g))
        # p-values in the Phred domain (uncapped, down to 1e-300): scipy's far tails are only good to ~1e-4 relative
        p = F["pvalue"][sl]
        m = (p > 0) & (O["pvalues"] > 0)
        dp = np.max(np.abs(-10 * np.log10(O["pvalues"][m]) + 10 * np.log10(p[m])) / np.maximum(1, -10 * np.log10(p[m])))
        assert dg <= TOL_PHRED and dp <= TOL_PHRED, (i, dg, dp)
        worst["sigma"] = max(worst["sigma"], ds / F["sigma"][i])
        worst["slope"] = max(worst["slope"], dl)
        worst["phred"] = max(worst["phred"], dg, dp)
    print("worst relative differences over", U, "units:", worst)


def test_fixture_is_reproducible(units):
    """pyref re-run live on the cheapest units of every class that has cheap ones"""
    F = units
    n_s = np.diff(F["off"])
    done = 0
    for k in range(len(F["cls_names"])):
        idx = np.where((F["cls"] == k) & (n_s <= 24) & (F["n_outer"] <= 3))[0][:4]
        for i in idx:
            y, x, er, sl = _unit(F, i)
            R = pyref.glmrob_nb(y, x, min_coverage=30, min_reads=3, extra_rob=bool(er))
            assert R["sigma"] == F["sigma"][i] and R["slope"] == F["slope"][i] and R["n_outer"] == F["n_outer"][i]
            assert np.array_equal(R["GQ"], F["gq"][sl])
            done += 1
    assert done >= 8


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
