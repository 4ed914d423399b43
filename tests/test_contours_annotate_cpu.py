"""CPU suite for the two "next" rows of SURVEY.md 8(f): the plot_rob_nb contour grid (host logic + the oracle's
toQvalue against scipy) and the annotate_vcf front end (VCF parsing, unit construction, WARN logic)."""
import os
import sys

import numpy as np
import pytest
from scipy import stats

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import orc  # noqa: E402
from needlestack_b200 import annotate, contours  # noqa: E402


def _unit(seed=1, n=40, depth=300, e=2e-3, sigma=0.3):
    rng = np.random.default_rng(seed)
    cov = rng.negative_binomial(20, 20 / (20 + depth), n).astype(np.int32) + 30
    ma = rng.negative_binomial(1 / sigma, 1 / (1 + sigma * e * cov)).astype(np.int32)
    ma[3] += 25
    return cov, ma


def _holm_last(p):
    """p.adjust(p, 'holm')[last] as R computes it: pmin(1, cummax((m - i + 1) * p[o]))[ro]"""
    m = len(p)
    o = np.argsort(p, kind="stable")
    adj = np.minimum(1, np.maximum.accumulate((m - np.arange(m)) * p[o]))
    out = np.empty(m)
    out[o] = adj
    return out[-1]


def test_oracle_toqvalue_matches_scipy():
    """plot_rob_nb.r:83-87 restated with scipy's negative binomial (independent of the oracle's nmath port)"""
    cov, ma = _unit()
    sigma, slope = 0.25, 2.2e-3
    gx = np.array([0, 50, 300, 300, 300, 700, 700, 10, 400])
    gy = np.array([0, 0, 1, 6, 40, 3, 30, 11, 400])
    got = orc.plot_toqvalue(cov, ma, sigma, slope, gx, gy)
    size = 1 / sigma

    def tail(y, x):
        mu = slope * x
        if mu == 0:
            return 1.0 if y == 0 else 0.0
        return stats.nbinom.sf(y - 1, size, size / (size + mu))

    p_obs = np.array([tail(y, x) for y, x in zip(ma, cov)])
    for g in range(len(gx)):
        if gy[g] > gx[g]:
            assert got[g] == 0
            continue
        want = -10 * np.log10(_holm_last(np.append(p_obs, tail(gy[g], gx[g]))))
        assert got[g] == pytest.approx(want, rel=1e-9, abs=1e-9), (gx[g], gy[g])


class _OracleCaller:
    """stands in for Caller.qvalue_grid so that the host-side grid logic is testable without a GPU"""

    def qvalue_grid(self, sigma, slope, cov, ma, gx, gy):
        return orc.plot_toqvalue(cov, ma, sigma, slope, gx, gy).reshape(np.shape(gx))


def test_grid_vectors_follow_the_reference_rules():
    x, y = contours.grid_vectors(40, 12)  # 480 <= 2500: every integer
    assert np.array_equal(x, np.arange(41)) and np.array_equal(y, np.arange(13))
    x, y = contours.grid_vectors(500, 30)  # too many points, ylim <= 50: all AO, sampled DP
    assert np.array_equal(y, np.arange(31)) and len(x) == round(2500 / 31) and x[0] == 0 and x[-1] == 500
    x, y = contours.grid_vectors(5000, 400)  # 50 x 50
    assert len(x) == 50 and len(y) == 50 and y[-1] == 400 and x[-1] == 5000
    assert np.all(x == np.round(x)) and np.all(y == np.round(y))


def test_contour_grid_host_logic():
    cov, ma = _unit(n=30, depth=120)
    reg = dict(coverage=cov, ma_count=ma, coef=dict(sigma=0.2, slope=2e-3))
    G = contours.contour_grid(_OracleCaller(), reg)
    max_dp = int(cov.max())
    assert G["xgrid"][-1] == max_dp and G["matgrid"].shape == (len(G["xgrid"]), len(G["ygrid"]))
    # the zoom limit is the first AO at the largest depth whose q-value reaches 100
    zq = orc.plot_toqvalue(cov, ma, 0.2, 2e-3, np.full(max_dp, max_dp), np.arange(1, max_dp + 1))
    first = np.nonzero(zq >= 100)[0]
    assert G["ylim_zoom"] == (first[0] + 1 if len(first) else pytest.approx(float("nan"), nan_ok=True))
    # q-values grow with AO at fixed DP, and every contour point reaches its level
    m = G["matgrid"]
    ix = len(G["xgrid"]) - 1
    ok = G["ygrid"] <= G["xgrid"][ix]
    assert np.all(np.diff(m[ix][ok]) >= -1e-9)
    for lev, (dps, aos) in G["lines"].items():
        for dp, ao in zip(dps, aos):
            if dp == 0 and ao == 0:
                continue
            i, j = list(G["xgrid"]).index(dp), list(G["ygrid"]).index(ao)
            assert m[i, j] >= lev and (j == 0 or not np.any(m[i, :j] >= lev))
    assert contours.contour_grid(_OracleCaller(), dict(coverage=cov, ma_count=ma, coef=dict(sigma=float("nan"), slope=float("nan")))) is None


VCF = """##fileformat=VCFv4.2
##FORMAT=<ID=GT,Number=1,Type=String,Description="g">
##FORMAT=<ID=AD,Number=R,Type=Integer,Description="a">
##FORMAT=<ID=DP,Number=1,Type=Integer,Description="d">
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3
1\t100\t.\tA\tT\t.\t.\tX=1\tGT:AD:DP\t0/0:90,1:91\t0/1:50,40:95\t./.:.:.
1\t200\t.\tC\tG,T\t.\t.\t.\tGT:DP:AD\t0/0:80:70,6,4\t0/0:60:60,0,0\t0/0:.:50,0,1
1\t300\t.\tG\tA\t.\t.\t.\tGT\t0/0\t0/0\t0/0
"""


def test_annotate_parsing_and_units(tmp_path):
    p = tmp_path / "in.vcf"
    p.write_text(VCF)
    header, names, recs = annotate.read_vcf(str(p))
    assert names == ["S1", "S2", "S3"] and len(recs) == 3 and header[-1].startswith("#CHROM")
    dp, AD = annotate.record_units(recs[0], 3)
    assert dp.tolist() == [91, 95, 0] and AD.tolist() == [[90, 1], [50, 40], [0, 0]]  # missing -> zeros (:57-60)
    dp, AD = annotate.record_units(recs[1], 3)
    assert dp.tolist() == [80, 60, 0] and AD.tolist() == [[70, 6, 4], [60, 0, 0], [50, 0, 1]]
    dp, AD = annotate.record_units(recs[2], 3)  # no AD at all: two zero entries -> one all-zero alternative
    assert AD.shape == (3, 2) and not AD.any() and not dp.any()
    D, Y, RP, owner = annotate.build_units(recs, 3)
    assert owner == [(0, 0), (1, 0), (1, 1), (2, 0)]
    assert Y.tolist() == [[1, 40, 0], [6, 0, 0], [4, 0, 1], [0, 0, 0]]
    assert RP.tolist() == [[90, 55, 0], [70, 60, 0], [70, 60, 0], [0, 0, 0]]  # DP - sum(alt AD), clamped at 0
    o = annotate.parse_args(["--input_vcf=x.vcf.bgz", "--min_coverage=10", "--extra_rob=TRUE"])
    assert o["out_vcf"] == "x_annotated_needlestack.vcf" and o["min_coverage"] == 10 and o["extra_rob"] is True
    assert o["min_reads"] == 5 and o["GQ_threshold"] == 50
    with pytest.raises(SystemExit):
        annotate.parse_args(["--out_vcf=y"])
    assert annotate.fmt(float("nan")) == "." and annotate.fmt(-0.0) == "0" and annotate.fmt(0.1 + 0.2) == "0.3"
