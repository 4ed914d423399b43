"""GPU parity for SURVEY.md 8(f) rows 3 and 4: the annotate_vcf front end (every unit against the oracle's
per-allele body) and the plot_rob_nb q-value grid (kernel k_qgrid against the oracle's toQvalue)."""
import numpy as np
import pytest

from needlestack_b200 import annotate, contours

pytestmark = pytest.mark.gpu


def _unit(seed, n, depth, e=2e-3, sigma=0.3):
    rng = np.random.default_rng(seed)
    cov = rng.negative_binomial(20, 20 / (20 + depth), n).astype(np.int32) + 30
    ma = rng.negative_binomial(1 / sigma, 1 / (1 + sigma * e * cov)).astype(np.int32)
    ma[n // 3] += max(8, depth // 10)
    return cov, ma


@pytest.mark.parametrize("n,depth,seed", [(25, 100, 1), (100, 300, 2), (300, 5000, 3), (1, 80, 4)])
def test_qvalue_grid_matches_oracle(caller, oracle, n, depth, seed):
    cov, ma = _unit(seed, n, depth)
    reg = caller.glmrob_nb(ma, cov, min_coverage=1, min_reads=1) if n >= 10 else None
    sigma = reg["coef"]["sigma"] if reg else 0.2
    slope = reg["coef"]["slope"] if reg else 1e-3
    rng = np.random.default_rng(seed + 10)
    m = 4000
    gx = rng.integers(0, int(cov.max()) + 1, m)
    gy = np.minimum(rng.integers(0, 60, m) * rng.integers(0, 2, m) + rng.integers(0, 8, m), gx + rng.integers(0, 3, m))
    gx[:3], gy[:3] = [0, 0, int(cov.max())], [0, 1, int(cov.max())]
    want = oracle.plot_toqvalue(cov, ma, sigma, slope, gx, gy)
    got = caller.qvalue_grid(sigma, slope, cov, ma, gx, gy)
    assert np.all(got[gy > gx] == 0)
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got))
    assert np.all(np.abs(got[fin] - want[fin]) <= 1e-6 * np.maximum(1, np.abs(want[fin])))  # BASELINE tolerance
    assert np.array_equal(got[~fin], want[~fin])


def test_contour_grid_end_to_end(caller, oracle):
    cov, ma = _unit(7, 60, 200)
    reg = caller.glmrob_nb(ma, cov, min_coverage=1, min_reads=1)
    G = contours.contour_grid(caller, reg)
    gx, gy = np.meshgrid(G["xgrid"], G["ygrid"], indexing="ij")
    want = oracle.plot_toqvalue(cov, ma, reg["coef"]["sigma"], reg["coef"]["slope"], gx, gy).reshape(gx.shape)
    assert np.all(np.abs(G["matgrid"] - want) <= 1e-6 * np.maximum(1, np.abs(want)))
    assert set(G["lines"]) == {10, 30, 50, 70, 100} and len(G["lines"][10][0]) > 0
    with pytest.raises(Exception):
        caller.qvalue_grid(float("nan"), 1e-3, cov, ma, [1], [1])


def _synthetic_vcf(path, n=40, L=30, seed=5):
    """DP / AD lines: error-level alleles, a few carriers, multi-allelic lines, one line where the ALT is the
    major allele in most samples (reference switching), missing cells"""
    rng = np.random.default_rng(seed)
    names = ["S%d" % i for i in range(n)]
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n##INFO=<ID=X,Number=1,Type=Integer,Description=\"x\">\n")
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        for ln in range(L):
            nalt = 1 + (ln % 3 == 0)
            dp = rng.negative_binomial(20, 20 / (20 + 150.0), n) + 60
            ad = np.zeros((n, 1 + nalt), dtype=np.int64)
            for a in range(nalt):
                ad[:, 1 + a] = rng.binomial(dp, 10 ** rng.uniform(-3.2, -2.2))
            for s in rng.choice(n, 2, replace=False):
                ad[s, 1] += rng.binomial(dp[s], rng.uniform(0.05, 0.5))
            if ln == 4:  # ALT is the major allele
                ad[:, 1] = dp - rng.binomial(dp, 0.01)
                ad[:3, 1] = rng.binomial(dp[:3], 0.4)
            ad[:, 0] = np.maximum(dp - ad[:, 1:].sum(axis=1), 0)
            cells = []
            for s in range(n):
                if ln == 7 and s == 5:
                    cells.append("./.:.:.")
                else:
                    cells.append("0/0:%s:%d" % (",".join(map(str, ad[s])), dp[s]))
            f.write("1\t%d\t.\tA\t%s\t.\t.\tX=1\tGT:AD:DP\t%s\n" % (1000 + ln, ",".join("CGT"[:nalt]), "\t".join(cells)))
    return names


def test_annotate_vcf_against_oracle(caller, oracle, tmp_path):
    src, dst = str(tmp_path / "in.vcf"), str(tmp_path / "out.vcf")
    names = _synthetic_vcf(src)
    n = len(names)
    per_rec = annotate.annotate_vcf(caller, src, dst, dict(min_coverage=50, min_reads=5))
    header, names2, recs = annotate.read_vcf(src)
    D, Y, RP, owner = annotate.build_units(recs, n)
    prm = oracle.call_params(min_coverage=50, min_reads=5)
    z = np.zeros(n, dtype=np.int32)
    n_fit = n_inv = 0
    for u, (r, a) in enumerate(owner):
        O = oracle.call_unit(D[u], Y[u], z, RP[u], z, prm)
        g = per_rec[r][a]
        assert g["inv_ref"] == bool(O["info"].ref_inv)
        n_inv += g["inv_ref"]
        if O["info"].glm.gated:
            assert g["sigma"] != g["sigma"] and np.all(g["gq"] == 0)
            continue
        n_fit += 1
        assert g["sigma"] == pytest.approx(O["info"].glm.sigma, rel=1e-6, abs=2e-8)
        assert g["slope"] == pytest.approx(O["info"].glm.slope, rel=1e-6)
        assert np.all(np.abs(g["gq"] - O["gq"]) <= 1e-6 * np.maximum(1, np.abs(O["gq"])))
    assert n_fit >= 20 and n_inv >= 1
    # the written file: header additions, INFO / FORMAT fields, "." for the allele that is not switched
    out_header, _, out = annotate.read_vcf(dst)
    assert sum(h.startswith("##INFO=<ID=ERR") or h.startswith("##FORMAT=<ID=QVAL") for h in out_header) == 3
    for r, f in enumerate(out):
        assert f[8].endswith(":QVAL:QVAL_INV") and "ERR=" in f[7] and "SIG=" in f[7] and f[7].startswith("X=1;")
        nalt = len(per_rec[r])
        q, qi = f[9].split(":")[-2:]
        assert len(q.split(",")) == nalt and len(qi.split(",")) == nalt
        inv = [g["inv_ref"] for g in per_rec[r]]
        if any(inv):
            assert "WARN=" in f[7] and "INV_REF" in f[7]
            for k, iv in enumerate(inv):
                assert (q.split(",")[k] == ".") == iv and (qi.split(",")[k] == ".") == (not iv)
        else:
            assert q == qi  # annotate_vcf.r:129-135 leaves both fields filled on ordinary lines
            assert float(q.split(",")[0]) == pytest.approx(per_rec[r][0]["gq"][0], rel=1e-12, abs=1e-12)
