"""The C oracle's per-allele body (qval_minAF, tumour/normal status, strand-bias annotations, output gates, genotypes)
against an independent numpy/scipy transcription of needlestack.r:206-252, :342-409 (tests/pyref_body.py).
tests/test_pyref_vs_oracle.py pins the regression the same way; together they are the closest thing to a pin of the
oracle without R.  The (sigma, slope, GQ) of each fit are taken from the oracle: this file checks what
needlestack.r does with them."""
import numpy as np

import orc
import pyref_body as B
from needlestack_b200 import synth

PHRED_TOL = 1e-6  # |d| <= 1e-6 * max(1, |value|), the parity tolerance of SURVEY.md 8c


def close(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= PHRED_TOL * np.maximum(1.0, np.abs(b))
    return bool(np.all(ok | same_inf))


def unit_rows(ref, counts, u):
    """the five count rows of SNV unit u = 3 * position + k: alleles in the order setdiff(c(A,T,C,G), ref)"""
    p, k = divmod(u, 3)
    r = int(ref[p])
    alt = [b for b in range(4) if b != r][k]
    c = counts[p]
    return c[0], c[1 + alt], c[5 + alt], c[1 + r], c[5 + r]


def plant_inverted(ref, counts, positions, rng):
    """positions where the reference genome carries the minor allele: the first ALT base holds almost every read of
    almost every sample, two samples are heterozygous for the reference base -> the regression runs on the REF counts
    (needlestack.r:330-337) and those two samples are called (QVAL_INV, default genotype 1/1)"""
    n = counts.shape[2]
    for p in positions:
        r = int(ref[p])
        alt = [b for b in range(4) if b != r][0]
        het = rng.choice(n, size=2, replace=False)
        for i in range(n):
            dp = int(counts[p, 0, i])
            others = sum(int(counts[p, 1 + b, i]) + int(counts[p, 5 + b, i]) for b in range(4) if b not in (r, alt))
            n_ref = int(rng.binomial(max(dp - others, 0), 0.5 if i in het else 0.004))
            n_alt = max(dp - others - n_ref, 0)
            for base, tot in ((r, n_ref), (alt, n_alt)):
                fwd = int(rng.binomial(tot, 0.5))
                counts[p, 1 + base, i] = fwd
                counts[p, 5 + base, i] = tot - fwd


def check_batch(ref, counts, O, pairs, afmin_power, max_fitted, rng):
    fitted = np.flatnonzero(O["gated"] == 0)
    called = np.flatnonzero((O["gated"] == 0) & (O["gate1"] == 1))
    some = fitted if len(fitted) <= max_fitted else rng.choice(fitted, size=max_fitted, replace=False)
    seen = dict(units=0, called=0, inv=0, status=set(), gt=set(), inf_qv=0)
    for u in sorted(set(some.tolist()) | set(called.tolist())):
        dp, vp, vm, rp, rm = (a.astype(float) for a in unit_rows(ref, counts, u))
        inv = bool(O["ref_inv"][u])
        ma_used = (rp + rm) if inv else (vp + vm)
        gq = O["gq"][u]
        qv, st = B.allele_body(dp, ma_used, gq, O["sigma"][u], O["slope"][u], afmin_power=afmin_power, pairs=pairs)
        assert close(qv, O["qval_minaf"][u]), (u, qv, O["qval_minaf"][u])
        seen["units"] += 1
        seen["inv"] += inv
        seen["inf_qv"] += int(np.isinf(qv).any())
        if pairs is not None:
            assert np.array_equal(st, O["status"][u]), (u, st, O["status"][u])
            seen["status"] |= set(st.tolist())
        gate1 = bool(np.sum(gq >= 50.0) > 0)
        assert gate1 == bool(O["gate1"][u])
        if not gate1:
            continue
        A = B.strand_annotations(vp, vm, rp, rm)
        assert close(A["sor"], O["sor"][u]) and close(A["rvsb"], O["rvsb"][u]) and close(A["fs"], O["fs"][u]), u
        pooled = O["pooled"][u]
        assert close([A["all_sor"], A["all_rvsb"], A["all_fs"]], pooled[:3]), (u, A, pooled)
        assert np.array_equal(pooled[3:], [rp.sum(), rm.sum(), vp.sum(), vm.sum(), dp.sum()])
        assert bool(np.sum((A["sor"] <= 100.0) & (gq >= 50.0)) > 0) == bool(O["gate2"][u])
        gt = B.genotypes(gq, A["sor"], ma_used, dp, qv, inv)
        assert np.array_equal(gt, O["gt"][u]), (u, gt, O["gt"][u])
        seen["called"] += 1
        seen["gt"] |= set(gt.tolist())
    return seen


def test_snv_body_without_pairs():
    rng = np.random.default_rng(11)
    ref, counts = synth.generate(synth.CONFIGS["C1"], P=1500, seed=321, p_variant=0.03, p_germline=0.01)
    plant_inverted(ref, counts, range(40, 1500, 150), rng)
    O = orc.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), orc.call_params(), n_threads=4)
    seen = check_batch(ref, counts, O, None, -1.0, 120, rng)
    assert seen["called"] >= 30 and seen["inv"] >= 8 and seen["gt"] >= {0, 1, 2}, seen


def test_snv_body_with_binomial_power():
    """--afmin_power given without a pairs file: toQvalueT for every sample (needlestack.r:377-378)"""
    rng = np.random.default_rng(12)
    ref, counts = synth.generate(synth.CONFIGS["C1"], P=600, seed=322, p_variant=0.03)
    O = orc.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), orc.call_params(afmin_power=0.05),
                           n_threads=4)
    seen = check_batch(ref, counts, O, None, 0.05, 60, rng)
    assert seen["called"] >= 8, seen


def test_snv_body_with_pairs():
    rng = np.random.default_rng(13)
    pairs = dict(tindex=[0, 2, 4, 6, 8, 10, 12, 14], nindex=[1, 3, 5, 7, 9, 11, 13, 15], only_t=[16, 17], only_n=[18, 19])
    ref, counts = synth.generate(synth.CONFIGS["C1"], P=2500, seed=323, p_variant=0.05, p_germline=0.01)
    O = orc.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)),
                           orc.call_params(pairs=pairs, afmin_power=0.01), n_threads=4)
    seen = check_batch(ref, counts, O, pairs, 0.01, 60, rng)
    code = {s: k for k, s in enumerate(B.STATUS)}
    assert seen["called"] >= 60, seen
    # every branch of the status table occurs
    assert seen["status"] == set(range(len(B.STATUS))) and seen["gt"] == {0, 1, 2, 3}, seen


def test_indel_body():
    """insertion / deletion alleles (needlestack.r:460-567, :623-730): the same body with all_DP = sum(DP) + sum(AO), the
    output gate on SB_threshold_indel and the genotype on SB_threshold_SNV (as the reference has it, :540 / :556)"""
    rng = np.random.default_rng(14)
    ref, counts = synth.generate(synth.CONFIGS["C1"], P=60, seed=324)
    sb_snv, sb_indel = 0.9, 0.3
    prm = orc.call_params(SB_threshold_SNV=sb_snv, SB_threshold_indel=sb_indel)
    n = counts.shape[2]
    called = gate2_differs = 0
    for p in range(60):
        dp = counts[p, 0].astype(float)
        r = int(ref[p])
        rp, rm = counts[p, 1 + r].astype(float), counts[p, 5 + r].astype(float)
        vp, vm = np.zeros(n), np.zeros(n)
        for i in rng.choice(n, size=1 + int(rng.integers(0, 3)), replace=False):
            tot = int(rng.binomial(int(dp[i]), rng.uniform(0.05, 0.6)))
            vp[i] = int(rng.binomial(tot, rng.uniform(0.1, 0.9)))
            vm[i] = tot - vp[i]
        O = orc.call_unit(dp, vp, vm, rp, rm, prm, is_indel=True)
        g = O["info"].glm
        if g.gated:
            continue
        gq = O["gq"]
        qv, _ = B.allele_body(dp, vp + vm, gq, g.sigma, g.slope)
        assert close(qv, O["qval_minaf"])
        if not np.sum(gq >= 50.0) > 0:
            continue
        A = B.strand_annotations(vp, vm, rp, rm)
        assert close(A["sor"], O["sor"]) and close(A["rvsb"], O["rvsb"]) and close(A["fs"], O["fs"])
        pooled = np.array(O["info"].pooled[:])
        assert close([A["all_sor"], A["all_rvsb"], A["all_fs"]], pooled[:3])
        assert pooled[7] == dp.sum() + (vp + vm).sum()
        gate2 = bool(np.sum((A["sor"] <= sb_indel) & (gq >= 50.0)) > 0)
        assert gate2 == bool(O["info"].gate2)
        gate2_differs += gate2 != bool(np.sum((A["sor"] <= sb_snv) & (gq >= 50.0)) > 0)
        assert np.array_equal(B.genotypes(gq, A["sor"], vp + vm, dp, qv, False, sb_threshold=sb_snv), O["gt"])
        called += 1
    assert called >= 30 and gate2_differs >= 3, (called, gate2_differs)
