"""Independent numpy/scipy transcription of the per-allele body of needlestack.r AFTER the regression (test
infrastructure, like tests/pyref.py; nothing here is shared with oracle/ or with the kernels):

  toQvalueN / toQvalueT            needlestack.r:242-252   qnbinom / qbinom quantile, NB upper tail, Holm (p.adjust default)
  tumour / normal status           needlestack.r:342-373   the strict-inequality table, contamination relabel
  SOR, RVSB, Fisher strand bias    needlestack.r:206-239   per sample and pooled, NA -> -1, Inf -> 99, Phred cap 1000
  output gates and genotypes       needlestack.r:381-409

Inputs are the count rows of one site-allele and the (sigma, slope, GQ) of its fit; R's vector semantics (NA
propagation in comparisons, which(), na.rm) are written out where they matter."""
import numpy as np
from scipy import stats

STATUS = (".", "SOMATIC", "GERMLINE_CONFIRMED", "GERMLINE_UNCONFIRMED", "GERMLINE_UNCONFIRMABLE", "UNKNOWN",
          "POSSIBLE_CONTAMINATION")


def nb_tail_incl(y, size, mu):
    """dnbinom(y, size, mu) + pnbinom(y, size, mu, lower.tail = FALSE) = P(Y >= y)"""
    p = size / (size + mu)
    return stats.nbinom.pmf(y, size, p) + stats.nbinom.sf(y, size, p)


def holm_last(p):
    """p.adjust(p) (method 'holm') evaluated at the LAST element of p"""
    p = np.asarray(p, dtype=float)
    m = len(p)
    o = np.argsort(p, kind="stable")
    adj = np.minimum(1.0, np.maximum.accumulate((m - np.arange(m)) * p[o]))
    out = np.empty(m)
    out[o] = adj
    return out[-1]


def to_qvalue(x, ma_count, coverage, sigma, slope, y_star):
    with np.errstate(divide="ignore"):
        p = nb_tail_incl(np.append(ma_count, y_star), 1.0 / sigma, slope * np.append(coverage, x))
        return -10.0 * np.log10(holm_last(p))


def to_qvalue_n(x, ma_count, coverage, sigma, slope, sigma_normal):
    size = 1.0 / sigma_normal
    y_star = stats.nbinom.ppf(0.01, size, size / (size + 0.5 * x))
    return to_qvalue(x, ma_count, coverage, sigma, slope, y_star)


def to_qvalue_t(x, ma_count, coverage, sigma, slope, afmin_power):
    y_star = stats.binom.ppf(0.01, x, afmin_power)
    return to_qvalue(x, ma_count, coverage, sigma, slope, y_star)


def sor(r_f, a_f, r_r, a_r):
    with np.errstate(divide="ignore", invalid="ignore"):
        ref_ratio = np.maximum(r_f, r_r) / np.minimum(r_f, r_r)
        alt_ratio = np.maximum(a_f, a_r) / np.minimum(a_f, a_r)
        R = (r_f / r_r) * (a_r / a_f)
        return np.log((R + 1.0 / R) / (ref_ratio / alt_ratio))


def fisher_phred(vp, vm, rp, rm):
    p = stats.fisher_exact([[int(vp), int(rp)], [int(vm), int(rm)]], alternative="two-sided")[1]
    with np.errstate(divide="ignore"):
        return min(-10.0 * np.log10(p), 1000.0)


def strand_annotations(vp, vm, rp, rm):
    vp, vm, rp, rm = (np.asarray(a, dtype=float) for a in (vp, vm, rp, rm))
    cp, cm = vp + rp, vm + rm
    with np.errstate(divide="ignore", invalid="ignore"):
        rvsb = np.maximum(vp * cm, vm * cp) / (vp * cm + vm * cp)
    rvsb[np.isnan(rvsb)] = -1.0
    s = sor(rp, vp, rm, vm)
    s[np.isnan(s)] = -1.0
    s[np.isinf(s)] = 99.0
    fs = np.array([fisher_phred(vp[i], vm[i], rp[i], rm[i]) for i in range(len(vp))])
    with np.errstate(divide="ignore", invalid="ignore"):
        a, b = vp.sum() * cm.sum(), vm.sum() * cp.sum()
        all_rvsb = np.float64(max(a, b)) / np.float64(a + b)
    if np.isnan(all_rvsb):
        all_rvsb = -1.0
    all_sor = float(sor(np.float64(rp.sum()), np.float64(vp.sum()), np.float64(rm.sum()), np.float64(vm.sum())))
    if np.isnan(all_sor):
        all_sor = -1.0
    if np.isinf(all_sor):
        all_sor = 99.0
    return dict(sor=s, rvsb=rvsb, fs=fs, all_sor=all_sor, all_rvsb=float(all_rvsb),
                all_fs=fisher_phred(vp.sum(), vm.sum(), rp.sum(), rm.sum()))


def allele_body(dp, ma_used, gq, sigma, slope, *, gq_threshold=50.0, sigma_normal=0.1, afmin_power=-1.0, pairs=None):
    """qval_minAF and somatic status of one fitted site-allele.  ma_used = the counts the regression saw (the
    reference allele's when the site is reference-inverted).  pairs = dict(tindex, nindex, only_t, only_n), 0-based."""
    dp = np.asarray(dp, dtype=float)
    n = len(dp)
    qn = lambda i: to_qvalue_n(dp[i], ma_used, dp, sigma, slope, sigma_normal)  # noqa: E731
    qt = lambda i: to_qvalue_t(dp[i], ma_used, dp, sigma, slope, afmin_power)  # noqa: E731
    qv = np.zeros(n)
    status = np.zeros(n, dtype=int)
    if pairs is None:
        for i in range(n):
            qv[i] = qn(i) if afmin_power == -1 else qt(i)
        return qv, status
    T, N = np.asarray(pairs["tindex"], dtype=int), np.asarray(pairs["nindex"], dtype=int)
    oT, oN = np.asarray(pairs.get("only_t", []), dtype=int), np.asarray(pairs.get("only_n", []), dtype=int)
    for i in N:
        qv[i] = qn(i)
    germline_somewhere = np.sum(gq[N] > gq_threshold) > 0
    for i in T:
        qv[i] = qn(i) if germline_somewhere else qt(i)
    for i in oN:
        qv[i] = qn(i)
    for i in oT:
        qv[i] = qt(i)
    code = {s: k for k, s in enumerate(STATUS)}
    thr = gq_threshold
    status[oT[gq[oT] > thr]] = code["UNKNOWN"]
    status[oN[gq[oN] > thr]] = code["GERMLINE_UNCONFIRMABLE"]
    gT, gN, qT, qN = gq[T], gq[N], qv[T], qv[N]

    def both(mask, name):
        status[T[mask]] = code[name]
        status[N[mask]] = code[name]

    both((gT < thr) & (qT > thr) & (gN > thr), "GERMLINE_UNCONFIRMED")
    both((gT < thr) & (qT < thr) & (gN > thr), "GERMLINE_UNCONFIRMABLE")
    both((gT > thr) & (qN < thr) & (gN < thr), "UNKNOWN")
    both((gT > thr) & (gN > thr), "GERMLINE_CONFIRMED")
    status[T[(gT > thr) & (qN > thr) & (gN < thr)]] = code["SOMATIC"]
    germ = np.isin(status, [code["GERMLINE_CONFIRMED"], code["GERMLINE_UNCONFIRMED"], code["GERMLINE_UNCONFIRMABLE"],
                            code["UNKNOWN"]])
    if germ.any():
        som = T[status[T] == code["SOMATIC"]]
        status[som] = code["POSSIBLE_CONTAMINATION"]
    return qv, status


def genotypes(gq, sbs, ma_used, dp, qv, ref_inv, *, gq_threshold=50.0, sb_threshold=100.0):
    """needlestack.r:403-409; codes 0 = 0/0, 1 = 0/1, 2 = 1/1, 3 = ./."""
    n = len(gq)
    gt = np.full(n, 2 if ref_inv else 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        af = np.asarray(ma_used, dtype=float) / np.asarray(dp, dtype=float)
    called = (gq >= gq_threshold) & (sbs <= sb_threshold)
    gt[called & (af < 0.75)] = 1
    gt[called & (af >= 0.75)] = 0 if ref_inv else 2
    gt[(gq < gq_threshold) & (qv < gq_threshold)] = 3
    return gt
