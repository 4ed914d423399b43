"""The streaming gate kernel's code paths (128-bit / 64-bit loads, scalar fallback, integer coverage tests, 32- and 64-bit
reference-inversion test, min_af, the even-n median tie) against the oracle: gate / inversion flags bit-exact, fits
within the usual tolerances."""
import dataclasses

import numpy as np
import pytest

import parity
from needlestack_b200 import NS_EMIT_ALL, NS_F_GATED, NS_F_REF_INV, make_params, synth

pytestmark = pytest.mark.gpu


def _gpu_dict(R):
    out = {k: R.dense(k) for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs")}
    out["gt"] = R.dense("gt", fill=0)
    out["status"] = R.dense("status", fill=0)
    out["pooled"] = R.dense("pooled")
    for k in ("sigma", "slope", "max_gq", "flags", "n_outer", "n_irls", "n_obj"):
        out[k] = R[k]
    return out


def _check(caller, oracle, ref, counts, **kw):
    O = oracle.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), oracle.call_params(**kw), n_threads=8)
    R = caller.call_snv(ref, counts, prm=make_params(emit_mode=NS_EMIT_ALL, **kw))
    rep = parity.compare(O, _gpu_dict(R), label="gate")
    assert rep["gate_mismatch"] == 0
    assert np.array_equal((R.flags & NS_F_GATED) != 0, O["gated"] != 0)
    assert np.array_equal((R.flags & NS_F_REF_INV) != 0, O["ref_inv"] != 0)
    return rep, R, O


@pytest.mark.parametrize("n", [33, 34, 36, 50, 64, 66, 100, 128, 130])
def test_sample_counts_cover_every_load_path(caller, oracle, n):
    """n = 36/64/100/128: 128-bit loads; 34/50: 64-bit; 33, 66 (not a multiple of 4 above 64), 130: scalar paths"""
    cfg = dataclasses.replace(synth.CONFIGS["C1"], n=n)
    ref, counts = synth.generate(cfg, P=300, seed=100 + n, p_variant=0.03, p_germline=0.02)
    for p in range(0, 300, 25):  # the ALT allele is the major one: swap the REF planes with an ALT's
        a = (int(ref[p]) + 1) % 4
        for o in (1, 5):
            counts[p, [o + int(ref[p]), o + a]] = counts[p, [o + a, o + int(ref[p])]]
    rep, R, O = _check(caller, oracle, ref, counts)
    assert rep["n_fitted"] > 20 and (R.flags & NS_F_REF_INV).sum() >= 12


def test_fractional_and_extreme_min_coverage(caller, oracle):
    cfg = dataclasses.replace(synth.CONFIGS["C1"], n=40, depth=40)
    ref, counts = synth.generate(cfg, P=400, seed=7, p_variant=0.03, p_germline=0.01)
    for mc in (29.5, 30.0, 40.0, 41.25, 0.0, 1e6):
        rep, R, O = _check(caller, oracle, ref, counts, min_coverage=mc)
    assert rep["n_fitted"] == 0  # min_coverage above every depth gates everything


def test_median_tie_on_even_n(caller, oracle):
    """exactly n/2 depths below min_coverage: the median is the mean of max{x < c} and min{x >= c}"""
    cfg = dataclasses.replace(synth.CONFIGS["C1"], n=36, depth=60)
    ref, counts = synth.generate(cfg, P=200, seed=9, p_variant=0.05, p_germline=0.0)
    rng = np.random.default_rng(3)
    for p in range(200):
        lo = rng.choice(36, 18, replace=False)
        d = counts[p, 0].copy()
        d[:] = np.maximum(d, 31)
        d[lo] = rng.integers(24, 30, 18)
        d[lo[0]] = 29 if p % 2 else 27          # max{x < 30}
        hi = np.setdiff1d(np.arange(36), lo)
        d[hi[0]] = 31 if p % 3 else 33          # min{x >= 30}: (29+31)/2 = 30 is NOT < 30, (27+31)/2 = 29 is
        counts[p, 0] = np.maximum(d, counts[p, 1:].reshape(2, 4, 36).sum(axis=(0, 1)))
    rep, R, O = _check(caller, oracle, ref, counts, min_coverage=30)
    g = (R.flags & NS_F_GATED) != 0
    assert g.any() and (~g).any()


def test_min_af_gate(caller, oracle):
    cfg = dataclasses.replace(synth.CONFIGS["C1"], n=48)
    ref, counts = synth.generate(cfg, P=400, seed=11, p_variant=0.05, p_germline=0.01)
    rep0, R0, _ = _check(caller, oracle, ref, counts, min_af=0.0)
    rep1, R1, _ = _check(caller, oracle, ref, counts, min_af=0.05)
    assert 0 < rep1["n_fitted"] < rep0["n_fitted"]


def test_huge_depths_take_the_64_bit_inversion_test(caller, oracle):
    """depths of 2^28 and more: 4x - 5y no longer fits 32 bits (and counts above the depth, as in malformed input)"""
    n, P = 36, 12
    rng = np.random.default_rng(5)
    counts = np.zeros((P, 9, n), dtype=np.int32)
    ref = np.zeros(P, dtype=np.uint8)  # REF = A
    for p in range(P):
        dp = rng.integers(3 * 10 ** 8, 5 * 10 ** 8, n)
        frac = 0.9 if p % 2 else 0.7      # T allele at 90 % (inverted) or 70 % (not) of the depth in most samples
        t = (dp * frac).astype(np.int64)
        t[: n // 6] = 0
        a = dp - t
        counts[p, 0] = dp
        counts[p, 1] = a // 2
        counts[p, 5] = a - a // 2
        counts[p, 2] = t // 2
        counts[p, 6] = t - t // 2
    counts[3, 2, :] = 2 ** 29   # T counts far above the depth of a 'small' position: the 32-bit form would wrap
    counts[3, 0, :] = 1000
    counts[3, 1, :] = counts[3, 5, :] = 0
    counts[3, 6, :] = 0
    # min_reads above every count: the inversion flags are what is under test, no unit is fitted
    O = oracle.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), oracle.call_params(min_reads=3e9), n_threads=4)
    R = caller.call_snv(ref, counts, prm=make_params(emit_mode=NS_EMIT_ALL, min_reads=3e9))
    assert np.array_equal((R.flags & NS_F_REF_INV) != 0, O["ref_inv"] != 0)
    assert np.array_equal((R.flags & NS_F_GATED) != 0, O["gated"] != 0) and R.n_fitted == 0
    inv = ((R.flags & NS_F_REF_INV) != 0).reshape(P, 3)
    assert inv[1, 0] and not inv[0, 0] and inv[3, 0]  # unit 0 of a REF = A position is the T allele


def test_unaligned_device_planes_take_the_scalar_path(caller):
    """ns_call_snv_device with count planes that are only 4-byte aligned: the vector-load forms do not apply; the
    result is that of the aligned call, bit for bit"""
    import torch
    cfg = dataclasses.replace(synth.CONFIGS["C1"], n=100)
    ref, counts = synth.generate(cfg, P=400, seed=31, p_variant=0.03, p_germline=0.01)
    prm = make_params(emit_mode=NS_EMIT_ALL)
    A = caller.call_snv(ref, counts, prm=prm)
    buf = torch.zeros(counts.size + 4, dtype=torch.int32, device="cuda")
    d_ref = torch.from_numpy(ref).cuda()
    for off in (1, 2):  # 4- and 8-byte aligned
        view = buf[off:off + counts.size]
        view.copy_(torch.from_numpy(counts.reshape(-1)))
        torch.cuda.synchronize()
        assert view.data_ptr() % 16 != 0
        B = caller.call_snv_device(d_ref.data_ptr(), view.data_ptr(), 400, 100, prm)
        assert np.array_equal(A.flags, B.flags)
        assert np.array_equal(A.sigma, B.sigma, equal_nan=True) and np.array_equal(A.slope, B.slope, equal_nan=True)
        assert np.array_equal(A.dense("gq"), B.dense("gq"), equal_nan=True)
