"""The C++ VCF writer of the library (ns_vcf_format_snv / _indel / _number, needlestack_b200/csrc/ns_vcf.cpp) against
the Python formatter driver.vcf_line / driver.rfmt (the executable specification of bin/needlestack.r:396-413,
:549-567, :712-730 and of R's cat() number format): byte-identical records for SNV, deletion and insertion units,
with the T/N status text, on oracle-computed values.  Host code only: no GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

import orc
import parity
from needlestack_b200 import _abi, api, driver, readcounts, synth


def _fake_result(O_rows, units, n):
    """Result-like dict (the layout Caller.call_* returns) holding the oracle's rows of the emitted `units`"""
    E = len(units)
    U = len(O_rows["sigma"])
    R = dict(n_emitted=E, emit_unit=np.array(units, dtype=np.int64), flags=O_rows["flags"].astype(np.uint32),
             iters=np.zeros(U, dtype=np.uint32), emit_index=np.full(U, -1, dtype=np.int32))
    R["emit_index"][units] = np.arange(E)
    for k in ("sigma", "slope", "max_gq"):
        R[k] = np.ascontiguousarray(O_rows[k], dtype=np.float64)
    for k in ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs"):
        R[k] = np.ascontiguousarray(O_rows[k][units], dtype=np.float64).reshape(E, n)
    R["pooled"] = np.ascontiguousarray(O_rows["pooled"][units], dtype=np.float64).reshape(E, 8)
    for k in ("gt", "status"):
        R[k] = np.ascontiguousarray(O_rows[k][units], dtype=np.uint8).reshape(E, n)
    return R


def test_number_format_matches_python_formatter():
    L = api.load_library()
    rng = np.random.default_rng(2)
    vals = list(10.0 ** rng.uniform(-320, 20, size=4000) * rng.choice([-1, 1], size=4000)) + \
        list(rng.uniform(-1e7, 1e7, size=4000)) + list(rng.integers(0, 10 ** 9, size=500).astype(float)) + \
        [0.0, -0.0, 1e-4, 9.9999995e-5, 9999999.4, 9999999.6, 1e7, 123456.75, 0.1 + 0.2, 1000.0, 5e-324,
         float("inf"), float("-inf"), -9.643275e-16, 1 / 3, 2 / 3, 1e15, 123456789012.0]
    buf = C.create_string_buffer(600)
    for x in vals:
        L.ns_vcf_format_number(float(x), 0, buf, 600)
        assert buf.value.decode() == driver.rfmt(float(x)), x
    L.ns_vcf_format_number(float("nan"), 0, buf, 600)
    assert buf.value == b"NA"
    L.ns_vcf_format_number(float("nan"), 1, buf, 600)
    assert buf.value == b"NaN"


@pytest.mark.parametrize("tn", [False, True])
def test_records_byte_identical_to_python_formatter(tmp_path, tn):
    n = 24
    cfg = synth.Config(**{**synth.CONFIGS["C4" if tn else "C1"].__dict__, "n": n, "tn_pairs": 12 if tn else 0})
    P = 260
    ref, counts = synth.generate(cfg, P=P, seed=91, p_variant=0.06, p_germline=0.03, p_refswap=0.02)
    ref[3::41] = 4
    cells = synth.random_indel_cells(P, n, seed=9, p_site=0.04)
    table = tmp_path / "t.tsv"
    synth.write_table(str(table), ref, counts, chrom="chrX", start=700, indel_cells=cells)
    names = [f"smp{i}" for i in range(n)]
    rng = np.random.default_rng(4)
    seq = "".join(rng.choice(list("ACGTacgtN"), size=700 + P + 12))
    fa = tmp_path / "r.fa"
    fa.write_text(">chrX some description\n" + "\n".join(seq[i:i + 61] for i in range(0, len(seq), 61)) + "\n>other\nACGT\n")
    pairs = synth.tn_pairs(cfg) if tn else None
    if tn:  # a contamination pattern: germline in pair 1, tumour-only signal in pair 4 (see tests/test_gpu_driver_pairs.py)
        for p in range(10, P, 37):
            if ref[p] > 3:
                continue
            r = int(ref[p])
            alt = [b for b in range(4) if b != r][0]
            for s_, af in ((1, 0.5), (13, 0.5), (4, 0.3)):
                y = int(af * counts[p, 0, s_])
                counts[p, 1 + alt, s_], counts[p, 5 + alt, s_] = y // 2, y - y // 2
        synth.write_table(str(table), ref, counts, chrom="chrX", start=700, indel_cells=cells)
    T = readcounts.load_table(table.read_bytes(), n)
    prm = orc.call_params(pairs=pairs, extra_rob=1, max_prop_extra_rob=0.8)
    o = driver.parse_args(["--out_file=x", "--extra_rob=TRUE"])
    fasta = driver.read_fasta(str(fa))
    L = api.load_library()
    fh = L.ns_fasta_load(str(fa).encode())
    assert fh
    # --- SNV units
    O = orc.call_snv_batch(T.ref, np.ascontiguousarray(T.counts.transpose(1, 0, 2)), prm, n_threads=4)
    O["flags"] = (O["ref_inv"] * _abi.NS_F_REF_INV + O["extra_rob"] * _abi.NS_F_EXTRA_ROB).astype(np.uint32)
    units = [int(u) for u in np.where(O["gate2"] != 0)[0]]
    assert len(units) > 10
    R = _fake_result(O, units, n)
    txt, off = api.vcf_records(T, R, names, fh, 50.0, indel=False, n_threads=3)
    statuses = set()
    for e, u in enumerate(units):
        p, k = divmod(u, 3)
        r = int(T.ref[p])
        alt = [b for b in range(4) if b != r][k]
        c = T.counts[p]
        pos = int(T.loc[p])
        row = {kk: R[kk][e] for kk in ("gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled")}
        want = driver.vcf_line("chrX", pos, "ATCG"[r], "ATCG"[alt], "snv", row, c[0], c[1 + alt], c[5 + alt], c[1 + r],
                               c[5 + r], O["sigma"][u], O["slope"][u], int(O["flags"][u]), O["max_gq"][u], o, fasta, names,
                               driver.context(fasta, "chrX", pos - 3, pos - 1), driver.context(fasta, "chrX", pos + 1, pos + 3))
        assert txt[off[e]:off[e + 1]].decode() == want, (e, u)
        statuses.update(f.split(":")[11].split("_FROM_")[0] for f in want.rstrip("\n").split("\t")[9:])
    if tn:
        assert "POSSIBLE_CONTAMINATION" in statuses and "GERMLINE_CONFIRMED" in statuses
    # --- indel units (deletions and insertions) through the per-unit oracle
    dp, vp, vm, rp, rm, ind = T.indel_unit_rows()
    Ui = len(T.indel_seq)
    keys = ("pvalue", "qvalue", "gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status")
    Oi = {k: np.zeros((Ui, n), dtype=np.uint8 if k in ("gt", "status") else np.float64) for k in keys}
    Oi.update(sigma=np.zeros(Ui), slope=np.zeros(Ui), max_gq=np.zeros(Ui), flags=np.zeros(Ui, dtype=np.uint32),
              pooled=np.zeros((Ui, 8)))
    iunits = []
    for u in range(Ui):
        Q = orc.call_unit(dp[u], vp[u], vm[u], rp[u], rm[u], prm, is_indel=True)
        for k in keys:
            Oi[k][u] = Q[k]
        inf = Q["info"]
        Oi["sigma"][u], Oi["slope"][u], Oi["max_gq"][u] = inf.glm.sigma, inf.glm.slope, inf.max_gq
        Oi["pooled"][u] = list(inf.pooled)
        Oi["flags"][u] = inf.ref_inv * _abi.NS_F_REF_INV + inf.glm.extra_rob * _abi.NS_F_EXTRA_ROB
        if inf.gate2:
            iunits.append(u)
    assert len(iunits) > 4 and {int(T.indel_type[u]) for u in iunits} == {1, 2}
    Ri = _fake_result(Oi, iunits, n)
    itxt, ioff = api.vcf_records(T, Ri, names, fh, 50.0, indel=True, n_threads=2)
    for e, u in enumerate(iunits):
        p = int(T.indel_pos[u])
        pos, s = int(T.loc[p]), T.indel_seq[u]
        if T.indel_type[u] == 1:
            before = driver.context(fasta, "chrX", pos + 1 - 3, pos + 1 - 1).upper()
            after = driver.context(fasta, "chrX", pos + 1 + len(s), pos + 1 + 3 + len(s) - 1).upper()
            ref_txt, alt_txt, typ = before[2:3] + s, before[2:3], "del"
        else:
            before = driver.context(fasta, "chrX", pos - 2, pos).upper()
            after = driver.context(fasta, "chrX", pos + 1, pos + 3).upper()
            ref_txt, alt_txt, typ = before[2:3], before[2:3] + s, "ins"
        row = {kk: Ri[kk][e] for kk in ("gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled")}
        want = driver.vcf_line("chrX", pos, ref_txt, alt_txt, typ, row, dp[u], vp[u], vm[u], rp[u], rm[u], Oi["sigma"][u],
                               Oi["slope"][u], int(Oi["flags"][u]), Oi["max_gq"][u], o, fasta, names, before, after)
        assert itxt[ioff[e]:ioff[e + 1]].decode() == want, (e, u)
    # no FASTA: the context is N's, as many as the window asked for
    txt0, off0 = api.vcf_records(T, R, names, None, 50.0)
    assert b";CONT=NNNxNNN" in txt0[off0[0]:off0[1]]
    L.ns_fasta_free(fh)
    T.close()
