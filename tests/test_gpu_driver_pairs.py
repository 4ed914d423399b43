"""The drop-in driver in tumour/normal mode (--pairs_file), end to end on a GPU: every VCF record against the SAME
formatter (driver.vcf_line, R's cat() number formatting) fed with the ORACLE's values for that unit.  Text fields
(CHROM, POS, REF, ALT, TYPE, CONT, WARN=..., FORMAT, GT, integer counts, SB, STATUS incl.
POSSIBLE_CONTAMINATION_FROM_<samples>) must be byte-identical; a printed number may differ from the oracle's in its
7th significant digit only when the two values straddle a rounding boundary (counted, bounded).
Reference: bin/needlestack.r:162-185 (pairs file), :342-373 (status), :396-413 (VCF line)."""
import io

import numpy as np
import pytest

from needlestack_b200 import _abi, driver, synth

pytestmark = pytest.mark.gpu

N, NP = 40, 18  # samples; pairs (T i <-> N 20+i), then T-only 18, 19 and N-only 38, 39
PAIRS = dict(tindex=list(range(NP)), nindex=list(range(20, 20 + NP)), only_t=[18, 19], only_n=[38, 39])
BASES = "ATCG"


def _plant(counts, ref, p, k, sample, af):
    """give `sample` an allelic fraction af of the k-th alt allele at position p (REF counts shrink accordingly)"""
    r = int(ref[p])
    alt = [b for b in range(4) if b != r][k]
    dp = int(counts[p, 0, sample])
    others = sum(int(counts[p, 1 + b, sample] + counts[p, 5 + b, sample]) for b in range(4) if b not in (r, alt))
    y = min(int(round(af * dp)), dp - others)
    counts[p, 1 + alt, sample], counts[p, 5 + alt, sample] = y // 2, y - y // 2
    rest = dp - others - y
    counts[p, 1 + r, sample], counts[p, 5 + r, sample] = rest // 2, rest - rest // 2


def _batch():
    cfg = synth.Config(**{**synth.CONFIGS["C4"].__dict__, "n": N, "tn_pairs": 20})
    P = 420
    ref, counts = synth.generate(cfg, P=P, seed=77, p_variant=0.08, p_germline=0.04, p_refswap=0.02)
    # contamination pattern: a germline variant confirmed in one pair + a tumour-only signal in another pair, same allele
    for j, p in enumerate(range(7, P, 29)):
        ref[p] = j % 4
        for s in range(N):
            for k in range(3):
                _plant(counts, ref, p, k, s, 0.0)
        _plant(counts, ref, p, 1, 2, 0.5), _plant(counts, ref, p, 1, 22, 0.5)      # GERMLINE_CONFIRMED
        _plant(counts, ref, p, 1, 5 + j % 8, 0.3)                                 # SOMATIC -> POSSIBLE_CONTAMINATION
        if j % 3 == 0:
            _plant(counts, ref, p, 1, 19, 0.4)                                    # T without a normal: UNKNOWN
        if j % 3 == 1:
            _plant(counts, ref, p, 1, 38, 0.5)                                    # N without a tumour: GERMLINE_UNCONFIRMABLE
    # reference-inverted positions (the genome carries the minor allele: ALT/DP > 0.8 in most samples): make a few
    # samples heterozygous, so that the regression on the REF counts has something to call (WARN=INV_REF); with eight
    # heterozygous samples the extra-robust filter removes them first (WARN=EXTRA_ROBUST_GL/INV_REF)
    done = 0
    for p in range(P):
        r = int(ref[p])
        for k, alt in enumerate([b for b in range(4) if b != r]):
            frac = (counts[p, 1 + alt] + counts[p, 5 + alt]) / np.maximum(counts[p, 0], 1)
            if (frac > 0.8).sum() > N // 2 and (p - 7) % 29:
                het = (3, 23) if done % 2 == 0 else (3, 23, 6, 26, 9, 29, 11, 31)
                for s_ in het:
                    _plant(counts, ref, p, k, s_, 0.5)
                done += 1
    assert done >= 4
    return cfg, ref, counts


def _expected_records(oracle, ref, counts, names, o, seq, start, extra_rob):
    prm = oracle.call_params(pairs={k: np.array(v, dtype=np.int32) for k, v in PAIRS.items()}, extra_rob=extra_rob,
                             max_prop_extra_rob=0.8)
    O = oracle.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), prm, n_threads=8)
    recs = []
    for u in np.where(O["gate2"] != 0)[0]:
        p, k = divmod(int(u), 3)
        r = int(ref[p])
        alt = [b for b in range(4) if b != r][k]
        c = counts[p]
        pos = start + p
        flags = (_abi.NS_F_REF_INV if O["ref_inv"][u] else 0) | (_abi.NS_F_EXTRA_ROB if O["extra_rob"][u] else 0)
        row = {kk: O[kk][u] for kk in ("gq", "qval_minaf", "sor", "rvsb", "fs", "gt", "status", "pooled")}
        recs.append(driver.vcf_line("chr3", pos, BASES[r], BASES[alt], "snv", row, c[0], c[1 + alt], c[5 + alt], c[1 + r],
                                    c[5 + r], O["sigma"][u], O["slope"][u], flags, O["max_gq"][u], o, None, names,
                                    seq[pos - 4:pos - 1], seq[pos:pos + 3]))
    return recs, O


def _is_number(t):
    try:
        float(t)
        return True
    except ValueError:
        return False


@pytest.mark.parametrize("extra_rob", [0, 1])
def test_driver_with_pairs_file_matches_oracle_fed_formatter(tmp_path, oracle, monkeypatch, extra_rob):
    cfg, ref, counts = _batch()
    P = len(ref)
    start = 5000
    table = tmp_path / "t.tsv"
    synth.write_table(str(table), ref, counts, chrom="chr3", start=start)
    names = [f"P{i:02d}{'T' if i < 20 else 'N'}" for i in range(N)]
    (tmp_path / "names.txt").write_text("".join(f"bam{i} {nm}\n" for i, nm in enumerate(names)))
    rows = ["TUMOR\tNORMAL"] + [f"{names[t]}\t{names[n]}" for t, n in zip(PAIRS["tindex"], PAIRS["nindex"])]
    rows += [f"{names[t]}\tNA" for t in PAIRS["only_t"]] + [f"NA\t{names[n]}" for n in PAIRS["only_n"]]
    (tmp_path / "pairs.txt").write_text("\n".join(rows) + "\n")
    rng = np.random.default_rng(3)
    seq = "".join(rng.choice(list("ACGT"), size=start + P + 10))
    (tmp_path / "ref.fa").write_text(">chr3\n" + "\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + "\n")
    monkeypatch.chdir(tmp_path)
    out = tmp_path / "out.vcf"
    argv = ["--out_file=" + str(out), "--fasta_ref=" + str(tmp_path / "ref.fa"), "--pairs_file=pairs.txt",
            "--chunk_lines=150"] + (["--extra_rob=TRUE"] if extra_rob else [])
    assert driver.main(argv, stdin=io.BytesIO(table.read_bytes())) == 0
    got = [ln + "\n" for ln in out.read_text().splitlines() if not ln.startswith("#")]
    o = driver.parse_args(argv)
    o["afmin_power"] = 0.01  # bin/needlestack.r:166
    want, O = _expected_records(oracle, ref, counts, names, o, seq, start, extra_rob)
    assert len(got) == len(want) >= 40
    n_num = n_diff = 0
    seen_status, seen_warn = set(), set()
    for g, w in zip(got, want):
        gf, wf = g.rstrip("\n").split("\t"), w.rstrip("\n").split("\t")
        assert len(gf) == len(wf) == 9 + N
        assert gf[:5] == wf[:5] and gf[6] == wf[6] and gf[8] == wf[8]
        gi, wi = gf[7].split(";"), wf[7].split(";")
        assert [x.split("=")[0] for x in gi] == [x.split("=")[0] for x in wi]
        toks = [(gf[5], wf[5])] + [(a.split("=", 1)[1], b.split("=", 1)[1]) for a, b in zip(gi, wi)]
        for a, b in zip(gf[9:], wf[9:]):
            fa, fb = a.split(":"), b.split(":")
            assert len(fa) == len(fb) == 12
            toks += list(zip(fa, fb))
            seen_status.add(fb[11].split("_FROM_")[0])
        seen_warn.update(x for x in wi if x.startswith("WARN="))
        for a, b in toks:
            if a == b:
                n_num += _is_number(b)
                continue
            # only a printed number may differ, and only by a unit of R's 7th significant digit
            assert _is_number(a) and _is_number(b), (a, b)
            n_num += 1
            n_diff += 1
            # (1e-9 absolute: -10 log10 of a Fisher p-value that equals 1 up to rounding prints as +-1e-15 with scipen = 100)
            assert abs(float(a) - float(b)) <= 2e-6 * abs(float(b)) + 1e-9, (a, b)
    # sigma is only defined to the 1e-8 (absolute) of the reference's Brent search: relative differences of ~1e-9 in the
    # printed values straddle a 7-digit rounding boundary about once in a hundred numbers
    assert n_diff <= 0.02 * n_num, (n_diff, n_num)
    # the statuses and warnings the case was built for are all there
    assert {"SOMATIC", "GERMLINE_CONFIRMED", "POSSIBLE_CONTAMINATION", "UNKNOWN", "GERMLINE_UNCONFIRMABLE", "."} <= seen_status
    contam = [f for g in got for f in g.split("\t")[9:] if "POSSIBLE_CONTAMINATION_FROM_" in f]
    # ... FROM_<every sample with a GERMLINE* / UNKNOWN status>, in sample order: the confirmed pair is always there
    assert contam and all({"P02T", "P22N"} <= set(f.rstrip("\n").split("FROM_")[1].split("_")) for f in contam)
    assert "WARN=INV_REF" in seen_warn
    if extra_rob:
        assert {"WARN=EXTRA_ROBUST_GL", "WARN=EXTRA_ROBUST_GL/INV_REF"} <= seen_warn
    print(f"extra_rob={extra_rob}: {len(got)} records, {n_num} printed numbers, {n_diff} differ in the last digit; "
          f"statuses {sorted(seen_status)}; warnings {sorted(seen_warn)}")
