"""Columnar loader (ns_table_parse, C++) against a line-by-line Python restatement of the reference's
parsing (bin/needlestack.r:316-329, :461-479, :624-643), and the host pieces of the drop-in driver."""
import numpy as np
import pytest

from needlestack_b200 import driver, readcounts, synth


def _table(tmp_path, P=300, seed=3, n=None, with_indels=True, crlf=False):
    cfg = synth.CONFIGS["C1"]
    ref, counts = synth.generate(cfg, P=P, seed=seed, p_variant=0.02)
    ref[::41] = 4
    cells = synth.random_indel_cells(P, cfg.n, seed=seed, p_site=0.05) if with_indels else None
    path = tmp_path / "t.tsv"
    synth.write_table(str(path), ref, counts, indel_cells=cells)
    data = path.read_bytes()
    if crlf:
        data = data.replace(b"\n", b"\r\n")
    return cfg, ref, counts, data


@pytest.mark.parametrize("threads,crlf", [(1, False), (4, False), (3, True)])
def test_loader_matches_reference_parsing(tmp_path, threads, crlf):
    cfg, ref, counts, data = _table(tmp_path, crlf=crlf)
    T = readcounts.load_table(data, cfg.n, n_threads=threads)
    R = readcounts.load_table_python(data, cfg.n)
    assert T.P == len(ref) and np.array_equal(T.counts, counts) and np.array_equal(T.ref, ref)
    assert np.array_equal(T.counts, R["counts"]) and np.array_equal(T.ref, R["ref"])
    assert np.array_equal(T.loc, R["loc"]) and [T.chrom_names[c] for c in T.chrom] == R["chrom"]
    assert len(T.indel_seq) == len(R["units"]) > 10
    for u, (p, typ, seq, vp, vm) in enumerate(R["units"]):
        assert (int(T.indel_pos[u]), int(T.indel_type[u]), T.indel_seq[u]) == (p, typ, seq)
        assert np.array_equal(T.indel_vp[u], vp) and np.array_equal(T.indel_vm[u], vm)
    # skipped reference bases contribute no indel units (needlestack.r:319)
    assert not np.any(ref[T.indel_pos] > 3)
    dp, vp, vm, rp, rm, ind = T.indel_unit_rows()
    p0 = int(T.indel_pos[0])
    assert np.array_equal(dp[0], counts[p0, 0]) and np.array_equal(rp[0], counts[p0, 1 + ref[p0]])
    T.close()


def test_loader_large_chunking_is_thread_independent(tmp_path):
    cfg, ref, counts, data = _table(tmp_path, P=3000, seed=8)
    A = readcounts.load_table(data, cfg.n, n_threads=1)
    B = readcounts.load_table(data, cfg.n, n_threads=7)
    assert np.array_equal(A.counts, B.counts) and A.indel_seq == B.indel_seq
    assert np.array_equal(A.indel_vp, B.indel_vp) and np.array_equal(A.indel_pos, B.indel_pos)


def test_loader_errors():
    from needlestack_b200 import NeedlestackError
    with pytest.raises(NeedlestackError, match="too few columns"):
        readcounts.load_table(b"h\nchr1\t1\tA\t10\t1\n", 1)
    with pytest.raises(NeedlestackError, match="non-integer"):
        readcounts.load_table(b"h\nchr1\t1\tA\t10\tx\t0\t0\t0\t0\t0\t0\t0\tNA\tNA\n", 1)
    with pytest.raises(NeedlestackError, match="indel"):
        readcounts.load_table(b"h\nchr1\t1\tA\t10\t1\t0\t0\t0\t0\t0\t0\t0\tNA\tAT\n", 1)
    T = readcounts.load_table(b"h\n", 3)
    assert T.P == 0


def test_r_number_formatting():
    f = driver.rfmt
    # R: cat(x) with options(digits = 7, scipen = 100)
    for x, want in ((0.0001234567891, "0.0001234568"), (1e-15, "0.000000000000001"), (123456789.123, "123456789"),
                    (1000.0, "1000"), (0.1 + 0.2, "0.3"), (175.08203174, "175.082"), (2 / 3, "0.6666667"),
                    (1234567.891, "1234568"), (99.99999999, "100"), (-0.0, "0"), (1e15, "1000000000000000"),
                    (0.5, "0.5"), (float("inf"), "Inf"), (-1.0, "-1"), (99.0, "99"), (1000.0000001, "1000")):
        assert f(x) == want, (x, f(x), want)
    assert f(float("nan")) == "NA" and f(float("nan"), nan="NaN") == "NaN" and f(7) == "7"


def test_cli_defaults_and_header(tmp_path):
    o = driver.parse_args(["--out_file=o.vcf", "--fasta_ref=ref.fa", "--min_reads=5", "--extra_rob=TRUE"])
    assert (o["min_coverage"], o["min_reads"], o["GQ_threshold"], o["SB_type"], o["max_prop_extra_rob"]) == \
        (30.0, 5.0, 50.0, "SOR", 0.8)  # bin/needlestack.r:64-92
    assert o["extra_rob"] is True and o["output_all_SNVs"] is False and o["afmin_power"] == -1.0
    with pytest.raises(SystemExit):
        driver.parse_args(["--fasta_ref=x"])
    h = driver.vcf_header(o, ["A", "B"]).splitlines()
    assert h[0] == "##fileformat=VCFv4.1" and h[2] == "##source=needlestack v1.1"
    assert h[5] == '##filter="QVAL > 50 & SOR_SNV < 100 & SOR_INDEL < 100 & min(AO) >= 5 & min(AF) >= 0 & min(DP) >= 30"'
    assert sum(1 for ln in h if ln.startswith("##FORMAT=<ID=FS,")) == 2  # the reference's duplicate line is kept
    assert h[-1] == "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB"
    (tmp_path / "names.txt").write_text("f1 S\nf2 S\nf3 T\n")
    _, names = driver.read_names(str(tmp_path / "names.txt"))
    assert names == ["S", "S_1", "T"]  # make.unique(sep = "_")
    (tmp_path / "pairs.txt").write_text("TUMOR NORMAL\nS S_1\nT NA\n")
    pr = driver.read_pairs(str(tmp_path / "pairs.txt"), names)
    assert pr == dict(tindex=[0], nindex=[1], only_t=[2], only_n=[])


def test_r_number_formatting_fast_path_equals_general_path():
    """rfmt's %g shortcut (1e-4 <= |x| < 1e7) against the general construction on random values and the edges"""
    import random
    from needlestack_b200 import driver

    def general(x):
        mant, exp = f"{x:.6e}".split("e")
        nsig = len(mant.replace("-", "").replace(".", "").rstrip("0")) or 1
        return f"{x:.{max(0, nsig - 1 - int(exp))}f}"

    random.seed(3)
    vals = [10 ** random.uniform(-4, 7) * random.choice([1, -1]) for _ in range(200000)]
    vals += [round(v, random.randint(0, 6)) for v in vals[:50000]]
    vals += [0.0001, 9999999.4, 9999999.5, 9999999.999, 0.00010000004, 99999.995, 0.30000000000000004, 1234567.0, 5e-5,
             123.4567891, 1e6, 0.1]
    for x in vals:
        if x != 0:
            assert driver.rfmt(x) == general(x), x
