"""End to end through the drop-in driver on a GPU: text table -> loader -> GPU batch API -> VCF text, checked
field by field against the oracle's per-unit results formatted the way R's cat() prints them."""
import io
import os

import numpy as np
import pytest

import parity
from needlestack_b200 import _abi, driver, readcounts, synth

pytestmark = pytest.mark.gpu


def test_driver_end_to_end(tmp_path, oracle, monkeypatch):
    cfg = synth.CONFIGS["C1"]
    P = 500
    ref, counts = synth.generate(cfg, P=P, seed=31, p_variant=0.03, p_germline=0.004)
    cells = synth.random_indel_cells(P, cfg.n, seed=31, p_site=0.02)
    table = tmp_path / "t.tsv"
    synth.write_table(str(table), ref, counts, chrom="chr7", start=1000, indel_cells=cells)
    (tmp_path / "names.txt").write_text("".join(f"bam{i} S{i + 1}\n" for i in range(cfg.n)))
    rng = np.random.default_rng(1)
    seq = "".join(rng.choice(list("ACGT"), size=1000 + P + 10))
    (tmp_path / "ref.fa").write_text(">chr7 test\n" + "\n".join(seq[i:i + 60] for i in range(0, len(seq), 60)) + "\n")
    monkeypatch.chdir(tmp_path)
    out = tmp_path / "out.vcf"
    rc = driver.main(["--out_file=" + str(out), "--fasta_ref=" + str(tmp_path / "ref.fa"), "--chunk_lines=137"],
                     stdin=io.BytesIO(table.read_bytes()))
    assert rc == 0
    lines = out.read_text().splitlines()
    body = [ln.split("\t") for ln in lines if not ln.startswith("#")]
    assert lines[0] == "##fileformat=VCFv4.1" and lines[-len(body) - 1].startswith("#CHROM")
    # oracle on the same units
    prm = oracle.call_params()
    O = oracle.call_snv_batch(ref, np.ascontiguousarray(counts.transpose(1, 0, 2)), prm)
    want_snv = [int(u) for u in np.where(O["gate2"] != 0)[0]]
    snv = [r for r in body if "TYPE=snv" in r[7]]
    assert len(snv) == len(want_snv) > 3
    bases = "ATCG"
    for rec, u in zip(snv, want_snv):
        p, k = divmod(u, 3)
        alt = [b for b in range(4) if b != ref[p]][k]
        assert rec[0] == "chr7" and int(rec[1]) == 1000 + p and rec[3] == bases[ref[p]] and rec[4] == bases[alt]
        assert len(rec) == 9 + cfg.n
        info = dict(kv.split("=") for kv in rec[7].split(";"))
        assert info["CONT"] == seq[1000 + p - 4:1000 + p - 1] + "x" + seq[1000 + p:1000 + p + 3]
        for key, val in (("ERR", O["slope"][u]), ("SIG", O["sigma"][u])):
            assert abs(float(info[key]) - val) <= 2e-6 * abs(val) + 1e-12
        assert abs(float(rec[5]) - O["max_gq"][u]) <= 1e-5 * max(1, O["max_gq"][u])
        assert int(info["DP"]) == int(counts[p, 0].sum()) and int(info["NS"]) == int((counts[p, 0] > 0).sum())
        for i in range(cfg.n):
            f = rec[9 + i].split(":")
            assert len(f) == 12
            assert f[0] == _abi.NS_GT[O["gt"][u][i]]
            assert abs(float(f[1]) - O["gq"][u][i]) <= 1e-5 * max(1, O["gq"][u][i])
            assert int(f[2]) == counts[p, 0, i] and int(f[4]) == counts[p, 1 + alt, i] + counts[p, 5 + alt, i]
            assert f[11] == "."
    # indel records: one per (position, unique sequence) that passes the gates; deletions before insertions
    T = readcounts.load_table(table.read_bytes(), cfg.n)
    dp, vp, vm, rp, rm, ind = T.indel_unit_rows()
    want = []
    for u in range(len(T.indel_seq)):
        R = oracle.call_unit(dp[u], vp[u], vm[u], rp[u], rm[u], prm, is_indel=True)
        if R["info"].gate2:
            want.append((int(T.indel_pos[u]), "del" if T.indel_type[u] == 1 else "ins", T.indel_seq[u]))
    got = []
    for r in body:
        info = dict(kv.split("=") for kv in r[7].split(";"))
        if info["TYPE"] in ("del", "ins"):
            s = r[3][1:] if info["TYPE"] == "del" else r[4][1:]
            got.append((int(r[1]) - 1000, info["TYPE"], s))
            assert r[3][0] == r[4][0] == seq[int(r[1]) - 1]  # the base before the indel (prev_bp)
    assert got == want and len(want) > 2
    # records come out in the reference's order: by line, SNVs then deletions then insertions
    order = [(int(r[1]), {"snv": 0, "del": 1, "ins": 2}[dict(kv.split("=") for kv in r[7].split(";"))["TYPE"]]) for r in body]
    assert order == sorted(order)
