"""The `Rscript needlestack.r` step end to end through the drop-in driver: text table (file on STDIN) -> loader -> GPU ->
VCF text, on a C2-shaped table (100 samples) WITH indel cells.  Usage: python profiles/driver_case.py [positions] [profile]"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from needlestack_b200 import driver, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
prof = len(sys.argv) > 2 and sys.argv[2] == "profile"
cfg = synth.CONFIGS["C2"]
ref, counts = synth.generate(cfg, P=P, seed=5)
d = tempfile.mkdtemp()
os.chdir(d)
synth.write_table("t.tsv", ref, counts, indel_cells=synth.random_indel_cells(P, cfg.n, seed=1, p_site=0.002))
open("names.txt", "w").write("".join(f"bam{i} S{i + 1}\n" for i in range(cfg.n)))
size = os.path.getsize("t.tsv")
for rep in range(3):
    with open("t.tsv", "rb") as f:
        t = time.time()
        driver.main(["--out_file=out.vcf"], stdin=f)
        dt = time.time() - t
    nrec = sum(1 for ln in open("out.vcf") if not ln.startswith("#"))
    print(f"rep {rep}: {size / 1e6:.0f} MB of table text, {P} positions x {cfg.n} samples, {nrec} VCF records, "
          f"{dt:.2f} s = {size / 1e6 / dt:.0f} MB/s, {3 * P / dt / 1e6:.2f} M site-alleles/s")
if prof:
    pr = cProfile.Profile()
    f = open("t.tsv", "rb")
    pr.enable()
    driver.main(["--out_file=out.vcf"], stdin=f)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30)
    print(s.getvalue()[:7000])
