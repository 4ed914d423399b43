"""The `Rscript needlestack.r` step end to end through the drop-in driver: text table on STDIN -> loader -> GPU ->
VCF text, on a C2-shaped table (100 samples).  Usage: python profiles/driver_case.py [positions] [chunk_lines]"""
import io
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from needlestack_b200 import driver, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = synth.CONFIGS["C2"]
ref, counts = synth.generate(cfg, P=P, seed=5)
d = tempfile.mkdtemp()
os.chdir(d)
synth.write_table("t.tsv", ref, counts, indel_cells=synth.random_indel_cells(P, cfg.n, seed=1, p_site=0.002))
open("names.txt", "w").write("".join(f"bam{i} S{i + 1}\n" for i in range(cfg.n)))
text = open("t.tsv", "rb").read()
for rep in range(2):
    t = time.time()
    driver.main(["--out_file=out.vcf", f"--chunk_lines={chunk}"], stdin=io.BytesIO(text))
    dt = time.time() - t
    nrec = sum(1 for ln in open("out.vcf") if not ln.startswith("#"))
    print(f"rep {rep}: {len(text) / 1e6:.0f} MB of table text, {P} positions x {cfg.n} samples, {nrec} VCF records, "
          f"{dt:.2f} s = {len(text) / 1e6 / dt:.0f} MB/s, {3 * P / dt / 1e6:.2f} M site-alleles/s")
