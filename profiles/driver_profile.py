"""Where the drop-in driver spends its time: cProfile of driver.main on a C2-shaped text table (tiled), plus wall-clock
stage timings.  Usage: python profiles/driver_profile.py [distinct positions] [tile]"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from needlestack_b200 import api, driver, readcounts, synth

P0 = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg = synth.CONFIGS["C2"]
ref, counts = synth.generate(cfg, P=P0, seed=5)
d = tempfile.mkdtemp()
os.chdir(d)
synth.write_table("t.tsv", ref, counts)
open("names.txt", "w").write("".join(f"bam{i} S{i + 1}\n" for i in range(cfg.n)))
raw = open("t.tsv", "rb").read()
nl = raw.index(b"\n") + 1
text = raw[:nl] + raw[nl:] * tile
open("tiled.tsv", "wb").write(text)
P = P0 * tile
print(f"table: {len(text) / 1e9:.2f} GB, {P} positions x {cfg.n} samples")
t = time.perf_counter()
c = api.Caller(0)
print(f"Caller(0) creation: {time.perf_counter() - t:.3f} s")
c.close()
for rep in range(3):
    with open("tiled.tsv", "rb") as f:
        t = time.perf_counter()
        driver.main(["--out_file=out.vcf"], stdin=f)
        dt = time.perf_counter() - t
    print(f"rep {rep}: {dt:.3f} s = {len(text) / 1e9 / dt:.2f} GB/s of text, {3 * P / dt / 1e6:.2f} M site-alleles/s")
t = time.perf_counter()
T = readcounts.load_table(text, cfg.n)
print(f"load_table alone: {time.perf_counter() - t:.3f} s")
with api.Caller(0) as c:
    prm = api.make_params()
    for rep in range(2):
        t = time.perf_counter()
        S = c.call_snv(T.ref, T.counts, prm=prm, per_unit=False)
        print(f"call_snv (rows only) rep {rep}: {time.perf_counter() - t:.3f} s, {S.n_emitted} rows")
    t = time.perf_counter()
    txt, off = api.vcf_records(T, S, [f"S{i}" for i in range(cfg.n)], None, 50.0)
    print(f"vcf_records: {time.perf_counter() - t:.3f} s, {len(txt) / 1e6:.1f} MB")
T.close()
pr = cProfile.Profile()
f = open("tiled.tsv", "rb")
pr.enable()
driver.main(["--out_file=out.vcf"], stdin=f)
pr.disable()
f.close()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25)
print(s.getvalue()[:6000])
