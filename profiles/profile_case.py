"""Small fixed case for ncu: one ns_call_snv over a C2-shaped batch (100 samples) with enough
common germline sites that k_fit_heavy has work.  Usage: python profiles/profile_case.py [P] [p_germline] [config]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import needlestack_b200 as ns
from needlestack_b200 import synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pg = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
cfgname = sys.argv[3] if len(sys.argv) > 3 else "C2"
ref, counts = synth.generate(synth.CONFIGS[cfgname], P=P, seed=99, p_germline=pg)
with ns.Caller(0) as c:
    for rep in range(2):
        t = time.time()
        R = c.call_snv(ref, counts)
        dt = time.time() - t
        tm = c.timings()
        print(f"rep {rep}: {R.n_fitted} fitted, {tm['n_heavy']} heavy, {R.n_emitted} emitted, wall {dt*1e3:.1f} ms, "
              f"gate {tm['gate_ms']:.2f} fit {tm['fit_ms']:.2f} heavy {tm['heavy_ms']:.2f} post {tm['post_ms']:.2f} ms")
