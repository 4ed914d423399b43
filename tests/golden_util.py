"""helpers shared by the CPU and GPU golden-fixture tests"""
import importlib.util
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN_DIR, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = make_golden.CASES


def load(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: d[k] for k in d.files}


def case_params(name):
    from needlestack_b200 import synth
    cfgname, P, seed, pv, pg, kw, tn = CASES[name]
    pairs = synth.tn_pairs(synth.CONFIGS[cfgname]) if tn else None
    return kw, pairs
