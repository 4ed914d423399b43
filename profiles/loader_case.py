import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np
from needlestack_b200 import synth, readcounts
cfg = synth.CONFIGS["C2"]
P = 60000
ref, counts = synth.generate(cfg, P=P, seed=5)
path = '/tmp/c2_table.tsv'
synth.write_table(path, ref, counts, indel_cells=synth.random_indel_cells(P, cfg.n, seed=1))
text = open(path,'rb').read()
print('text MB', round(len(text)/1e6), 'cores', os.cpu_count())
for rep in range(4):
    t=time.time(); T = readcounts.load_table(text, cfg.n); dt=time.time()-t; pinned=T.pinned if hasattr(T,'pinned') else None; T.close()
    print('rep', rep, 's', round(dt,3), 'MB/s', round(len(text)/1e6/dt))
