import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from mgcod_b200 import synth
from mgcod_b200.engine import Engine, synth_metagenome
import tempfile
d = tempfile.mkdtemp()
paths = synth.write_model_set(d, kind='parametric')
L = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
buf, offs, truth = synth_metagenome(5, L, min_len=L, max_len=L, threads=4)
e = Engine([paths[c] for c in (11,4,15,101)])
for _ in range(4):
    r = e.decode_buffer(buf, offs)
print(e.stats())
