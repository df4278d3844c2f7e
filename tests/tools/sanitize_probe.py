"""Small decodes that touch every kernel path (clean, ambiguous letters + gaps, repeats overflow, windows, recoded stops); run under
compute-sanitizer:  compute-sanitizer --tool memcheck python tests/tools/sanitize_probe.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from mgcod_b200 import synth
from mgcod_b200.engine import Engine, synth_metagenome

d = tempfile.mkdtemp()
paths = synth.write_model_set(d, kind="parametric", bins=range(30, 71, 10))
eng = Engine([paths[c] for c in (11, 4, 15, 101)])
buf, offs, truth = synth_metagenome(3, 40000, threads=2)
res = eng.decode_buffer(buf, offs)
print("clean", [res.num_genes(m) for m in range(4)])
lens = np.diff(offs)
wo, wl, wc = eng.window_labels(lens, np.full(len(lens), 2500, np.int32), 5000, 3, True)
ml, st = eng.merge_labels_flat(wl, wo, 3)
c, o, p = eng.recoded_stops(res.genes(1))
print("windows", int(wo[-1]), "recoded", int(c.sum()))
seqs = [buf[offs[i]:offs[i + 1]].tobytes() for i in range(len(truth))]
s0 = bytearray(seqs[0]); s0[1000:1100] = b"N" * 100; s0[3000] = ord("R"); s0[-30:] = b"N" * 30
res = eng.decode([bytes(s0), b"", b"ACG", seqs[1][:95], b"ATG" * 3000 + b"TAA"])
print("amb", [res.num_genes(m) for m in range(4)])
for r in eng.decode_stream([(buf, offs), (buf, offs)]):
    pass
print("done")
eng.close()
