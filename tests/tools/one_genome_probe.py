"""Latency of decoding ONE long contig (a complete bacterial genome's worth) under the four models."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mgcod_b200 import synth
from mgcod_b200.engine import Engine, synth_metagenome

d = tempfile.mkdtemp()
synth.write_model_set(d)
eng = Engine([f"{d}/mgm_{c}.mod" for c in (11, 4, 15, 101)])
for L in (200_000, 1_000_000, 5_000_000, 20_000_000):
    buf, offs, _ = synth_metagenome(7, L, min_len=L, max_len=L, threads=8)
    for _ in range(3):
        r = eng.decode_buffer(buf, offs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        r = eng.decode_buffer(buf, offs)
    dt = (time.perf_counter() - t0) / 5
    st = eng.stats()
    print(f"{L:>9d} bp x 4 models: {1e3 * dt:7.2f} ms wall per decode_buffer; device phases: "
          + " ".join(f"{k}={st[k]:.2f}" for k in ("pack", "orf_find", "orf_prep", "orf_score", "orf_final", "order", "chain", "d2h", "total"))
          + f"; genes {st['n_genes']}")
