#!/usr/bin/env python
"""Timing of the GFF writer alone (host code, runs without a GPU): four files of ~132 k genes each written side by side, as
bench.py's file_e2e leg and Data.run_mgm do.   python tests/tools/writer_bench.py [genes_per_model] [contigs]"""
import os, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from mgcod_b200.engine import Result, GENE_DTYPE

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 132_000
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 4297
rng = np.random.default_rng(1)
genes = []
for m in range(4):
    g = np.zeros(ng, GENE_DTYPE)
    g["contig"] = np.sort(rng.integers(0, nc, ng))
    g["left"] = rng.integers(1, 400_000, ng)
    g["right"] = g["left"] + rng.integers(90, 3000, ng)
    g["strand"] = rng.choice([-1, 1], ng)
    g["gc"] = rng.integers(25, 70, ng)
    g["logodd"] = rng.random(ng) * 200
    genes.append(g)
totals = [rng.random(nc) * 1000 for _ in range(4)]
r = Result.from_genes(genes, totals)
ids = [f"contig_{i}" for i in range(nc)]
lens = np.full(nc, 500_000, np.int64)
d = tempfile.mkdtemp()
for rep in range(4):
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=4) as tp:
        for j in [tp.submit(r.write, mi, ids, lens, "x.fa", "m.mod", code, os.path.join(d, f"mgm_{code}.gff"), "gff")
                  for mi, code in enumerate((11, 4, 15, 101))]:
            j.result()
    dt = time.perf_counter() - t0
    size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
    print(f"4 files, {size / 1e6:.1f} MB in {1e3 * dt:.1f} ms ({size / dt / 1e9:.2f} GB/s)")
