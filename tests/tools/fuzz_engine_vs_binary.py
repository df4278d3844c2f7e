#!/usr/bin/env python
"""GPU box: the CUDA engine against the staged reference binary (oracle/_ref/gmhmmp2) on fresh synthetic metagenomes of any
size — the comparison of tests/test_live_binary.py as a tool.

    python tests/tools/fuzz_engine_vs_binary.py [--seeds 987654 31337 4242] [--bp 40000000]

Prints the number of genes compared and every differing gene line.  MGCOD_B200_CHAIN=thread / warp sends all chain pieces
through one of the two chain kernels (batches of this size use both by default)."""
import argparse
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
BINARY = os.path.join(ROOT, "oracle", "_ref", "gmhmmp2")
CODES = (11, 4, 15, 101)


def binary_lines(mod, fasta, out):
    subprocess.run([BINARY, "-M", mod, "-s", fasta, "-f", "gff", "-o", out], check=True, capture_output=True)
    rows = []
    for line in open(out):
        if line.startswith("#") or not line.strip():
            continue
        t = line.rstrip("\n").split("\t")
        rows.append((int(t[0].split("_")[1]), int(t[3]), int(t[4]), t[6], t[5], t[8].split(";", 1)[1].strip()))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[987654, 31337, 4242])
    ap.add_argument("--bp", type=int, default=40_000_000)
    a = ap.parse_args()
    from mgcod_b200 import synth
    from mgcod_b200.engine import Engine, synth_metagenome
    from tests import fixtures as fx
    tmp = tempfile.mkdtemp()
    synth.write_model_set(tmp)
    paths = [f"{tmp}/mgm_{c}.mod" for c in CODES]
    eng = Engine(paths)
    total = bad = 0
    for seed in a.seeds:
        buf, offs, truth = synth_metagenome(seed, a.bp, threads=8)
        fa = os.path.join(tmp, f"in_{seed}.fa")
        with open(fa, "wb") as fh:
            for i in range(len(truth)):
                fh.write(b">contig_%d\n" % i + buf[offs[i]:offs[i + 1]].tobytes() + b"\n")
        t0 = time.perf_counter()
        with ThreadPoolExecutor(4) as ex:
            ref = list(ex.map(lambda k: binary_lines(paths[k], fa, os.path.join(tmp, f"o{seed}_{k}.gff")), range(4)))
        t1 = time.perf_counter()
        res = eng.decode_buffer(buf, offs)
        for mi in range(4):
            mine = [(int(g["contig"]), int(g["left"]), int(g["right"]), "+" if g["strand"] > 0 else "-", "%.2f" % g["logodd"],
                     fx.gene_attr_text(g["domain"], g["partial_left"], g["partial_right"], g["gc"], int(g["right"]) - int(g["left"]) + 1))
                    for g in res.genes(mi)]
            diff = sorted(set(ref[mi]) ^ set(mine))
            total += len(ref[mi])
            bad += len(diff)
            for d in diff[:10]:
                print("DIFF", seed, CODES[mi], d)
        print(f"seed {seed}: {len(truth)} contigs, binary {t1 - t0:.1f} s, genes so far {total}, differing lines {bad}", flush=True)
    print(f"{total} genes compared, {bad} differing lines")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
