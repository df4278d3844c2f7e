#!/usr/bin/env python
"""Differential fuzz: the LIVE reference binary (/root/reference/dependencies/gmhmmp2, call form of
bin/data.py:40-41) against the CPU oracle on fresh synthetic metagenomes under the four parametric
82-section parameter files.  Build container only (the binary's checkout is absent on the GPU box).

    python tests/tools/fuzz_vs_binary.py [--seeds 987654 31337 4242] [--bp 4000000] [--jobs 8]

Compares the gene lines of the GFF text (contig, left, right, %.2f score, strand, attributes incl. gene_type and gc)
and prints every difference with the GC fraction of the gene, so that tie effects are visible.  The same comparison for the
CUDA engine runs on the GPU box against the staged binary: tests/test_live_binary.py.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GM = os.environ.get("GMHMMP2", "/root/reference/dependencies/gmhmmp2")
CODES = (11, 4, 15, 101)


def gene_lines(text):
    out = []
    for line in text.split("\n"):
        if line and not line.startswith("#"):
            t = line.split("\t")
            out.append((t[0], int(t[3]), int(t[4]), t[5], t[6], t[8].split("gene_id")[1].split(";", 1)[1].strip()))
    return out


def run_binary(mod, fasta, out):
    subprocess.run([GM, "-M", mod, "-s", fasta, "-f", "gff", "-o", out], check=True, capture_output=True)
    return open(out).read()


def oracle_lines(model, recs):
    out = []
    for rid, s in recs:
        genes, _ = model.decode(s)
        for g in genes:
            attr = f"gene_type {'archaea' if g['domain'] else 'bacteria'};"
            if g["partial_left"] or g["partial_right"]:
                attr += f" partial {g['partial_left']}{g['partial_right']};"
            attr += f" gc {g['gc']}; length {g['right'] - g['left'] + 1};"
            out.append((rid, g["left"], g["right"], "%.2f" % g["logodd"], "+" if g["strand"] > 0 else "-", attr))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[987654, 31337, 4242])
    ap.add_argument("--bp", type=int, default=4_000_000)
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--kind", default="parametric")
    args = ap.parse_args()
    from mgcod_b200 import synth
    from mgcod_b200.engine import synth_metagenome
    from oracle import gmhmm
    tmp = tempfile.mkdtemp()
    paths = synth.write_model_set(tmp, kind=args.kind)
    total_genes = total_diff = 0
    for seed in args.seeds:
        buf, offs, truth = synth_metagenome(seed, args.bp, threads=args.jobs)
        recs = [(f"contig_{i}", buf[offs[i]:offs[i + 1]].tobytes().decode()) for i in range(len(truth))]
        seqs = dict(recs)
        fa = os.path.join(tmp, f"s{seed}.fa")
        with open(fa, "w") as fh:
            for rid, s in recs:
                fh.write(f">{rid}\n{s}\n")
        with ThreadPoolExecutor(args.jobs) as ex:
            futs = {c: ex.submit(run_binary, paths[c], fa, os.path.join(tmp, f"s{seed}_{c}.gff")) for c in CODES}
            ofut = {c: ex.submit(oracle_lines, gmhmm.Model(paths[c]), recs) for c in CODES}
            for c in CODES:
                ref = gene_lines(futs[c].result())
                mine = ofut[c].result()
                total_genes += len(ref)
                if ref == mine:
                    continue
                sr, sm = set(ref), set(mine)
                for kind, rows in (("binary-only", sorted(sr - sm)), ("oracle-only", sorted(sm - sr))):
                    for r in rows:
                        s = seqs[r[0]][r[1] - 1:r[2]]
                        gcn = sum(ch in "GC" for ch in s)
                        print(f"seed {seed} code {c} {kind}: {r}  GC {gcn}/{len(s)} = {100.0 * gcn / len(s):.4f}")
                total_diff += len(sr - sm)
        print(f"seed {seed}: done, {total_genes} genes so far, {total_diff} differing", flush=True)
    print(f"TOTAL genes {total_genes} differing {total_diff}")
    return 1 if total_diff else 0


if __name__ == "__main__":
    sys.exit(main())
