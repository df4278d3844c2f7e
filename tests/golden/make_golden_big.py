#!/usr/bin/env python
"""Large multi-section golden from the REFERENCE BINARY (build container only).

    python tests/golden/make_golden_big.py

Runs /root/reference/dependencies/gmhmmp2 (call form of bin/data.py:40-41: -M <mod> -s <fasta> -f gff -o <out>) on a
24 Mbp synthetic metagenome (the generator of the timed workload, BASELINE.json config 5) under the four parametric
82-section parameter files — the configuration bench.py times — and stores in big_parametric.json.xz

  seed / total_bp / seq_sha256 / model_sha256     what has to be regenerated to compare
  digests[code][contig]                            first 16 hex digits of sha256 over the contig's gene lines
                                                   "left right strand %.2f attributes\n"
  n_genes[code]
  ties[]                                           every predicted gene whose 100*GC/len is an exact .5 tie, with the
                                                   binary's gc / score / gene_type: the binary computes
                                                   (int)((double)GC/len*100 + 0.5), which sends most ties DOWN
                                                   (round 1 used integer round-half-up: 1.6e-4 of all genes wrong)

The 150 kb golden of round 1 (synth_parametric.json) held 0.14 expected tie genes; this one holds several hundred.
"""
import hashlib
import json
import lzma
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
GM = "/root/reference/dependencies/gmhmmp2"
CODES = (11, 4, 15, 101)
SEED, TOTAL_BP = 424242, 24_000_000

from mgcod_b200 import synth  # noqa: E402
from mgcod_b200.engine import synth_metagenome  # noqa: E402


def contig_digests(rows_by_contig, n_contigs):
    out = []
    for c in range(n_contigs):
        h = hashlib.sha256()
        for r in rows_by_contig.get(c, ()):
            h.update(("%d %d %s %s %s\n" % r).encode())
        out.append(h.hexdigest()[:16])
    return out


def main():
    tmp = tempfile.mkdtemp()
    paths = synth.write_model_set(tmp, kind="parametric")
    buf, offs, truth = synth_metagenome(SEED, TOTAL_BP, threads=8)
    n = len(truth)
    fa = os.path.join(tmp, "big.fa")
    with open(fa, "w") as fh:
        for i in range(n):
            fh.write(f">contig_{i}\n{buf[offs[i]:offs[i + 1]].tobytes().decode()}\n")

    def run(code):
        out = os.path.join(tmp, f"o{code}.gff")
        subprocess.run([GM, "-M", paths[code], "-s", fa, "-f", "gff", "-o", out], check=True, capture_output=True)
        rows = {}
        for line in open(out):
            if line.startswith("#") or not line.strip():
                continue
            t = line.rstrip("\n").split("\t")
            attrs = t[8].split(";", 1)[1].strip()          # without the running gene_id
            rows.setdefault(int(t[0].split("_")[1]), []).append((int(t[3]), int(t[4]), t[6], t[5], attrs))
        return rows

    with ThreadPoolExecutor(4) as ex:
        res = dict(zip(CODES, ex.map(run, CODES)))
    meta = {"seed": SEED, "total_bp": TOTAL_BP, "n_contigs": n, "seq_sha256": hashlib.sha256(buf.tobytes()).hexdigest(),
            "model_sha256": {str(c): hashlib.sha256(open(paths[c], "rb").read()).hexdigest() for c in CODES},
            "digests": {}, "n_genes": {}, "ties": []}
    isgc = (buf == ord("G")) | (buf == ord("C"))
    import numpy as np
    gcpre = np.concatenate([[0], np.cumsum(isgc, dtype=np.int64)])
    for code in CODES:
        rows = res[code]
        meta["digests"][str(code)] = contig_digests(rows, n)
        meta["n_genes"][str(code)] = sum(len(v) for v in rows.values())
        for c, rr in rows.items():
            for left, right, strand, score, attrs in rr:
                L = right - left + 1
                g = int(gcpre[offs[c] + right] - gcpre[offs[c] + left - 1])
                if (200 * g) % L == 0 and (100 * g) % L != 0:
                    gc = int(attrs.split("gc ")[1].split(";")[0])
                    meta["ties"].append([code, c, left, right, strand, score, gc, attrs.split("gene_type ")[1].split(";")[0],
                                         g, L])
    down = sum(1 for t in meta["ties"] if t[6] * t[9] * 2 < 200 * t[8])
    print(f"{n} contigs, genes {meta['n_genes']}, {len(meta['ties'])} tie genes ({down} rounded down by the binary)")
    with lzma.open(os.path.join(HERE, "big_parametric.json.xz"), "wt", preset=9) as fh:
        json.dump(meta, fh, separators=(",", ":"))


if __name__ == "__main__":
    main()
