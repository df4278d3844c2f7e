#!/usr/bin/env python
"""Window labels of the bench's two-code sample (contigs of the first 100 Mbp whose true code is 11 or 15): label histogram,
contigs made of noise labels only (they stop the reference's smoothing with an IndexError), and a digest to compare runs.

    python tests/tools/cli_iso_probe.py [--shape2-first]
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from mgcod_b200.engine import Engine, synth_metagenome  # noqa: E402
from mgcod_b200.pipeline import Pipeline  # noqa: E402


def main():
    work = tempfile.mkdtemp()
    model_paths = bench.write_models(os.path.join(work, "models"))
    buf, offs, truth = synth_metagenome(bench.SEED, 1_000_000_000, threads=16)
    if "--shape2-first" in sys.argv:
        eng = Engine(model_paths)
        buf2, offs2, _ = synth_metagenome(bench.SEED + 1000, 300_000_000, min_len=5000, max_len=5000, threads=16)
        eng.decode_buffer(buf2, offs2)
        r = eng.decode_buffer(buf, offs)
        print("main engine genes", [r.num_genes(m) for m in range(4)])
    nrec = int(np.searchsorted(offs, 100_000_000, side="left"))
    two = [i for i in range(nrec) if int(truth[i]) in (11, 15)]
    fa = os.path.join(work, "sample2.fasta")
    with open(fa, "wb") as fh:
        for i in two:
            fh.write(b">contig_%d synthetic\n" % i)
            fh.write(buf[offs[i]:offs[i + 1]].tobytes() + b"\n")
    P = Pipeline(fa, os.path.join(work, "pred") + "/", "sample2", True, False, 3, 5, 5000, None, 30, False, False, work,
                 model_dir=os.path.dirname(model_paths[0]))
    os.makedirs(os.path.join(work, "pred"))
    try:
        P()
        print("pipeline ok; contigs with a switch:", sum(1 for v in P.predicted_switch_regions.values() if v and v[1]))
    except IndexError as ex:
        print("IndexError:", ex)
    labels = P.genome_label if isinstance(P.genome_label, list) else []
    allab = np.concatenate(labels) if labels else np.zeros(0, dtype=np.int64)
    print("contigs", len(two), "windows", allab.shape[0], "label histogram", np.bincount(allab, minlength=4).tolist())
    print("labels sha", hashlib.sha256(allab.astype(np.int64).tobytes()).hexdigest()[:16],
          "genes sha", [hashlib.sha256(df[["start", "end", "logodd"]].to_numpy().tobytes()).hexdigest()[:12] for df in P.label_to_gene_coords.values()])
    counts = np.bincount(allab, minlength=4)
    noise = set(np.argsort(counts)[::-1][2:].tolist())
    for k, (rec, l) in enumerate(zip(P.sequence, labels)):
        if len(l) and all(int(x) in noise for x in l):
            print("all-noise contig: index", k, rec.id, "len", len(rec), "labels", l.tolist(), "true code", int(truth[int(rec.id.split('_')[1])]))


if __name__ == "__main__":
    main()
