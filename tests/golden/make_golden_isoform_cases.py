#!/usr/bin/env python
"""Differential cases for the isoform comparison: random pairs of gene sets run through the UNMODIFIED reference
``Pipeline.compare_sets_of_predicted_genes`` (bin/pipeline.py:559-951) in the build container, with the pandas-1 /
Biopython stand-ins of make_golden_isoforms.py.  Each case perturbs one tiling of genes two ways so that every branch of
the comparison is hit many times: identical genes (with score ties), shifted starts, shifted stops inside and outside
the tolerance, set-2 genes spanning several set-1 genes and the reverse (with sums that tie), strand flips, genes
private to one set, before the first and after the last gene of the other.

Writes iso_cases.json.xz: inputs and the reference's two tables (or the exception type where the reference fails).

    python tests/golden/make_golden_isoform_cases.py
"""
import json
import os
import random
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_isoforms as MI   # noqa: E402
import make_golden_results as G     # noqa: E402

COLS = ["sequence", "source", "feature", "start", "end", "logodd", "strand", "phase", "attributes"]


def frame(genes, code):
    rows = [["c1", f"GeneMark.hmm2_{code}", "CDS", s, e, lo, st, 0, f"gene_id {k + 1}; gene_type bacteria; gc 40; length {e - s + 1};"]
            for k, (s, e, st, lo) in enumerate(genes)]
    df = pd.DataFrame(rows, columns=COLS)
    for c, t in (("start", "int64"), ("end", "int64"), ("logodd", "float64"), ("phase", "int64")):
        df[c] = df[c].astype(t)
    df["length"] = df["end"] - df["start"] + 1
    df["sequence"] = df["sequence"].astype(str)
    df["contig"] = list(df.sequence.values)
    return df


def case_sequence(seed, length):
    """the contig of a case (only its GC content matters: parents of isoform pairs get a fresh ``gc`` attribute)"""
    return "".join("ACGT"[k] for k in np.random.RandomState(seed).randint(0, 4, length))


def make_case(rng):
    n = rng.randint(3, 14)
    pos = rng.randint(1, 200)
    g1, g2 = [], []
    score = lambda: round(rng.uniform(2, 120), 2)
    k = 0
    while k < n:
        ln = 3 * rng.randint(30, 200)
        s, e = pos, pos + ln - 1
        st = rng.choice("+-")
        kind = rng.choice(["same", "same", "tie", "start", "start", "end_in", "end_out", "merge", "merge_far", "split", "inner",
                           "only1", "only2", "flip", "both_shift"])
        a = score()
        if kind == "same":
            g1.append((s, e, st, a)); g2.append((s, e, st, score()))
        elif kind == "tie":
            g1.append((s, e, st, a)); g2.append((s, e, st, a))
        elif kind == "start":
            d = 3 * rng.randint(1, 20)
            if rng.random() < 0.5:
                g1.append((s + d, e, st, a)); g2.append((s, e, st, score()))
            else:
                g1.append((s, e, st, a)); g2.append((s + d, e, st, score()))
        elif kind in ("end_in", "end_out"):
            d = 3 * (rng.randint(1, 9) if kind == "end_in" else rng.randint(12, 40))
            if rng.random() < 0.5:
                g1.append((s, e, st, a)); g2.append((s, e + d, st, score()))
            else:
                g1.append((s, e + d, st, a)); g2.append((s, e, st, score()))
            e += d
        elif kind in ("merge", "merge_far", "split", "inner"):
            m = rng.randint(1, 3)
            margin = rng.choice([3, 6, 27, 30]) if kind in ("merge", "inner") else rng.choice([33, 90, 300])
            parts, p = [], s + margin
            for _ in range(m):
                pl = 3 * rng.randint(30, 120)
                parts.append((p, p + pl - 1, st if rng.random() < 0.85 else "+-"[st == "+"], score()))
                p += pl + rng.choice([-3, 0, 12, 60])
            big_e = parts[-1][1] + margin
            tot = round(sum(x[3] for x in parts), 2)
            big = (s, big_e, st, tot if rng.random() < 0.25 else score() * m)
            if kind in ("merge", "merge_far"):
                g1.extend(parts); g2.append(big)
            else:
                g1.append(big); g2.extend(parts)
            e = big_e
        elif kind == "only1":
            g1.append((s, e, st, a))
        elif kind == "only2":
            g2.append((s, e, st, a))
        elif kind == "flip":
            g1.append((s, e, st, a)); g2.append((s, e, "+-"[st == "+"], score()))
        elif kind == "both_shift":
            g1.append((s, e, st, a)); g2.append((s + 3 * rng.randint(1, 5), e + 3 * rng.randint(1, 5), st, score()))
            e += 15
        pos = e + rng.choice([-3, 1, 20, 150])
        k += 1
    g1.sort(key=lambda g: g[0]); g2.sort(key=lambda g: g[0])
    length = max([g[1] for g in g1 + g2] + [pos]) + 50
    seed = rng.randrange(1 << 30)
    return {"seq_seed": seed, "seq_len": length, "seq": case_sequence(seed, length), "genes_1": g1, "genes_2": g2, "tolerance": rng.choice([30, 30, 100]), "consecutive_gene_labels": rng.choice([2, 3, 5]),
            "labels": rng.choice([[0, 2], [0, 1], [1, 3], [2, 3], [0, 3]])}


def main():
    warnings.filterwarnings("ignore")
    MI.install_pandas1_shims()
    Seq = G.install_bio_shim()          # the str subclass of the stand-in
    sys.path.insert(0, f"{MI.REF}/bin")
    import pipeline as refpipe
    from Bio.SeqRecord import SeqRecord
    rng = random.Random(4242)
    cases = []
    errors = 0
    for it in range(300):
        c = make_case(rng)
        if not c["genes_1"] or not c["genes_2"]:
            continue
        P = refpipe.Pipeline("x.fa", "pred/", "fuzz", True, False, 3, c["consecutive_gene_labels"], 5000, 2500, c["tolerance"], False,
                             False, "cwd")
        l1, l2 = c["labels"]
        frames = {lab: frame([], code) for lab, code in enumerate(G.CODES)}
        frames[l1] = frame(c["genes_1"], G.CODES[l1])
        frames[l2] = frame(c["genes_2"], G.CODES[l2])
        P.label_to_gene_coords = frames
        P.sequence = [SeqRecord(Seq(c["seq"]), "c1", "c1")]
        P.updated_genome_labels = [np.array([l1, l1, l2, l2])]
        try:
            coords, g2g, iso = P.compare_sets_of_predicted_genes()
            c["expected"] = {
                "iso": MI.frame_dump(iso, ["source", "start", "end", "strand", "logodd", "attributes", "gene_label"]),
                "coords": MI.frame_dump(coords.reset_index(drop=True), ["start", "end", "strand", "source", "gene_label", "updated_gene_label"]),
            }
        except Exception as ex:      # the reference itself fails on some inputs (label list vs rows, empty tables)
            c["expected"] = {"error": type(ex).__name__}
            errors += 1
        del c["seq"]
        cases.append(c)
    G.xz_write(f"{HERE}/iso_cases.json.xz", json.dumps(cases).encode())
    print(len(cases), "cases,", errors, "where the reference raises")


if __name__ == "__main__":
    main()
