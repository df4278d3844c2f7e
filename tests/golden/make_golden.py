#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory by running the REFERENCE in the build container.

Needs /root/reference (read-only): the closed binary dependencies/gmhmmp2 and bin/pipeline.py.  Nothing here
runs on the GPU box; the fixtures it writes are committed and are what tests/ compare against.

    python tests/golden/make_golden.py

Fixtures
  phage2.fa.xz                    the reference's example input (examples/phage2.fasta, BASELINE.json configs 0/1)
  trained/mgm_<code>.mod.xz       order-4/2 single-section parameter files trained on phage2
  phage2_trained_<code>.gff.xz    gmhmmp2 -f gff output   (and .lst.xz for code 15)
  synth_parametric.json           sha256 of the regenerated synthetic contigs + parametric 82-section models
  synth_parametric_<code>.gff.xz  gmhmmp2 output on those
  odd_<name>.*                    small multi-record / edge inputs and gmhmmp2's GFF for them
  labels.json                     reference Pipeline.find_best_sequence_of_labels on seeded random label arrays
  pipeline_phage2.json            reference identify_most_common_codes / find_best / segment outputs for phage2
                                  (--isoforms -w 5000 -st 2500 -n 3) and for a two-contig input
"""
import hashlib
import json
import lzma
import os
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GM = f"{REF}/dependencies/gmhmmp2"
CODES = (11, 4, 15, 101)

from mgcod_b200 import synth  # noqa: E402
from mgcod_b200.engine import synth_metagenome  # noqa: E402


def xz_write(path, data: bytes):
    with lzma.open(path, "wb", preset=9) as fh:
        fh.write(data)


def strip_dates(text: str) -> str:
    return "\n".join(l for l in text.split("\n") if not l.startswith("# output date")) + ""


def run_gm(mod, fasta, fmt="gff"):
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o." + fmt)
        subprocess.run([GM, "-M", mod, "-s", fasta, "-f", fmt, "-o", out], check=True, capture_output=True)
        return open(out).read()


def reference_pipeline_module():
    """Import the reference's bin/pipeline.py unmodified, with stand-ins for the packages absent here."""
    bio = types.ModuleType("Bio")
    su = types.ModuleType("Bio.SeqUtils")
    su.GC = lambda s: 100.0 * sum(c in "GCgc" for c in str(s)) / max(1, len(s))
    bio.SeqUtils = su
    bio.SeqIO = types.ModuleType("Bio.SeqIO")
    sys.modules.update({"Bio": bio, "Bio.SeqUtils": su, "Bio.SeqIO": bio.SeqIO})
    sys.path.insert(0, f"{REF}/bin")
    import pipeline as refpipe
    return refpipe


class Rec:
    def __init__(self, rid, seq):
        self.id, self.seq = rid, seq

    def __len__(self):
        return len(self.seq)


def ref_pipeline_outputs(refpipe, gff_texts, records, window, stride, n, short_contigs):
    import io
    import pandas as pd
    P = refpipe.Pipeline("x.fa", "pred/", "x", True, False, n, 5, window, stride, 30, short_contigs, False, "/tmp")
    P.sequence = sorted(records, key=lambda r: r.id) if len(records) > 1 else records
    P.nr_contigs = len(records)
    dfs = []
    for code in CODES:
        df = pd.read_csv(io.StringIO(gff_texts[code]), sep="\t", header=None, comment="#",
                         names=["sequence", "source", "feature", "start", "end", "logodd", "strand", "phase", "attributes"])
        df["source"] = df["source"] + f"_{code}"
        df["length"] = df["end"] - df["start"] + 1
        df["sequence"] = df["sequence"].astype(str)
        df["contig"] = df["sequence"].values
        dfs.append(df)
    P.gene_coords_11, P.gene_coords_4, P.gene_coords_15, P.gene_coords_101 = dfs
    P.label_to_gene_coords = dict(enumerate(dfs))
    # numpy >= 1.24 refuses the ragged np.array(list) at pipeline.py:385; give the module a tolerant np.array
    class NP:
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def array(x, *a, **kw):
            try:
                return np.array(x, *a, **kw)
            except ValueError:
                o = np.empty(len(x), dtype=object)
                for i, v in enumerate(x):
                    o[i] = v
                return o
    refpipe.np = NP()
    labels, scores = P.identify_most_common_codes()
    refpipe.np = np
    labels = [np.asarray(l) for l in labels]
    scores = [np.asarray(s, dtype=float) for s in scores]
    allab = np.concatenate(labels)
    updated = []
    for l in labels:
        try:
            updated.append(l if np.unique(l).shape[0] == 1 else P.find_best_sequence_of_labels(l, allab, n))
        except IndexError:          # the reference walks off the array when a contig starts with a noise label
            updated.append(np.array([-1]))
    out = {"contigs": [r.id for r in P.sequence], "labels": [l.tolist() for l in labels],
           "scores": [np.round(s, 2).tolist() for s in scores], "updated": [np.asarray(u).tolist() for u in updated]}
    if np.unique(np.concatenate(updated)).shape[0] > 1 and not any((np.asarray(u) < 0).any() for u in updated):
        P.updated_genome_labels = updated
        c1, c2, sl = P.segment_genome_based_on_window_labels()
        out["confident_1"] = {k: [list(map(int, t)) for t in v] for k, v in c1.items()}
        out["confident_2"] = {k: [list(map(int, t)) for t in v] for k, v in c2.items()}
        out["segment_labels"] = {k: [int(x) for x in v] for k, v in sl.items()}
    sums = [float(df["logodd"].sum()) for df in dfs]
    out["file_sums"] = [round(s, 2) for s in sums]
    out["file_label"] = CODES[int(np.argmax(sums))]
    return out


def main():
    tmp = tempfile.mkdtemp()
    # ---- phage2 + trained models ---------------------------------------------------------------------
    name, desc, seq = synth.read_fasta(f"{REF}/examples/phage2.fasta")[0]
    fa_text = f">{desc}\n{seq}\n"
    xz_write(f"{HERE}/phage2.fa.xz", fa_text.encode())
    phage_fa = os.path.join(tmp, "phage2.fasta")
    open(phage_fa, "w").write(fa_text)
    x = synth.encode(seq)
    os.makedirs(f"{HERE}/trained", exist_ok=True)
    gffs = {}
    for c in CODES:
        mp = os.path.join(tmp, f"mgm_{c}.mod")
        synth.write_mod(mp, synth.trained_sections(x, c, 4, 2))
        xz_write(f"{HERE}/trained/mgm_{c}.mod.xz", open(mp, "rb").read())
        gffs[c] = run_gm(mp, phage_fa)
        xz_write(f"{HERE}/phage2_trained_{c}.gff.xz", strip_dates(gffs[c]).encode())
    xz_write(f"{HERE}/phage2_trained_15.lst.xz", strip_dates(run_gm(os.path.join(tmp, "mgm_15.mod"), phage_fa, "lst")).encode())

    # ---- synthetic contigs + parametric 82-section models --------------------------------------------
    buf, offs, truth = synth_metagenome(424242, 150000, threads=2)
    pdir = os.path.join(tmp, "par")
    paths = synth.write_model_set(pdir, kind="parametric")
    sfa = os.path.join(tmp, "synth.fa")
    with open(sfa, "w") as fh:
        for i in range(len(truth)):
            fh.write(f">contig_{i} code={truth[i]}\n{buf[offs[i]:offs[i + 1]].tobytes().decode()}\n")
    meta = {"seed": 424242, "total_bp": 150000, "contigs": int(len(truth)), "truth": truth.tolist(),
            "seq_sha256": hashlib.sha256(buf.tobytes()).hexdigest(),
            "model_sha256": {str(c): hashlib.sha256(open(paths[c], "rb").read()).hexdigest() for c in CODES}}
    json.dump(meta, open(f"{HERE}/synth_parametric.json", "w"), indent=1)
    sgff = {}
    for c in CODES:
        sgff[c] = run_gm(paths[c], sfa)
        xz_write(f"{HERE}/synth_parametric_{c}.gff.xz", strip_dates(sgff[c]).encode())

    # ---- odd inputs: multi-record with empty / tiny / lower-case records, small multi-section model ----
    odd_mod = os.path.join(tmp, "odd.mod")
    synth.write_mod(odd_mod, synth.trained_sections(x, 4, 3, 1, bins=[22, 31, 40], domains=("B", "A")))
    xz_write(f"{HERE}/odd_model_code4.mod.xz", open(odd_mod, "rb").read())
    odd_fa = os.path.join(tmp, "odd.fa")
    rc = seq[20000:26000][::-1].translate(str.maketrans("ACGT", "TGCA"))
    with open(odd_fa, "w") as fh:
        fh.write(">tiny one\nACGTACGTACGTAAATTT\n>slice|a b\n" + "\n".join(seq[63400:65300][i:i + 70] for i in range(0, 1900, 70)) + "\n")
        fh.write(">empty\n>lower\n" + seq[100000:104000].lower() + "\n>two\nAC\n>rc\n" + rc + "\n>nostop\n" + "ATGGCA" * 40 + "\n")
    xz_write(f"{HERE}/odd_input.fa.xz", open(odd_fa, "rb").read())
    xz_write(f"{HERE}/odd_code4.gff.xz", strip_dates(run_gm(odd_mod, odd_fa)).encode())
    xz_write(f"{HERE}/odd_code4.lst.xz", strip_dates(run_gm(odd_mod, odd_fa, "lst")).encode())

    # ---- label smoothing vectors ----------------------------------------------------------------------
    refpipe = reference_pipeline_module()
    rng = np.random.default_rng(20260101)
    cases = []
    for it in range(400):
        n = int(rng.integers(2, 90))
        k = int(rng.integers(2, 6))
        nlab = int(rng.integers(2, 5))
        p = rng.dirichlet(np.ones(nlab) * 0.6)
        # piecewise-constant with noise, like real window labels
        lab = []
        while len(lab) < n:
            lab += [int(rng.choice(nlab, p=p))] * int(rng.integers(1, 15))
        lab = np.array(lab[:n])
        flip = rng.random(n) < 0.12
        lab[flip] = rng.integers(0, nlab, flip.sum())
        if np.unique(lab).shape[0] == 1:
            continue
        multi = bool(it % 7 == 0)
        got = refpipe.Pipeline.find_best_sequence_of_labels(lab.copy(), lab.copy(), k, multi)
        cases.append({"labels": lab.tolist(), "n": k, "allow_multiple": multi, "expected": np.asarray(got).tolist()})
    json.dump(cases, open(f"{HERE}/labels.json", "w"))

    # ---- reference pipeline outputs -------------------------------------------------------------------
    out = {"phage2_w5000_st2500_n3": ref_pipeline_outputs(refpipe, gffs, [Rec("genome", seq)], 5000, 2500, 3, False)}
    recs = [Rec(f"contig_{i}", buf[offs[i]:offs[i + 1]].tobytes().decode()) for i in range(len(truth))]
    out["synth_default_stride_n3"] = ref_pipeline_outputs(refpipe, sgff, recs, 5000, None, 3, False)
    out["synth_short_contigs_w4000_st2000_n2"] = ref_pipeline_outputs(refpipe, sgff, recs, 4000, 2000, 2, True)
    json.dump(out, open(f"{HERE}/pipeline_ref.json", "w"))
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f"  {f:40s} {os.path.getsize(os.path.join(HERE, f)):8d}")


if __name__ == "__main__":
    main()
