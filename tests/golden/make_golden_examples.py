#!/usr/bin/env python
"""Golden outputs for BASELINE.json configs[2] and configs[3]: the reference's other two example inputs,
examples/GCA_002924505.1_ASM292450v1_genomic.fasta (one bacterial contig) and examples/both.fasta (that contig and
phage2 in one file), through the UNMODIFIED reference Pipeline + Results with gmhmmp2 driven by the committed trained
parameter files (stand-ins for Biopython / pandas 1 as in make_golden_isoforms.py).

  gca.fa.xz                                  the GCA record (both.fasta = phage2.fa.xz + gca.fa.xz, checked below)
  gca_trained_<code>.gff.xz                  gmhmmp2's predictions for it under the four models
  ex_<case>.gff.xz / .tab.xz / _stats.tab.xz Results files; cases: gca (default), both (default), both_iso (--isoforms)
  ex_<case>.json.xz                          labels per file / per window, merged labels, segments

    python tests/golden/make_golden_examples.py
"""
import datetime
import json
import lzma
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_isoforms as MI   # noqa: E402
import make_golden_results as G     # noqa: E402

REF = "/root/reference"


def main():
    warnings.filterwarnings("ignore")
    MI.install_pandas1_shims()
    G.install_bio_shim()
    sys.path.insert(0, f"{REF}/bin")
    import pipeline as refpipe
    import results as refres

    class NP:   # numpy >= 1.24 refuses the ragged np.array(list) at pipeline.py:385
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
    gca = open(f"{REF}/examples/GCA_002924505.1_ASM292450v1_genomic.fasta").read()
    both = open(f"{REF}/examples/both.fasta").read()
    phage = lzma.open(f"{HERE}/phage2.fa.xz", "rt").read()
    G.xz_write(f"{HERE}/gca.fa.xz", gca.encode())
    seqs = lambda t: sorted((l for l in "".join(x if x.startswith(">") else x.strip() for x in ("\n" + y for y in t.split("\n"))).split("\n") if l))
    tmp = tempfile.mkdtemp()
    cwd = os.path.join(tmp, "mgcod")
    os.makedirs(f"{cwd}/dependencies")
    os.symlink(f"{REF}/dependencies/gmhmmp2", f"{cwd}/dependencies/gmhmmp2")
    for c in G.CODES:
        open(f"{cwd}/dependencies/mgm_{c}.mod", "wb").write(lzma.open(f"{HERE}/trained/mgm_{c}.mod.xz").read())
    date = datetime.datetime(2026, 1, 1, 12, 0, 0)
    for case, text, iso in (("gca", gca, False), ("both", both, False), ("both_iso", both, True)):
        fa = os.path.join(tmp, f"{case}.fasta")
        open(fa, "w").write(text)
        pred, outd = os.path.join(tmp, f"pred_{case}") + "/", os.path.join(tmp, f"out_{case}") + "/"
        os.makedirs(pred)
        os.makedirs(outd)
        P = refpipe.Pipeline(fa, pred, case, iso, False, 3, 5, 5000, None, 30, False, False, cwd)
        P()
        dump = {"contigs": [r.id for r in P.sequence], "lengths": [len(r) for r in P.sequence]}
        if iso:
            dump.update(labels=[np.asarray(l).tolist() for l in P.genome_label] if not isinstance(P.genome_label, int) else P.genome_label,
                        updated=[np.asarray(l).tolist() for l in P.updated_genome_labels],
                        scores=[np.round(np.asarray(s, dtype=float), 2).tolist() for s in P.logodd_scores])
            if getattr(P, "confident_1", None) is not None:
                dump.update(confident_1={k: [list(map(int, t)) for t in v] for k, v in P.confident_1.items()},
                            confident_2={k: [list(map(int, t)) for t in v] for k, v in P.confident_2.items()},
                            gene_label_to_gcode={k: {str(i): str(v) for i, v in d.items()} for k, d in P.gene_label_to_gcode.items()})
        else:
            dump.update(file_label=int(P.genome_label),
                        file_sums=[round(float(df["logodd"].sum()), 2) for df in P.label_to_gene_coords.values()])
        if case == "gca":
            for c in G.CODES:
                G.xz_write(f"{HERE}/gca_trained_{c}.gff.xz", G.strip_dates(open(f"{pred}mgm_{c}_{case}.gff").read()).encode())
        R = refres.Results(P, case, pred, fa, outd, date, False, False, False, False, cwd)
        R()
        gff = open(f"{outd}{case}.gff").read().replace(tmp, "TMP")
        G.xz_write(f"{HERE}/ex_{case}.gff.xz", G.strip_dates(gff).encode())
        G.xz_write(f"{HERE}/ex_{case}.tab.xz", open(f"{outd}{case}.tab", "rb").read())
        G.xz_write(f"{HERE}/ex_{case}_stats.tab.xz", open(f"{outd}{case}_stats.tab", "rb").read())
        G.xz_write(f"{HERE}/ex_{case}.json.xz", json.dumps(dump).encode())
        print(case, {k: v for k, v in dump.items() if k in ("contigs", "lengths", "file_label", "file_sums", "gene_label_to_gcode")})
        print(open(f"{outd}{case}.tab").read())
    # both.fasta is the two committed records
    def records(t):
        out = {}
        for block in t.split(">")[1:]:
            head, _, body = block.partition("\n")
            out[head.split()[0]] = body.replace("\n", "")
        return out
    assert records(both) == {**records(phage), **records(gca)}, "both.fasta is not phage2 + GCA"
    print("both.fasta == phage2 + gca; record order in the file:", [l[1:].split()[0] for l in both.split("\n") if l.startswith(">")])


if __name__ == "__main__":
    main()
