#!/usr/bin/env python
"""Golden outputs of the reference in --isoforms mode on inputs WITH a genetic-code switch: the isoform comparison
(Pipeline.compare_sets_of_predicted_genes + helpers, bin/pipeline.py:494-951, 1170-1238) and the switch branches of
Results (bin/results.py:57-75, 101-157, 181-251).

Runs the UNMODIFIED reference classes from /root/reference/bin in the build container.  The reference was written for
pandas 1.x; this image has pandas 3, so the script restores the four pandas-1 behaviours the code relies on before
importing it (nothing in the reference is edited):

  * ``DataFrame.append`` (removed in 2.0): concatenation, a Series counts as one row;
  * assigning a string into a float column (``parent.loc[0, 'logodd'] = '.'``) upcasts to object instead of raising;
  * ``Series.values`` is writable and writes through (copy-on-write made it read-only);
  * text columns are object dtype, not the new ``str`` dtype;
  * and one numpy-1.23 behaviour: ``np.array`` of ragged per-contig lists (bin/pipeline.py:385) gives an object array.

Biopython is absent as well: make_golden_results.install_bio_shim() stands in for it.  Writes next to this file

  iso_<case>.gff.xz / .tab.xz / _stats.tab.xz     files of Results (output-date line stripped, tmp dir -> TMP)
  iso_<case>.json.xz                               window labels, segments, Pipeline.coords / coords_w_isoforms columns
  iso_mixed_mgm_<code>.gff.xz                      gmhmmp2's predictions for the multi-contig case (phage2: phage2_trained_*)

    python tests/golden/make_golden_isoforms.py
"""
import datetime
import json
import lzma
import os
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_results as G   # noqa: E402

REF = "/root/reference"


def install_pandas1_shims():
    def append(self, other, ignore_index=False, verify_integrity=False, sort=False):
        def row(o):
            return o.to_frame().T if isinstance(o, pd.Series) else o
        objs = [self] + ([row(o) for o in other] if isinstance(other, (list, tuple)) else [row(other)])
        return pd.concat(objs, ignore_index=ignore_index, sort=sort)

    pd.DataFrame.append = append
    pd.options.future.infer_string = False
    from pandas.core.internals import blocks
    orig = blocks.Block.coerce_to_target_dtype
    blocks.Block.coerce_to_target_dtype = lambda self, other, raise_on_upcast=False: orig(self, other, raise_on_upcast=False)
    values = pd.Series.values

    def writable_values(self):
        a = values.fget(self)
        if isinstance(a, np.ndarray) and not a.flags.writeable:
            try:
                a.flags.writeable = True
            except ValueError:
                a = a.copy()
        return a

    pd.Series.values = property(writable_values)


def cases():
    phage = lzma.open(f"{HERE}/phage2.fa.xz", "rt").read()
    seq = "".join(l.strip() for l in phage.split("\n") if not l.startswith(">"))
    # phage2 switches 15 -> 11 -> 15 at 104547 and 133177
    mixed = (f">zeta two codes\n{seq[80000:125000]}\n>alpha\n{seq[106000:131000]}\n>mid x\n{seq[0:50000]}\n"
             f">beta back\n{seq[120000:173000]}\n")
    return {
        "phage2": (phage, dict(window=5000, stride=2500, n=3, genes=5, tol=30)),
        "mixed": (mixed, dict(window=5000, stride=2500, n=3, genes=5, tol=30)),
        "phage2_t300": (phage, dict(window=5000, stride=5000, n=2, genes=3, tol=300)),
        # two codes in the input but no contig with a switch: the isoform table stays empty
        "two_single": (f">a15\n{seq[0:50000]}\n>b11\n{seq[106000:131000]}\n", dict(window=5000, stride=2500, n=3, genes=5, tol=30)),
    }


def frame_dump(df, cols):
    out = {}
    for c in cols:
        if c in df.columns:
            out[c] = [None if (isinstance(v, float) and v != v) else (v if isinstance(v, str) else float(v) if c == "logodd" else int(v))
                      for v in df[c].values]
    return out


def main():
    warnings.filterwarnings("ignore")
    install_pandas1_shims()
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
    tmp = tempfile.mkdtemp()
    cwd = os.path.join(tmp, "mgcod")
    os.makedirs(f"{cwd}/dependencies")
    os.symlink(f"{REF}/dependencies/gmhmmp2", f"{cwd}/dependencies/gmhmmp2")
    for c in G.CODES:
        open(f"{cwd}/dependencies/mgm_{c}.mod", "wb").write(lzma.open(f"{HERE}/trained/mgm_{c}.mod.xz").read())
    date = datetime.datetime(2026, 1, 1, 12, 0, 0)
    for case, (text, a) in cases().items():
        fa = os.path.join(tmp, f"{case}.fasta")
        open(fa, "w").write(text)
        pred, outd = os.path.join(tmp, f"pred_{case}") + "/", os.path.join(tmp, f"out_{case}") + "/"
        os.makedirs(pred)
        os.makedirs(outd)
        P = refpipe.Pipeline(fa, pred, case, True, False, a["n"], a["genes"], a["window"], a["stride"], a["tol"], False, False, cwd)
        P()
        dump = {
            "args": a,
            "contigs": [r.id for r in P.sequence],
            "labels": [np.asarray(l).tolist() for l in P.genome_label],
            "updated": [np.asarray(l).tolist() for l in P.updated_genome_labels],
            "confident_1": {k: [list(map(int, t)) for t in v] for k, v in P.confident_1.items()},
            "confident_2": {k: [list(map(int, t)) for t in v] for k, v in P.confident_2.items()},
            "gene_label_to_gcode": {k: {str(i): str(v) for i, v in d.items()} for k, d in P.gene_label_to_gcode.items()},
            "coords": frame_dump(P.coords.reset_index(drop=True), ["contig", "start", "end", "strand", "logodd", "source", "gene_label",
                                                                   "updated_gene_label"]),
            "coords_w_isoforms": frame_dump(P.coords_w_isoforms, ["contig", "source", "start", "end", "strand", "logodd", "attributes",
                                                                  "gene_label"]),
        }
        if case in ("mixed", "two_single"):    # the binary's predictions, so that the comparison can be tested without a GPU
            for c in G.CODES:
                G.xz_write(f"{HERE}/iso_{case}_mgm_{c}.gff.xz", G.strip_dates(open(f"{pred}mgm_{c}_{case}.gff").read().replace(tmp, "TMP")).encode())
        R = refres.Results(P, case, pred, fa, outd, date, False, False, True, False, cwd)      # -NT
        R()
        import hashlib
        dump["nt_sha256"] = hashlib.sha256(open(f"{outd}proteins_nt_{case}.fasta", "rb").read()).hexdigest()
        gff = open(f"{outd}{case}.gff").read().replace(tmp, "TMP")
        G.xz_write(f"{HERE}/iso_{case}.gff.xz", G.strip_dates(gff).encode())
        G.xz_write(f"{HERE}/iso_{case}.tab.xz", open(f"{outd}{case}.tab", "rb").read())
        G.xz_write(f"{HERE}/iso_{case}_stats.tab.xz", open(f"{outd}{case}_stats.tab", "rb").read())
        G.xz_write(f"{HERE}/iso_{case}.json.xz", json.dumps(dump).encode())
        iso = P.coords_w_isoforms
        print(case, "rows", iso.shape[0], "parents", int((iso.logodd == ".").sum()), P.gene_label_to_gcode)
        print(open(f"{outd}{case}.tab").read())


if __name__ == "__main__":
    main()
