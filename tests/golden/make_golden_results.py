#!/usr/bin/env python
"""Golden outputs of the reference's Results class (bin/results.py) in default mode.

Runs the UNMODIFIED reference Pipeline + Results (bin/pipeline.py, bin/data.py, bin/results.py under /root/reference)
in the build container, with gmhmmp2 driven by the committed trained parameter files and a small stand-in for
Biopython (absent here).  Writes next to this file:

  results_<case>.gff.xz / .tab.xz / _stats.tab.xz    the three files Results writes (dates stripped)
  results_<case>_nt.fasta.xz                          -NT output (gene sequences, recoded stops upper-cased)
  recoded_ref.json                                    Results.find_recoded_stops on seeded random sequences

    python tests/golden/make_golden_results.py
"""
import datetime
import json
import lzma
import os
import random
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference"
CODES = (11, 4, 15, 101)


def install_bio_shim():
    comp = str.maketrans("ACGTacgtNn", "TGCAtgcaNn")

    class Seq(str):
        def __getitem__(self, k):
            r = str.__getitem__(self, k)
            return Seq(r) if isinstance(k, slice) else r

        def reverse_complement(self):
            return Seq(str(self)[::-1].translate(comp))

        def lower(self):
            return Seq(str.lower(self))

        def upper(self):
            return Seq(str.upper(self))

        def __add__(self, o):
            return Seq(str(self) + str(o))

    class SeqRecord:
        def __init__(self, seq, id="", description=""):
            self.seq, self.id, self.description = seq, id, description

        def __len__(self):
            return len(self.seq)

    def parse(path, fmt):
        rid = desc = None
        chunks = []
        for line in open(path):
            line = line.rstrip("\n")
            if line.startswith(">"):
                if rid is not None:
                    yield SeqRecord(Seq("".join(chunks)), rid, desc)
                desc = line[1:]
                rid = desc.split()[0] if desc.split() else ""
                chunks = []
            elif rid is not None:
                chunks.append(line.strip())
        if rid is not None:
            yield SeqRecord(Seq("".join(chunks)), rid, desc)

    def write(records, path, fmt):
        with open(path, "w") as fh:
            for r in records:
                # Biopython: ">id description" unless the description already starts with the id
                head = r.description if r.description.startswith(r.id + " ") or r.description == r.id else f"{r.id} {r.description}"
                fh.write(f">{head}\n")
                s = str(r.seq)
                for i in range(0, len(s), 60):
                    fh.write(s[i:i + 60] + "\n")

    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqio.parse, seqio.write = parse, write
    su = types.ModuleType("Bio.SeqUtils")
    su.GC = lambda s: 100.0 * sum(c in "GCgc" for c in str(s)) / max(1, len(s))
    sr = types.ModuleType("Bio.SeqRecord")
    sr.SeqRecord = SeqRecord
    data = types.ModuleType("Bio.Data")
    ct = types.ModuleType("Bio.Data.CodonTable")
    ct.CodonTable = object
    data.CodonTable = ct
    bio.SeqIO, bio.SeqUtils, bio.SeqRecord, bio.Data = seqio, su, sr, data
    sys.modules.update({"Bio": bio, "Bio.SeqIO": seqio, "Bio.SeqUtils": su, "Bio.SeqRecord": sr, "Bio.Data": data,
                        "Bio.Data.CodonTable": ct})
    return Seq


def xz_write(path, data: bytes):
    with lzma.open(path, "wb", preset=9) as fh:
        fh.write(data)


def strip_dates(text):
    return "\n".join(l for l in text.split("\n") if not l.startswith("# output date"))


def main():
    Seq = install_bio_shim()
    sys.path.insert(0, f"{REF}/bin")
    import pipeline as refpipe
    import results as refres

    # ---- find_recoded_stops on random sequences -----------------------------------------------------------
    rng = random.Random(77)
    cases = []
    for it in range(200):
        n = rng.choice([3, 6, 9, 30, 31, 32, 60, 150, 299, 300])
        s = "".join(rng.choice("acgt" if it % 3 else "tag") for _ in range(n))
        out = {}
        work = Seq(s)
        for stop in ("tag", "tga", "taa"):
            pos, work = refres.Results.find_recoded_stops(work, stop)
            out[stop.upper()] = pos
        cases.append({"seq": s, "recodings": out, "marked": str(work)})
    json.dump(cases, open(f"{HERE}/recoded_ref.json", "w"))

    # ---- full default-mode runs ---------------------------------------------------------------------------
    tmp = tempfile.mkdtemp()
    cwd = os.path.join(tmp, "mgcod")
    os.makedirs(f"{cwd}/dependencies")
    os.symlink(f"{REF}/dependencies/gmhmmp2", f"{cwd}/dependencies/gmhmmp2")
    for c in CODES:
        open(f"{cwd}/dependencies/mgm_{c}.mod", "wb").write(lzma.open(f"{HERE}/trained/mgm_{c}.mod.xz").read())
    phage = lzma.open(f"{HERE}/phage2.fa.xz", "rt").read()
    seq = "".join(l.strip() for l in phage.split("\n") if not l.startswith(">"))
    inputs = {
        "phage2": phage,
        "three_contigs": f">zeta part\n{seq[1000:31000]}\n>alpha\n{seq[60000:100000]}\n>mid x y\n{seq[120000:150000]}\n",
    }
    date = datetime.datetime(2026, 1, 1, 12, 0, 0)
    for case, text in inputs.items():
        fa = os.path.join(tmp, f"{case}.fasta")
        open(fa, "w").write(text)
        pred, outd = os.path.join(tmp, f"pred_{case}") + "/", os.path.join(tmp, f"out_{case}") + "/"
        os.makedirs(pred)
        os.makedirs(outd)
        P = refpipe.Pipeline(fa, pred, case, False, False, 3, 5, 5000, 5000, 30, False, False, cwd)
        P()
        R = refres.Results(P, case, pred, fa, outd, date, False, False, True, False, cwd)
        R()
        gff = open(f"{outd}{case}.gff").read().replace(tmp, "TMP")
        xz_write(f"{HERE}/results_{case}.gff.xz", strip_dates(gff).encode())
        xz_write(f"{HERE}/results_{case}.tab.xz", open(f"{outd}{case}.tab", "rb").read())
        xz_write(f"{HERE}/results_{case}_stats.tab.xz", open(f"{outd}{case}_stats.tab", "rb").read())
        if case != "phage2":   # 45 KB; the three-contig file covers the format
            xz_write(f"{HERE}/results_{case}_nt.fasta.xz", open(f"{outd}proteins_nt_{case}.fasta", "rb").read())
        print(case, "genome_label", P.genome_label, "genes", P.coords_w_isoforms.shape[0])


if __name__ == "__main__":
    main()
