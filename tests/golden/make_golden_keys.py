#!/usr/bin/env python
"""Golden files for the OPTIONAL keys of the .mod grammar, from the reference binary (build container only).

    python tests/golden/make_golden_keys.py

One 30 kb slice of the reference's example genome (examples/phage2.fasta) is decoded by
/root/reference/dependencies/gmhmmp2 (call form of bin/data.py:40-41) under parameter files that differ only in optional
keys.  keys.json.xz holds, per case, the parameter file text, and either the gene lines of the binary's GFF or the error text
it printed.  What the cases pin (DESIGN.md §2 rule 9):

  to_mgm            $TO_MGM in one section only: file-level, TR = ln(to_mgm * to_atypical)
  to_atypical       $TO_ATYPICAL_FIRST_BACTERIA / _SECOND_ARCHAEA: per-domain TR, and the domain choice includes it
  lexicographic     conflicting values in sections 9 / 35 / 100: the section whose LABEL sorts last as a string decides
  ignored           $TO_NATIVE, $BUILD, an unknown key, $RBS 0, $PROMOTER 0, $GENE_GC_DIST 0: no effect
  gc_dist           $GENE_GC_DIST 1 with a 101-line matrix: accepted, no effect on the -M path
  errors            $GENE_GC_DIST 1 without / with a short matrix, negative $TO_MGM: the binary's messages
  site models       $RBS 1 etc.: the binary asks for more keys; this build refuses every non-zero switch
"""
import json
import lzma
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GM = f"{REF}/dependencies/gmhmmp2"

from mgcod_b200 import synth  # noqa: E402


def main():
    tmp = tempfile.mkdtemp()
    seq = synth.read_fasta(f"{REF}/examples/phage2.fasta")[0][2]
    x = synth.encode(seq)
    slice_ = seq[40000:70000]
    fa = os.path.join(tmp, "in.fa")
    open(fa, "w").write(f">slice\n{slice_}\n")

    def model_text(bins, inserts):
        """inserts: {section label or 'all': [lines]} placed in front of $NON_MAT"""
        secs = synth.trained_sections(x, 11, 3, 2, bins=bins, domains=("B", "A"))
        out = []
        for s in secs:
            t = synth.section_text(s).split("\n")
            at = next(i for i, l in enumerate(t) if l.startswith("$NON_MAT"))
            t[at:at] = inserts.get("all", []) + inserts.get(s.label, [])
            out.append("\n".join(t))
        return "".join(out)

    mat = ["$GENE_GC_DIST_MAT"] + [f"{i} {0.9 if i == 30 else 0.001}" for i in range(101)]
    cases = {
        "plain": ([25, 35], {}),
        "to_mgm": ([25, 35], {"__B_GC_35": ["$TO_MGM 0.3"]}),
        "to_atypical": ([25, 35], {"__A_GC_25": ["$TO_ATYPICAL_FIRST_BACTERIA 0.9", "$TO_ATYPICAL_SECOND_ARCHAEA 0.2"]}),
        "lexicographic": ([9, 35, 100], {"__B_GC_9": ["$TO_MGM 0.3", "$TO_ATYPICAL_SECOND_ARCHAEA 0.7"], "__B_GC_35": ["$TO_MGM 0.8"],
                                         "__B_GC_100": ["$TO_MGM 0.6"], "__A_GC_100": ["$TO_ATYPICAL_SECOND_ARCHAEA 0.1"]}),
        "ignored": ([25, 35], {"all": ["$TO_NATIVE 0.9", "$BUILD test", "$SOME_UNKNOWN_KEY 3", "$RBS 0", "$PROMOTER 0", "$SC_RBS 0",
                                        "$SC_PROMOTER 0", "$GENE_GC_DIST 0"]}),
        "gc_dist": ([25, 35], {"all": ["$GENE_GC_DIST 1"] + mat}),
        "err_gc_dist_missing": ([25, 35], {"all": ["$GENE_GC_DIST 1"]}),
        "err_gc_dist_short": ([25, 35], {"all": ["$GENE_GC_DIST 1", "$GENE_GC_DIST_MAT", "30 0.5", "40 0.5"]}),
        "err_to_mgm_negative": ([25, 35], {"all": ["$TO_MGM -1"]}),
        "err_rbs": ([25, 35], {"all": ["$RBS 1"]}),
        "err_promoter": ([25, 35], {"all": ["$PROMOTER 1"]}),
    }
    out = {"sequence": slice_, "cases": {}}
    for name, (bins, ins) in cases.items():
        text = model_text(bins, ins)
        mp, gff = os.path.join(tmp, f"{name}.mod"), os.path.join(tmp, f"{name}.gff")
        open(mp, "w").write(text)
        r = subprocess.run([GM, "-M", mp, "-s", fa, "-f", "gff", "-o", gff], capture_output=True, text=True, errors="replace")
        rec = {"mod": text}
        if r.returncode == 0 and os.path.exists(gff):
            rec["genes"] = [l.rstrip("\n") for l in open(gff) if l.strip() and not l.startswith("#")]
        else:
            rec["error"] = (r.stdout + r.stderr).strip().split("\n")[-1]
        out["cases"][name] = rec
        print(name, len(rec.get("genes", [])), rec.get("error", ""))
    assert out["cases"]["ignored"]["genes"] == out["cases"]["plain"]["genes"] == out["cases"]["gc_dist"]["genes"]
    assert out["cases"]["to_mgm"]["genes"] != out["cases"]["plain"]["genes"]
    with lzma.open(os.path.join(HERE, "keys.json.xz"), "wt", preset=9) as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
