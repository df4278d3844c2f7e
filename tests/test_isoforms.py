"""Isoform comparison (mgcod_b200.isoforms) against the reference's own compare_sets_of_predicted_genes on the binary's
predictions: tests/golden/iso_*.json.xz (make_golden_isoforms.py).  Host logic only — no GPU."""
import io
import json
import os
import types

import numpy as np
import pandas as pd
import pytest

from tests import fixtures as fx
from mgcod_b200 import isoforms
from mgcod_b200.data import GFF_COLUMNS, Record

CASES = {"phage2": ("phage2.fa.xz", "phage2_trained_{}.gff.xz"),
         "phage2_t300": ("phage2.fa.xz", "phage2_trained_{}.gff.xz"),
         "mixed": (None, "iso_mixed_mgm_{}.gff.xz"),
         "two_single": (None, "iso_two_single_mgm_{}.gff.xz")}      # two codes in the input, no contig with a switch


def mixed_fasta():
    seq = fx.phage2()[0][1]
    return [("zeta", "zeta two codes", seq[80000:125000]), ("alpha", "alpha", seq[106000:131000]),
            ("mid", "mid x", seq[0:50000]), ("beta", "beta back", seq[120000:173000])]


def two_single_fasta():
    seq = fx.phage2()[0][1]
    return [("a15", "a15", seq[0:50000]), ("b11", "b11", seq[106000:131000])]


def records_of(case):
    if case == "mixed":
        recs = [Record(i, d, s) for i, d, s in mixed_fasta()]
    elif case == "two_single":
        recs = [Record(i, d, s) for i, d, s in two_single_fasta()]
    else:
        recs = [Record(i, i, s) for i, s in fx.phage2()]
    return sorted(recs, key=lambda r: r.id)


def fake_pipeline(case):
    """what Pipeline holds when it calls the comparison, with gmhmmp2's own GFFs as the gene sets"""
    ref = json.loads(fx.xz_text(f"iso_{case}.json.xz"))
    frames = {}
    for lab, code in enumerate(fx.CODES):
        df = pd.read_csv(io.StringIO(fx.xz_text(CASES[case][1].format(code))), sep="\t", header=None, comment="#", names=GFF_COLUMNS)
        df["source"] = df["source"] + f"_{code}"
        df["length"] = df["end"] - df["start"] + 1
        df["sequence"] = df["sequence"].astype(str)
        df["contig"] = list(df.sequence.values)
        frames[lab] = df
    a = ref["args"]
    P = types.SimpleNamespace(species=case, verbose=False, label_to_gene_coords=frames, sequence=records_of(case),
                              updated_genome_labels=[np.array(u) for u in ref["updated"]], tolerance=a["tol"],
                              consecutive_gene_labels=a["genes"], label_to_gcode=dict(enumerate(fx.CODES)))
    assert [r.id for r in P.sequence] == ref["contigs"]
    return P, ref


def column(df, c):
    out = []
    for v in df[c].values:
        if isinstance(v, float) and v != v:
            out.append(None)
        elif isinstance(v, str):
            out.append(v)
        else:
            out.append(float(v) if c == "logodd" else int(v))
    return out


@pytest.mark.parametrize("case", list(CASES))
def test_compare_sets_matches_reference(case):
    P, ref = fake_pipeline(case)
    coords, g2g, iso = isoforms.compare_sets_of_predicted_genes(P)
    assert {k: {str(i): str(v) for i, v in d.items()} for k, d in g2g.items()} == ref["gene_label_to_gcode"]
    want = ref["coords_w_isoforms"]
    assert iso.shape[0] == len(want["start"])
    for c in ("contig", "source", "start", "end", "strand", "logodd", "attributes"):
        assert column(iso, c) == want[c], c
    # gene_label: the reference leaves it unset (NaN) on some parents; where it is set it must agree
    got = column(iso, "gene_label")
    want_labels = want.get("gene_label", [])        # the reference's empty table has no such column
    assert [g for g, w in zip(got, want_labels) if w is not None] == [w for w in want_labels if w is not None]
    # the plain table: same genes in the same order (ties of the sort by start included), labels and smoothed labels.  For a
    # contig with one code the reference appends that model's genes of every contig (bin/pipeline.py:574-576); the mirror
    # keeps the contig's own, so those blocks are compared after dropping the foreign rows.
    coords = coords.reset_index(drop=True)
    wc = ref["coords"]
    single = {c for c, d in ref["gene_label_to_gcode"].items() if len(d) == 1}
    keep, i = [], 0
    for c in ref["contigs"]:                         # the reference's table is one block per contig, in this order
        if c in single:                              # block = the whole table of that model
            frame = P.label_to_gene_coords[fx.CODES.index(int(ref["gene_label_to_gcode"][c]["1"]))]
            keep += [k for k in range(i, i + len(frame)) if wc["contig"][k] == c]
            i += len(frame)
        else:
            while i < len(wc["start"]) and wc["gene_label"][i] is not None and wc["contig"][i] == c:
                keep.append(i)
                i += 1
    assert i == len(wc["start"])
    for c in ("contig", "start", "end", "strand", "source", "gene_label", "updated_gene_label"):
        if c in wc:          # no gene_label column when no contig went through the comparison
            assert column(coords, c) == [wc[c][k] for k in keep], c


def test_update_gene_ids_and_insert():
    rows = [{"start": 10, "attributes": "gene_id 1; x"}, {"start": 50, "attributes": "gene_id 2; gene_type a;"},
            {"start": 40, "attributes": "gene_id 2.11; Parent=2; y"}, {"start": 50, "attributes": "gene_id 2.15; Parent=2; z"},
            {"start": 90, "attributes": "gene_id 3; w"}]
    out = isoforms.insert_unseen_genes([dict(r) for r in rows], [{"start": 45, "attributes": "gene_id 77; u"}])
    # last row starting at or before 45 is the child at 40 (index 2); its sibling is skipped; the gene takes id 3
    assert [r["attributes"] for r in out] == ["gene_id 1; x", "gene_id 2; gene_type a;", "gene_id 2.11; Parent=2; y",
                                              "gene_id 2.15; Parent=2; z", "gene_id 3; u", "gene_id 4; w"]
    out = isoforms.insert_unseen_genes([dict(r) for r in rows], [{"start": 5, "attributes": "gene_id 9; u"}])
    assert [r["attributes"].split(";")[0] for r in out] == ["gene_id 0", "gene_id 2", "gene_id 3", "gene_id 3.11", "gene_id 3.15",
                                                            "gene_id 4"]
    out = isoforms.insert_unseen_genes([dict(r) for r in rows], [{"start": 95, "attributes": "gene_id 9; u"}])
    assert out[-1]["attributes"] == "gene_id 4; u"


def test_gc_percent():
    assert isoforms.gc_percent("") == 0.0
    assert isoforms.gc_percent("GCgcSsATat") == 60.0


class HostResults:
    """Results with the in-frame stop scan done by the host restatement instead of the GPU kernel (test-only: the product
    class always uses the engine)."""

    @staticmethod
    def make(P, case, outdir):
        import datetime
        from mgcod_b200.results import Results, _COMP

        class R(Results):
            def extract_sequences(self):
                out = []
                coords = self.results.coords_w_isoforms
                for n in range(self.results.nr_contigs):
                    rec = self.results.sequence[n]
                    cur = coords[coords.contig == rec.id]
                    cur = cur[cur.logodd != "."]
                    for s, e, st in zip(cur.start.values, cur.end.values, cur.strand.values):
                        seq = rec.seq[s - 1:e]
                        if st == "-":
                            seq = seq[::-1].translate(_COMP)
                        seq = seq.lower()
                        d = {}
                        for name in ("TAG", "TGA", "TAA"):
                            d[name], seq = Results.find_recoded_stops(seq, name.lower())
                        out.append(d)
                return out

        return R(P, case, "PRED/", f"TMP/{case}.fasta", outdir, datetime.datetime(2026, 1, 1, 12, 0, 0), False, False, False, False,
                 "TMP/mgcod")


@pytest.mark.parametrize("case", list(CASES))
def test_results_files_of_a_switch_run(case, tmp_path):
    """.gff / .tab / _stats.tab of a run with a switch, byte for byte (the output-date line aside)"""
    P, ref = fake_pipeline(case)
    P.coords, P.gene_label_to_gcode, P.coords_w_isoforms = isoforms.compare_sets_of_predicted_genes(P)
    P.nr_contigs = len(P.sequence)
    P.genome_label = [np.array(l) for l in ref["labels"]]
    P.confident_1 = {k: [tuple(t) for t in v] for k, v in ref["confident_1"].items()}
    P.confident_2 = {k: [tuple(t) for t in v] for k, v in ref["confident_2"].items()}
    out = str(tmp_path) + "/"
    R = HostResults.make(P, case, out)
    R()
    strip = lambda t: "\n".join(l for l in t.split("\n") if not l.startswith("# output date"))
    assert open(f"{out}{case}.tab").read() == fx.xz_text(f"iso_{case}.tab.xz")
    assert open(f"{out}{case}_stats.tab").read() == fx.xz_text(f"iso_{case}_stats.tab.xz")
    assert strip(open(f"{out}{case}.gff").read()) == fx.xz_text(f"iso_{case}.gff.xz")


def _case_text(case):
    if case == "mixed":
        return "".join(f">{d}\n{s}\n" for _, d, s in mixed_fasta())
    if case == "two_single":
        return "".join(f">{d}\n{s}\n" for _, d, s in two_single_fasta())
    return fx.xz_text("phage2.fa.xz")


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("via_cli", [False, True])
def test_switch_run_end_to_end_on_the_gpu(case, via_cli, tmp_path):
    """FASTA -> CUDA decode under four models -> window labels -> isoform comparison -> Results files, equal to the
    reference's files for the same command line (--isoforms -w .. -st .. -n .. -g .. -t ..)"""
    import datetime
    from mgcod_b200.pipeline import Pipeline
    from mgcod_b200.results import Results
    tmp = str(tmp_path)
    ref = json.loads(fx.xz_text(f"iso_{case}.json.xz"))
    a = ref["args"]
    mdir = os.path.dirname(fx.trained_models(tmp)[11])
    fa = os.path.join(tmp, f"{case}.fasta")
    open(fa, "w").write(_case_text(case))
    pred, out = os.path.join(tmp, "pred") + "/", os.path.join(tmp, "out") + "/"
    if via_cli:
        from mgcod_b200 import cli
        assert cli.main(["-i", fa, "-p", pred, "-o", out, "--isoforms", "-w", str(a["window"]), "-st", str(a["stride"]), "-n", str(a["n"]),
                         "-g", str(a["genes"]), "-t", str(a["tol"]), "--model_dir", mdir]) == 0
    else:
        os.makedirs(pred)
        os.makedirs(out)
        P = Pipeline(fa, pred, case, True, False, a["n"], a["genes"], a["window"], a["stride"], a["tol"], False, False, "CWD",
                     model_dir=mdir)
        P()
        assert [u.tolist() for u in P.updated_genome_labels] == ref["updated"]
        assert {k: {str(i): str(v) for i, v in d.items()} for k, d in P.gene_label_to_gcode.items()} == ref["gene_label_to_gcode"]
        Results(P, case, pred, fa, out, datetime.datetime(2026, 1, 1, 12, 0, 0), False, False, False, False, "CWD")()
    golden = fx.xz_text(f"iso_{case}.gff.xz").split("\n")
    mine = fx.strip_dates(open(f"{out}{case}.gff").read()).split("\n")
    assert len(golden) == len(mine)
    for g, m in zip(golden, mine):
        if g.startswith("# File with"):     # paths differ (temp dirs); the rest of the header must agree
            assert m.split(":")[0] == g.split(":")[0]
            continue
        assert g == m
    assert open(f"{out}{case}.tab").read() == fx.xz_text(f"iso_{case}.tab.xz")
    assert open(f"{out}{case}_stats.tab").read() == fx.xz_text(f"iso_{case}_stats.tab.xz")


@pytest.mark.gpu
def test_isoforms_run_on_cached_predictions(tmp_path):
    """`cli -i g.fa` followed by `cli -i g.fa --isoforms` with the same -p directory: the second run finds all four
    mgm_<code>_<species>.gff files cached (bin/data.py:39-51 skips gmhmmp2 then), decodes again for the window labels
    instead of failing, leaves the cached files alone and writes the reference's files"""
    from mgcod_b200 import cli
    tmp = str(tmp_path)
    case = "mixed"
    ref = json.loads(fx.xz_text(f"iso_{case}.json.xz"))
    a = ref["args"]
    mdir = os.path.dirname(fx.trained_models(tmp)[11])
    fa = os.path.join(tmp, f"{case}.fasta")
    open(fa, "w").write(_case_text(case))
    pred, out1, out2 = os.path.join(tmp, "pred") + "/", os.path.join(tmp, "out1") + "/", os.path.join(tmp, "out2") + "/"
    assert cli.main(["-i", fa, "-p", pred, "-o", out1, "--model_dir", mdir]) == 0
    cached = {c: os.stat(f"{pred}mgm_{c}_{case}.gff").st_mtime_ns for c in fx.CODES}
    assert cli.main(["-i", fa, "-p", pred, "-o", out2, "--isoforms", "-w", str(a["window"]), "-st", str(a["stride"]), "-n", str(a["n"]),
                     "-g", str(a["genes"]), "-t", str(a["tol"]), "--model_dir", mdir]) == 0
    assert cached == {c: os.stat(f"{pred}mgm_{c}_{case}.gff").st_mtime_ns for c in fx.CODES}
    assert open(f"{out2}{case}.tab").read() == fx.xz_text(f"iso_{case}.tab.xz")
    assert open(f"{out2}{case}_stats.tab").read() == fx.xz_text(f"iso_{case}_stats.tab.xz")
    # default mode on the cache (no decode at all): the frames come from the files, as in the reference
    out3 = os.path.join(tmp, "out3") + "/"
    assert cli.main(["-i", fa, "-p", pred, "-o", out3, "--model_dir", mdir]) == 0
    assert open(f"{out3}{case}.tab").read() == open(f"{out1}{case}.tab").read()
    assert fx.strip_dates(open(f"{out3}{case}.gff").read()) == fx.strip_dates(open(f"{out1}{case}.gff").read())


def test_comparison_against_reference_on_random_gene_sets():
    """300 perturbed pairs of gene sets (tests/golden/make_golden_isoform_cases.py): both tables as the reference builds them"""
    import sys
    sys.path.insert(0, fx.GOLDEN)
    from make_golden_isoform_cases import case_sequence, frame
    cases = json.loads(fx.xz_text("iso_cases.json.xz"))
    assert len(cases) >= 250
    seen_parents = 0
    for k, c in enumerate(cases):
        l1, l2 = c["labels"]
        frames = {lab: frame([], code) for lab, code in enumerate(fx.CODES)}
        frames[l1] = frame([tuple(g) for g in c["genes_1"]], fx.CODES[l1])
        frames[l2] = frame([tuple(g) for g in c["genes_2"]], fx.CODES[l2])
        P = types.SimpleNamespace(species="fuzz", verbose=False, label_to_gene_coords=frames,
                                  sequence=[Record("c1", "c1", case_sequence(c["seq_seed"], c["seq_len"]))],
                                  updated_genome_labels=[np.array([l1, l1, l2, l2])], tolerance=c["tolerance"],
                                  consecutive_gene_labels=c["consecutive_gene_labels"], label_to_gcode=dict(enumerate(fx.CODES)))
        want = c["expected"]
        if "error" in want:
            with pytest.raises(Exception):
                isoforms.compare_sets_of_predicted_genes(P)
            continue
        coords, _, iso = isoforms.compare_sets_of_predicted_genes(P)
        for col in ("source", "start", "end", "strand", "logodd", "attributes"):
            assert column(iso, col) == want["iso"][col], (k, col)
        coords = coords.reset_index(drop=True)
        for col in ("start", "end", "strand", "source", "gene_label", "updated_gene_label"):
            assert column(coords, col) == want["coords"][col], (k, col)
        seen_parents += sum(1 for v in want["iso"]["logodd"] if v == ".")
    assert seen_parents > 100


def test_one_pass_renumbering_equals_stepwise_insertion():
    """the one-pass form of insert_unseen_genes against the step-by-step restatement of the reference on random tables,
    including shapes it must decline (shared ids, a parent with a foreign id, a gene in front of the first row)"""
    import random
    rng = random.Random(5)
    used_fast = declined = 0
    for it in range(400):
        rows, gid, pos = [], rng.choice([1, 1, 7, 99]), rng.randint(1, 50)
        for _ in range(rng.randint(1, 30)):
            s = pos
            kind = rng.random()
            if kind < 0.25:       # isoform triple
                pid = gid if rng.random() > 0.04 else gid + 40      # case-E parent keeps a foreign id
                rows.append({"start": s, "attributes": f"gene_id {pid}; gene_type bacteria; gc 30; length 300;"})
                rows.append({"start": s + rng.choice([0, 30]), "attributes": f"gene_id {gid}.11; Parent={gid}; gene_type bacteria;"})
                rows.append({"start": s, "attributes": f"gene_id {gid}.15; Parent={gid}; gene_type bacteria;"})
            elif kind < 0.29:     # several rows sharing one id
                for k in range(rng.randint(2, 3)):
                    rows.append({"start": s + 100 * k, "attributes": f"gene_id {gid}; gene_type bacteria; gc 31;"})
            else:
                rows.append({"start": s, "attributes": f"gene_id {gid}; gene_type bacteria; gc 3{gid % 10}; length 1{gid};"})
            gid += 1
            pos += rng.randint(50, 400)
        unseen = sorted(({"start": rng.randint(1 if rng.random() < 0.05 else rows[0]["start"], pos + 200),
                          "attributes": f"gene_id {900 + k}; gene_type bacteria; gc 50;"} for k in range(rng.randint(1, 8))),
                        key=lambda g: g["start"])
        want = isoforms._insert_unseen_stepwise([dict(r) for r in rows], unseen)
        fast = isoforms._insert_unseen_renumbered([dict(r) for r in rows], unseen)
        if fast is None:
            declined += 1
        else:
            used_fast += 1
            assert fast == want, it
        assert isoforms.insert_unseen_genes([dict(r) for r in rows], unseen) == want
    assert used_fast > 150 and declined > 20


# ---- BASELINE.json configs[2] / configs[3]: the reference's other example inputs ---------------------------------------

def _example_text(case):
    gca, phage = fx.xz_text("gca.fa.xz"), fx.xz_text("phage2.fa.xz")
    return gca if case == "gca" else phage + gca          # both.fasta: phage2 first


def _both_gff(code):
    """gmhmmp2's GFF for both.fasta = its GFFs for the two records, gene ids continuing across records"""
    import re
    first = [l for l in fx.xz_text(f"phage2_trained_{code}.gff.xz").split("\n") if l and not l.startswith("#")]
    second = [l for l in fx.xz_text(f"gca_trained_{code}.gff.xz").split("\n") if l and not l.startswith("#")]
    n = len(first)
    second = [re.sub(r"gene_id (\d+);", lambda m: f"gene_id {int(m.group(1)) + n};", l) for l in second]
    return "\n".join(first + second) + "\n"


def test_both_fasta_isoforms_host_part():
    """configs[3] with --isoforms: two contigs, both with switches between codes 15 and 101, five segments on one of them;
    comparison + Results files from the binary's GFFs (no GPU)"""
    ref = json.loads(fx.xz_text("ex_both_iso.json.xz"))
    frames = {}
    for lab, code in enumerate(fx.CODES):
        df = pd.read_csv(io.StringIO(_both_gff(code)), sep="\t", header=None, comment="#", names=GFF_COLUMNS)
        df["source"] = df["source"] + f"_{code}"
        df["length"] = df["end"] - df["start"] + 1
        df["sequence"] = df["sequence"].astype(str)
        df["contig"] = list(df.sequence.values)
        frames[lab] = df
    recs = sorted((Record(i, i, s) for i, s in fx.parse_fasta_text(_example_text("both"))), key=lambda r: r.id)
    assert [r.id for r in recs] == ref["contigs"] and [len(r) for r in recs] == ref["lengths"]
    P = types.SimpleNamespace(species="both_iso", verbose=False, label_to_gene_coords=frames, sequence=recs, nr_contigs=2,
                              updated_genome_labels=[np.array(u) for u in ref["updated"]], tolerance=30, consecutive_gene_labels=5,
                              label_to_gcode=dict(enumerate(fx.CODES)), genome_label=[np.array(l) for l in ref["labels"]],
                              confident_1={k: [tuple(t) for t in v] for k, v in ref["confident_1"].items()},
                              confident_2={k: [tuple(t) for t in v] for k, v in ref["confident_2"].items()})
    P.coords, P.gene_label_to_gcode, P.coords_w_isoforms = isoforms.compare_sets_of_predicted_genes(P)
    assert {k: {str(i): str(v) for i, v in d.items()} for k, d in P.gene_label_to_gcode.items()} == ref["gene_label_to_gcode"]
    import tempfile
    out = tempfile.mkdtemp() + "/"
    HostResults.make(P, "both_iso", out)()
    strip = lambda t: "\n".join(l for l in t.split("\n") if not l.startswith("# output date"))
    assert open(f"{out}both_iso.tab").read() == fx.xz_text("ex_both_iso.tab.xz")
    assert open(f"{out}both_iso_stats.tab").read() == fx.xz_text("ex_both_iso_stats.tab.xz")
    assert strip(open(f"{out}both_iso.gff").read()) == fx.xz_text("ex_both_iso.gff.xz")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["gca", "both", "both_iso"])
def test_reference_examples_end_to_end_on_the_gpu(case, tmp_path):
    """configs[2] (GCA_002924505.1, default mode) and configs[3] (both.fasta, default and --isoforms with the automatic
    stride): predictions of the four models, labels, and the three Results files equal the reference's"""
    import datetime
    from mgcod_b200.pipeline import Pipeline
    from mgcod_b200.results import Results
    tmp = str(tmp_path)
    ref = json.loads(fx.xz_text(f"ex_{case}.json.xz"))
    mdir = os.path.dirname(fx.trained_models(tmp)[11])
    fa = os.path.join(tmp, f"{case}.fasta")
    open(fa, "w").write(_example_text("gca" if case == "gca" else "both"))
    pred, out = os.path.join(tmp, "pred") + "/", os.path.join(tmp, "out") + "/"
    os.makedirs(pred)
    os.makedirs(out)
    P = Pipeline(fa, pred, case, case == "both_iso", False, 3, 5, 5000, None, 30, False, False, "CWD", model_dir=mdir)
    P()
    assert [r.id for r in P.sequence] == ref["contigs"]
    for c in fx.CODES:      # the engine's GFFs against gmhmmp2's
        want = fx.xz_text(f"gca_trained_{c}.gff.xz") if case == "gca" else _both_gff(c)
        assert fx.gff_gene_lines(open(f"{pred}mgm_{c}_{case}.gff").read()) == fx.gff_gene_lines(want)
    if case == "both_iso":
        assert [l.tolist() for l in P.genome_label] == ref["labels"]
        assert [u.tolist() for u in P.updated_genome_labels] == ref["updated"]
        assert [np.round(s, 2).tolist() for s in P.logodd_scores] == ref["scores"]
        assert {k: [list(t) for t in v] for k, v in P.confident_1.items()} == ref["confident_1"]
        assert {k: [list(t) for t in v] for k, v in P.confident_2.items()} == ref["confident_2"]
    else:
        assert P.genome_label == ref["file_label"]
        assert [round(float(df["logodd"].sum()), 2) for df in P.label_to_gene_coords.values()] == ref["file_sums"]
    Results(P, case, pred, fa, out, datetime.datetime(2026, 1, 1, 12, 0, 0), False, False, False, False, "CWD")()
    golden = fx.xz_text(f"ex_{case}.gff.xz").split("\n")
    mine = fx.strip_dates(open(f"{out}{case}.gff").read()).split("\n")
    assert len(golden) == len(mine)
    for g, m in zip(golden, mine):
        if g.startswith("# File with"):
            assert m.split(":")[0] == g.split(":")[0]
            continue
        assert g == m
    assert open(f"{out}{case}.tab").read() == fx.xz_text(f"ex_{case}.tab.xz")
    assert open(f"{out}{case}_stats.tab").read() == fx.xz_text(f"ex_{case}_stats.tab.xz")


class FakeEngine:
    """stands in for the CUDA engine in Results.extract_sequences: the in-frame stop scan by the host restatement, with the
    engine's return layout (counts[g, 3], offsets[3g + 1], positions; stop order TAA, TAG, TGA)"""

    def __init__(self, seqs):
        self.seqs = seqs

    def recoded_stops(self, genes):
        from mgcod_b200.results import Results, _COMP
        counts, offs, pos = np.zeros((len(genes), 3), dtype=np.int32), [0], []
        for g in range(len(genes)):
            s = self.seqs[int(genes["contig"][g])][int(genes["left"][g]) - 1:int(genes["right"][g])]
            if genes["strand"][g] < 0:
                s = s[::-1].translate(_COMP)
            s = s.lower()
            for t, stop in enumerate(("taa", "tag", "tga")):
                hits, _ = Results.find_recoded_stops(s, stop)
                counts[g, t] = len(hits)
                pos.extend(hits)
                offs.append(len(pos))
        return counts, np.array(offs, dtype=np.int64), np.array(pos, dtype=np.int32)


@pytest.mark.parametrize("case", list(CASES))
def test_product_results_class_with_nt_output_on_a_switch_run(case, tmp_path):
    """the product Results class itself (extract_sequences, -NT records, attribute rewriting, all three files) with only the
    device call replaced: the nucleotide FASTA of a switch run has the reference's digest"""
    import datetime
    import hashlib
    from mgcod_b200.results import Results
    P, ref = fake_pipeline(case)
    P.coords, P.gene_label_to_gcode, P.coords_w_isoforms = isoforms.compare_sets_of_predicted_genes(P)
    P.nr_contigs = len(P.sequence)
    P.genome_label = [np.array(l) for l in ref["labels"]]
    P.confident_1 = {k: [tuple(t) for t in v] for k, v in ref["confident_1"].items()}
    P.confident_2 = {k: [tuple(t) for t in v] for k, v in ref["confident_2"].items()}
    ids = [r.id for r in P.sequence]
    P.data = types.SimpleNamespace(ensure_resident=lambda: (FakeEngine([r.seq for r in P.sequence]), ids))
    out = str(tmp_path) + "/"
    Results(P, case, "PRED/", f"TMP/{case}.fasta", out, datetime.datetime(2026, 1, 1, 12, 0, 0), False, False, True, False, "TMP/mgcod")()
    assert hashlib.sha256(open(f"{out}proteins_nt_{case}.fasta", "rb").read()).hexdigest() == ref["nt_sha256"]
    strip = lambda t: "\n".join(l for l in t.split("\n") if not l.startswith("# output date"))
    assert open(f"{out}{case}.tab").read() == fx.xz_text(f"iso_{case}.tab.xz")
    assert open(f"{out}{case}_stats.tab").read() == fx.xz_text(f"iso_{case}_stats.tab.xz")
    assert strip(open(f"{out}{case}.gff").read()) == fx.xz_text(f"iso_{case}.gff.xz")
