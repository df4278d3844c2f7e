"""Results mirror (mgcod_b200/results.py) against outputs of the reference's own Results class
(tests/golden/make_golden_results.py ran bin/pipeline.py + bin/results.py unmodified in the build container)."""
import datetime
import json
import os

import numpy as np
import pytest

from tests import fixtures as fx

HERE = os.path.dirname(os.path.abspath(__file__))


def test_find_recoded_stops_restatement_matches_reference():
    from mgcod_b200.results import Results
    cases = json.load(open(os.path.join(HERE, "golden", "recoded_ref.json")))
    assert len(cases) == 200
    for c in cases:
        work = c["seq"]
        for stop in ("tag", "tga", "taa"):
            pos, work = Results.find_recoded_stops(work, stop)
            assert pos == c["recodings"][stop.upper()], c["seq"]
        assert work == c["marked"]


def test_translate_seq_recodes_and_marks():
    from types import SimpleNamespace
    from mgcod_b200.results import Results
    nt = ("1 x", "d", "atgTGAaaatagtaa")
    rid, desc, aa = Results.translate_seq(SimpleNamespace(source="GeneMark.hmm2_4"), nt, {"TAG": [], "TGA": [1], "TAA": []})
    assert (rid, desc, aa) == ("1 x", "d", "mWk")           # TGA -> W (upper: recoded), TAG and TAA are stops, dropped
    _, _, aa = Results.translate_seq(SimpleNamespace(source="GeneMark.hmm2_11"), ("1 x", "d", "atgnnnaaataa"), {"TAG": [], "TGA": [], "TAA": []})
    assert aa == "mxk"


def _inputs():
    phage = fx.xz_text("phage2.fa.xz")
    seq = "".join(l.strip() for l in phage.split("\n") if not l.startswith(">"))
    return {"phage2": phage,
            "three_contigs": f">zeta part\n{seq[1000:31000]}\n>alpha\n{seq[60000:100000]}\n>mid x y\n{seq[120000:150000]}\n"}


def _run(case, tmp, via_cli):
    from mgcod_b200.pipeline import Pipeline
    from mgcod_b200.results import Results
    paths = fx.trained_models(tmp)
    mdir = os.path.dirname(paths[11])
    fa = os.path.join(tmp, f"{case}.fasta")
    open(fa, "w").write(_inputs()[case])
    pred, out = os.path.join(tmp, f"pred_{case}_{via_cli}") + "/", os.path.join(tmp, f"out_{case}_{via_cli}") + "/"
    if via_cli:
        from mgcod_b200 import cli
        assert cli.main(["-i", fa, "-p", pred, "-o", out, "-NT", "--model_dir", mdir, "--logfile", os.path.join(tmp, "log", "x.log")]) == 0
    else:
        os.makedirs(pred)
        os.makedirs(out)
        P = Pipeline(fa, pred, case, False, False, 3, 5, 5000, 5000, 30, False, False, "CWD", model_dir=mdir)
        P()
        Results(P, case, pred, fa, out, datetime.datetime(2026, 1, 1, 12, 0, 0), False, False, True, False, "CWD")()
    return fa, out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["phage2", "three_contigs"])
@pytest.mark.parametrize("via_cli", [False, True])
def test_results_files_equal_reference(case, via_cli, tmp_path):
    tmp = str(tmp_path)
    fa, out = _run(case, tmp, via_cli)
    golden = fx.xz_text(f"results_{case}.gff.xz").split("\n")
    mine = fx.strip_dates(open(f"{out}{case}.gff").read()).split("\n")
    assert len(golden) == len(mine)
    for g, m in zip(golden, mine):
        if g.startswith("# File with"):     # paths differ (temp dirs); the rest of the header must agree
            assert m.split(":")[0] == g.split(":")[0]
            continue
        assert g == m
    assert open(f"{out}{case}.tab").read() == fx.xz_text(f"results_{case}.tab.xz")
    assert open(f"{out}{case}_stats.tab").read() == fx.xz_text(f"results_{case}_stats.tab.xz")
    if case == "three_contigs":
        assert open(f"{out}proteins_nt_{case}.fasta").read() == fx.xz_text(f"results_{case}_nt.fasta.xz")


@pytest.mark.gpu
def test_recoded_stops_kernel_matches_host_restatement(tmp_path):
    """every gene of four models on contigs with gaps and ambiguous letters; both strands, partial genes"""
    from mgcod_b200.engine import Engine
    from mgcod_b200.results import Results
    recs = fx.parse_fasta_text(fx.xz_text("amb_input.fa.xz"))
    paths = fx.trained_models(str(tmp_path))
    eng = Engine([paths[c] for c in fx.CODES])
    seqs = [s for _, s in recs]
    res = eng.decode(seqs)
    comp = str.maketrans("ACGTacgt", "TGCAtgca")
    for mi in range(4):
        g = res.genes(mi)
        counts, offs, pos = eng.recoded_stops(g)
        assert counts.shape == (len(g), 3)
        for k in range(len(g)):
            s = seqs[g["contig"][k]][g["left"][k] - 1:g["right"][k]]
            if g["strand"][k] < 0:
                s = s[::-1].translate(comp)
            s = s.lower()
            for t, stop in enumerate(("taa", "tag", "tga")):
                want, _ = Results.find_recoded_stops(s, stop)
                got = pos[offs[3 * k + t]:offs[3 * k + t + 1]].tolist()
                assert got == want and counts[k, t] == len(want), (mi, k, stop)
    eng.close()


@pytest.mark.gpu
def test_multi_genome_driver(tmp_path):
    """bin/multiprocess_mgcod.py mirror: a list of genomes through one engine; outputs equal the single-genome CLI's goldens,
    a missing file is skipped"""
    from mgcod_b200 import multi
    tmp = str(tmp_path)
    paths = fx.trained_models(tmp)
    mdir = os.path.dirname(paths[11])
    files = []
    for case, text in _inputs().items():
        fa = os.path.join(tmp, f"{case}.fasta")
        open(fa, "w").write(text)
        files.append(fa)
    lst = os.path.join(tmp, "genomes.txt")
    open(lst, "w").write("\n".join(files + [os.path.join(tmp, "missing.fasta")]) + "\n")
    out, pred = os.path.join(tmp, "out") + "/", os.path.join(tmp, "pred") + "/"
    rc = multi.main(["--genomes", lst, "-o", out, "-p", pred, "-NT", "--model_dir", mdir, "--threads", "4",
                     "--logfile", os.path.join(tmp, "log", "m.log")])
    assert rc == 0
    for case in _inputs():
        assert open(f"{out}{case}.tab").read() == fx.xz_text(f"results_{case}.tab.xz")
        assert open(f"{out}{case}_stats.tab").read() == fx.xz_text(f"results_{case}_stats.tab.xz")
