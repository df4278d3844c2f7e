"""CPU: the host layer above the C-ABI end to end — Data / Pipeline / Results exactly as the command line drives them — with the
decode taken from the CPU oracle and the device aggregation calls from tests/hostfake.py (numpy restatements with the engine's
return layouts).  What runs is the product's Python: lazy frames from gene arrays, the one-call labelling path, vectorised
single-code contigs, the isoform comparison, the three result files — against the reference's own files (tests/golden/iso_*)."""
import datetime
import json
import os

import numpy as np
import pandas as pd
import pytest

from tests import fixtures as fx
from tests import hostfake
from tests.test_isoforms import CASES, _case_text


def _fake_data_class(result, eng):
    from mgcod_b200 import data as data_mod

    class FakeData(data_mod.Data):
        def _decode(self):
            self.engine = eng
            eng._resident_owner = self
            return result

        def ensure_resident(self):
            return eng, [r.id for r in self._file_order]
    return FakeData


@pytest.mark.parametrize("case", list(CASES))
def test_pipeline_and_results_on_the_host(case, tmp_path, monkeypatch):
    import mgcod_b200.pipeline as pm
    from mgcod_b200.results import Results
    tmp = str(tmp_path)
    ref = json.loads(fx.xz_text(f"iso_{case}.json.xz"))
    a = ref["args"]
    paths = fx.trained_models(tmp)
    fa = os.path.join(tmp, f"{case}.fasta")
    open(fa, "w").write(_case_text(case))
    recs = fx.parse_fasta_text(open(fa).read())
    result, genes = hostfake.oracle_result([paths[c] for c in fx.CODES], [s for _, s in recs], threads=4)
    eng = hostfake.HostEngine(genes, [s for _, s in recs])
    monkeypatch.setattr(pm, "Data", _fake_data_class(result, eng))
    pred, out = os.path.join(tmp, "pred") + "/", os.path.join(tmp, "out") + "/"
    os.makedirs(pred)
    os.makedirs(out)
    P = pm.Pipeline(fa, pred, case, True, False, a["n"], a["genes"], a["window"], a["stride"], a["tol"], False, False, "CWD",
                    model_dir=os.path.dirname(paths[11]))
    P()
    assert [u.tolist() for u in P.updated_genome_labels] == ref["updated"]
    assert [l.tolist() for l in P.genome_label] == ref["labels"] if isinstance(P.genome_label, list) else True
    Results(P, case, pred, fa, out, datetime.datetime(2026, 1, 1, 12, 0, 0), False, False, False, False, "CWD")()
    P.close()
    assert open(f"{out}{case}.tab").read() == fx.xz_text(f"iso_{case}.tab.xz")
    assert open(f"{out}{case}_stats.tab").read() == fx.xz_text(f"iso_{case}_stats.tab.xz")
    golden = fx.xz_text(f"iso_{case}.gff.xz").split("\n")
    mine = fx.strip_dates(open(f"{out}{case}.gff").read()).split("\n")
    assert len(golden) == len(mine)
    for g, m in zip(golden, mine):
        if not g.startswith("# File with"):
            assert g == m
    # the four mgm_<code>_<species>.gff files were written in the background and are complete
    for code in fx.CODES:
        assert os.path.getsize(f"{pred}mgm_{code}_{case}.gff") > 1000


def test_frames_from_arrays_equal_frames_parsed_from_the_gff_files(tmp_path):
    """GeneFrames (built from the decoded gene arrays) against what the reference does — pd.read_csv of the GFF text the
    writer produced (bin/data.py:66-96 + the contig column of bin/pipeline.py:73-76): equal column by column, dtypes included"""
    from mgcod_b200.data import GFF_COLUMNS, GeneFrames
    tmp = str(tmp_path)
    recs, paths, _ = fx.synthetic_case(tmp)
    seqs, ids = [s for _, s in recs], [i for i, _ in recs]
    res, genes = hostfake.oracle_result([paths[c] for c in fx.CODES], seqs, threads=4)
    gf = GeneFrames(genes, ids)
    for mi, code in enumerate(fx.CODES):
        out = os.path.join(tmp, f"o{code}.gff")
        res.write(mi, ids, [len(s) for s in seqs], "x.fa", "m.mod", code, out, "gff")
        df = pd.read_csv(out, sep="\t", header=None, comment="#", names=GFF_COLUMNS)
        df["source"] = df["source"] + f"_{code}"
        df["length"] = df["end"] - df["start"] + 1
        df["sequence"] = df["sequence"].astype(str)
        df.loc[:, "contig"] = list(df.sequence.values)
        assert list(df.dtypes) == list(gf[mi].dtypes) and df.equals(gf[mi])
        some = ids[len(ids) // 2]
        assert df[df.contig == some].equals(gf.contig_frame(mi, some))
        a, b = gf.contig_range(mi, some)
        assert df[df.contig == some].equals(GeneFrames(genes, ids).contig_frame(mi, some)) and b - a == (df.contig == some).sum()


def test_cents_follow_printf_rounding():
    from mgcod_b200.data import cents_of
    v = np.array([1.005, 2.675, 0.125, 10.0, 0.004999999, 33.335, 1e-9, 123456.785])
    assert cents_of(v).tolist() == [int(round(float("%.2f" % x) * 100)) for x in v]
