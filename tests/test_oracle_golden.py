"""CPU: the oracle (oracle/gmhmm_oracle.c) is pinned against outputs of the reference binary (tests/golden),
and the product's GFF/LST writer reproduces the binary's text byte for byte (date lines excepted)."""
import os

import pytest

from tests import fixtures as fx


@pytest.fixture(scope="module")
def tmpdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("golden"))


def _written(result, model_index, ids, lens, seqfile, modfile, code, tmpdir, fmt):
    out = os.path.join(tmpdir, f"w_{code}.{fmt}")
    result.write(model_index, ids, lens, seqfile, modfile, code, out, fmt)
    return fx.strip_dates(open(out).read())


def _golden_header_names(text):
    seq = mod = None
    for l in text.split("\n"):
        if l.startswith("# File with sequence: "):
            seq = l[len("# File with sequence: "):]
        if l.startswith("# File with MetaGeneMark parameters: "):
            mod = l[len("# File with MetaGeneMark parameters: "):]
    return seq, mod


@pytest.mark.parametrize("code", fx.CODES)
def test_phage2_trained_models(code, tmpdir):
    from mgcod_b200.engine import Result
    from oracle import gmhmm
    (rid, seq), = fx.phage2()
    paths = fx.trained_models(tmpdir)
    golden = fx.xz_text(f"phage2_trained_{code}.gff.xz")
    genes, totals = fx.genes_array_from_oracle(gmhmm.Model(paths[code]), [seq])
    res = Result.from_genes([genes], [totals])
    sf, mf = _golden_header_names(golden)
    assert _written(res, 0, [rid], [len(seq)], sf, mf, code, tmpdir, "gff") == golden
    if code == 15:
        assert _written(res, 0, [rid], [len(seq)], sf, mf, code, tmpdir, "lst") == fx.xz_text("phage2_trained_15.lst.xz")


@pytest.mark.parametrize("code", fx.CODES)
def test_gca_example_trained_models(code, tmpdir):
    """BASELINE.json configs[2]: examples/GCA_002924505.1 (189 826 bp, AT-rich, poly-A ends) under the four trained models"""
    from mgcod_b200.engine import Result
    from oracle import gmhmm
    (rid, seq), = fx.parse_fasta_text(fx.xz_text("gca.fa.xz"))
    paths = fx.trained_models(tmpdir)
    golden = fx.xz_text(f"gca_trained_{code}.gff.xz")
    genes, totals = fx.genes_array_from_oracle(gmhmm.Model(paths[code]), [seq])
    res = Result.from_genes([genes], [totals])
    sf, mf = _golden_header_names(golden)
    assert _written(res, 0, [rid], [len(seq)], sf, mf, code, tmpdir, "gff") == golden


@pytest.mark.parametrize("code", fx.CODES)
def test_synthetic_parametric_models(code, tmpdir):
    from mgcod_b200.engine import Result
    from oracle import gmhmm
    recs, paths, _ = fx.synthetic_case(tmpdir)
    golden = fx.xz_text(f"synth_parametric_{code}.gff.xz")
    genes, totals = fx.genes_array_from_oracle(gmhmm.Model(paths[code]), [s for _, s in recs])
    res = Result.from_genes([genes], [totals])
    sf, mf = _golden_header_names(golden)
    assert _written(res, 0, [i for i, _ in recs], [len(s) for _, s in recs], sf, mf, code, tmpdir, "gff") == golden


def test_odd_records(tmpdir):
    """multi-record input with an empty record, 2-nt and 18-nt records, lower case, a stop-free record and a
    three-bin bacterial+archaeal model of orders 3/1 (exercises gap filling and the order-1 codon rule)."""
    from mgcod_b200.engine import Result
    from oracle import gmhmm
    recs = fx.parse_fasta_text(fx.xz_text("odd_input.fa.xz"))
    mp = fx.write_xz("odd_model_code4.mod.xz", os.path.join(tmpdir, "odd.mod"))
    genes, totals = fx.genes_array_from_oracle(gmhmm.Model(mp), [s for _, s in recs])
    res = Result.from_genes([genes], [totals])
    for fmt in ("gff", "lst"):
        golden = fx.xz_text(f"odd_code4.{fmt}.xz")
        sf, mf = _golden_header_names(golden)
        assert _written(res, 0, [i for i, _ in recs], [len(s) for _, s in recs], sf, mf, 4, tmpdir, fmt) == golden


AMB_MODELS = {"code4": ("odd_model_code4.mod.xz", 4), "trained11": ("trained/mgm_11.mod.xz", 11),
              "trained101": ("trained/mgm_101.mod.xz", 101)}


@pytest.mark.parametrize("tag", sorted(AMB_MODELS))
def test_ambiguous_letters_and_gaps(tag, tmpdir):
    """IUPAC letters, U, lower-case n, runs of 39/40/41/700 ambiguous letters, gaps at either record end and a record
    of nothing but N (tests/golden/make_golden_amb.py): constant window terms, half-GC counting, record splitting."""
    from mgcod_b200.engine import Result
    from oracle import gmhmm
    rel, code = AMB_MODELS[tag]
    recs = fx.parse_fasta_text(fx.xz_text("amb_input.fa.xz"))
    mp = fx.write_xz(rel, os.path.join(tmpdir, f"amb_{tag}.mod"))
    genes, totals = fx.genes_array_from_oracle(gmhmm.Model(mp), [s for _, s in recs])
    res = Result.from_genes([genes], [totals])
    golden = fx.xz_text(f"amb_{tag}.gff.xz")
    sf, mf = _golden_header_names(golden)
    assert _written(res, 0, [i for i, _ in recs], [len(s) for _, s in recs], sf, mf, code, tmpdir, "gff") == golden
