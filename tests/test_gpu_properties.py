"""GPU: edge cases and size-independent properties at sizes the oracle would not finish quickly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CODES = (11, 4, 15, 101)


@pytest.fixture(scope="module")
def engine(model_dir):
    from mgcod_b200.engine import Engine
    e = Engine([f"{model_dir}/mgm_{c}.mod" for c in CODES])
    yield e
    e.close()


def test_letters_outside_the_iupac_alphabet_are_rejected_loudly(engine):
    """gmhmmp2: "error, unexpected letter found in FASTA sequence" for anything but ACGTU + RYSWKMBDHVN"""
    from mgcod_b200 import _lib
    from mgcod_b200.engine import MgcodError
    for bad in ("X", "-", "*", "1", "E"):
        with pytest.raises(MgcodError) as e:
            engine.decode(["ACGT" * 100 + bad + "ACGT" * 100])
        assert e.value.code == _lib.MGC_E_ALPHABET
        assert "unexpected letter" in str(e.value)
    engine.decode(["ACGT" * 100 + "NRYSWKMBDHVUnryswkmbdhvu" + "ACGT" * 100])   # accepted


def test_tiny_and_empty_contigs(engine):
    res = engine.decode(["", "A", "ACGTAC", "ATG" + "GCA" * 40 + "TAA", "ATGAAATAA"])
    for mi in range(4):
        g = res.genes(mi)
        assert set(g["contig"].tolist()) <= {3}
    assert engine.stats()["n_genes"] >= 0


def test_small_orders_and_few_bins_match_oracle(small_model_dir):
    """order 3/1 models with four GC bins per domain (gap filling above/below the provided bins)"""
    from mgcod_b200.engine import Engine, synth_metagenome
    from oracle import gmhmm
    from tests import parity
    buf, offs, truth = synth_metagenome(21, 80000, threads=2)
    seqs = [buf[offs[i]:offs[i + 1]].tobytes() for i in range(len(truth))]
    eng = Engine([f"{small_model_dir}/mgm_{c}.mod" for c in CODES])
    res = eng.decode_buffer(buf, offs)
    for mi, c in enumerate(CODES):
        orc = gmhmm.Model(f"{small_model_dir}/mgm_{c}.mod")
        assert not parity.compare_candidates(eng.candidates(mi), orc, seqs)
        rows, _ = parity.oracle_genes(orc, seqs)
        assert not parity.compare_genes(res.genes(mi), rows)
    eng.close()


def test_full_size_properties(engine):
    """64 Mbp x 4 models: deterministic, genes ordered and inside their contigs, stop-codon/frame consistent,
    chain totals consistent with the genes, window sums add up to the per-model sums, labels follow the planted code."""
    from mgcod_b200.engine import synth_metagenome
    buf, offs, truth = synth_metagenome(20260101, 64_000_000, threads=8)
    lens = np.diff(offs)
    r1 = engine.decode_buffer(buf, offs)
    g1 = [r1.genes(m) for m in range(4)]
    c1 = r1.model_cents().copy()
    r2 = engine.decode_buffer(buf, offs)
    for m in range(4):
        g2 = r2.genes(m)
        assert g1[m].tobytes() == g2.tobytes(), "decode is not deterministic"
        g = g1[m]
        assert (np.diff(g["contig"]) >= 0).all()
        same = np.diff(g["contig"]) == 0
        assert (np.diff(g["left"])[same] > 0).all() and (np.diff(g["right"])[same] > 0).all(), "chain order violated"
        assert (g["left"] >= 1).all() and (g["right"] <= lens[g["contig"]]).all()
        assert ((g["right"] - g["left"] + 1) % 3 == 0).all() and (g["logodd"] > 0).all()
        assert (g["right"] - g["left"] + 1 >= 90).all()
        # complete genes end in a stop codon of the model's genetic code on their strand
        stops = {11: {b"TAA", b"TAG", b"TGA"}, 4: {b"TAA", b"TAG"}, 15: {b"TAA", b"TGA"}, 101: {b"TAG", b"TGA"}}[CODES[m]]
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        for k in np.random.default_rng(m).choice(len(g), 3000, replace=False):
            x = g[k]
            base = offs[x["contig"]]
            if x["strand"] > 0 and not x["partial_right"]:
                assert buf[base + x["right"] - 3: base + x["right"]].tobytes() in stops
            if x["strand"] < 0 and not x["partial_left"]:
                assert buf[base + x["left"] - 1: base + x["left"] + 2].tobytes().translate(comp)[::-1] in stops
        # chain totals: non-negative, and positive exactly where genes were found.  (They are NOT bounded by the sum of
        # the traced genes' scores: the reference's shared back-pointer of reverse-strand ORFs can trace a path that
        # differs from the one the total was computed on — DESIGN.md "chain".)
        tot = r2.contig_totals(m)
        has = np.bincount(g["contig"], minlength=len(lens)) > 0
        assert (tot >= 0).all() and ((tot > 0) == has).all()
        assert int(np.rint(g["logodd"] * 100).astype(np.int64).sum()) == int(c1[m])
    # non-overlapping windows (stride == window) partition the genes by left coordinate
    w = 5000
    offs_w, labels, cents = engine.window_labels(lens, np.full(len(lens), w, np.int32), w, 1, True)
    long_enough = lens > w
    for m in range(4):
        g = g1[m]
        per_contig = np.bincount(g["contig"], weights=np.rint(g["logodd"] * 100), minlength=len(lens)).astype(np.int64)
        wsum = np.add.reduceat(cents[:, m], offs_w[:-1])
        # the spanning-gene fallback can only add to empty windows, so window sums are >= the gene sums
        assert (wsum[long_enough] >= per_contig[long_enough]).all()
        assert (wsum[long_enough] - per_contig[long_enough]).sum() <= 0.02 * per_contig.sum()
    # per-contig majority label recovers the planted genetic code for the long contigs
    code_of = np.array(CODES)
    hit = tot_c = 0
    for c in np.flatnonzero(lens > 20000):
        lab = labels[offs_w[c]:offs_w[c + 1]]
        hit += code_of[np.bincount(lab, minlength=4).argmax()] == truth[c]
        tot_c += 1
    assert hit / tot_c > 0.97


def test_decode_stream_equals_batch_by_batch(engine):
    """batch i+1 travels on the copy stream while batch i is decoded; results equal those of separate calls, a
    prefetch that is never adopted is harmless"""
    from mgcod_b200.engine import synth_metagenome
    batches = []
    for seed in (5, 6, 7):
        buf, offs, _ = synth_metagenome(seed, 60000 + 20000 * seed, threads=2)
        batches.append((buf, offs))
    want = [[engine.decode_buffer(b, o).genes(m) for m in range(4)] for b, o in batches]
    got = [[r.genes(m) for m in range(4)] for r in engine.decode_stream(iter(batches))]
    assert len(got) == 3
    for w, g in zip(want, got):
        for a, b in zip(w, g):
            assert np.array_equal(a, b)
    engine.prefetch(*batches[0])
    r = engine.decode_buffer(*batches[1])      # different batch: the prefetched letters are dropped
    assert np.array_equal(r.genes(0), want[1][0])
    r = engine.decode_buffer(*batches[0])
    assert np.array_equal(r.genes(2), want[0][2])


@pytest.mark.parametrize("kc,kn,gcode,domains,bins", [
    (0, 0, 11, ("B",), [40]),
    (1, 0, 4, ("B", "A"), [35, 45]),
    (2, 1, 25, ("B",), [30, 38, 46]),
    (2, 2, 1, ("B", "A"), [40]),            # genetic code 1: ATG is the only start codon
    (4, 0, 15, ("B",), [33, 44]),
    (5, 2, 101, ("B", "A"), [36, 42]),      # order 5: tables too large for shared memory, gathered from global memory
    (5, 3, 11, ("B",), [41]),
])
def test_model_orders_codes_and_domains_match_oracle(kc, kn, gcode, domains, bins, tmp_path):
    """every Markov order the parser accepts (0..5 coding, 0..3 non-coding), the genetic codes outside Mgcod's four (25, 1),
    bacterial-only and two-domain files, one to three GC bins — genes and chain totals against the oracle"""
    import lzma, os
    from mgcod_b200 import synth
    from mgcod_b200.engine import Engine, synth_metagenome
    from oracle import gmhmm
    from tests import parity
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phage2.fa.xz")
    template = "".join(l.strip() for l in lzma.open(golden, "rt") if not l.startswith(">")).upper()
    mod = str(tmp_path / "m.mod")
    synth.write_mod(mod, synth.trained_sections(synth.encode(template), gcode, kc, kn, bins=bins, domains=domains))
    buf, offs, truth = synth_metagenome(300 + kc * 10 + kn, 90000, threads=2)
    seqs = [buf[offs[i]:offs[i + 1]].tobytes() for i in range(len(truth))] + [template[20000:50000].encode()]
    eng = Engine([mod])
    res = eng.decode(seqs)
    orc = gmhmm.Model(mod)
    rows, totals = parity.oracle_genes(orc, seqs)
    assert len(rows) > 20
    diffs = parity.compare_genes(res.genes(0), rows)
    assert not diffs, "; ".join(diffs[:6])
    np.testing.assert_allclose(res.contig_totals(0), totals, rtol=1e-9, atol=1e-9)
    diffs = parity.compare_candidates(eng.candidates(0), orc, seqs)
    assert not diffs, "; ".join(diffs[:6])
    eng.close()


@pytest.mark.parametrize("walk", ["pair", "single"])
def test_models_of_different_orders_in_one_engine(walk, tmp_path, monkeypatch):
    """four parameter files with different Markov orders (4/2, 3/1, 2/0, 1/1) in ONE engine: the pair walk runs its generic
    kernel (orders read per work unit, a different table layout and replica count per model, tables reloaded when the model
    changes); MGCOD_B200_WALK=single sends the same input through the single-position walk.  Both equal the oracle."""
    import lzma, os
    from mgcod_b200 import synth
    from mgcod_b200.engine import Engine, synth_metagenome
    from oracle import gmhmm
    from tests import parity
    if walk == "single":
        monkeypatch.setenv("MGCOD_B200_WALK", "single")
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phage2.fa.xz")
    template = "".join(l.strip() for l in lzma.open(golden, "rt") if not l.startswith(">")).upper()
    specs = [(4, 2, 11, ("B", "A"), [30, 45]), (3, 1, 4, ("B", "A"), [35]), (2, 0, 15, ("B",), [33, 44]), (1, 1, 101, ("B", "A"), [40])]
    mods = []
    for i, (kc, kn, gcode, domains, bins) in enumerate(specs):
        mods.append(str(tmp_path / f"m{i}.mod"))
        synth.write_mod(mods[-1], synth.trained_sections(synth.encode(template), gcode, kc, kn, bins=bins, domains=domains))
    buf, offs, truth = synth_metagenome(777, 150000, threads=2)
    seqs = [buf[offs[i]:offs[i + 1]].tobytes() for i in range(len(truth))] + [template[60000:90000].encode()]
    eng = Engine(mods)
    res = eng.decode(seqs)
    for mi, mod in enumerate(mods):
        orc = gmhmm.Model(mod)
        rows, totals = parity.oracle_genes(orc, seqs)
        assert len(rows) > 20
        diffs = parity.compare_genes(res.genes(mi), rows)
        assert not diffs, f"model {mi}: " + "; ".join(diffs[:6])
        np.testing.assert_allclose(res.contig_totals(mi), totals, rtol=1e-9, atol=1e-9)
    eng.close()


def test_low_complexity_repeats_grow_the_scratch_buffers(engine):
    """(ATG)n / (CAT)n repeats hold a start codon every three letters — far more starts per letter than the scratch
    capacities assume; the engine must notice, grow and still agree with the oracle"""
    from mgcod_b200.engine import synth_metagenome
    from oracle import gmhmm
    from tests import parity
    import tempfile
    from mgcod_b200 import synth
    buf, offs, _ = synth_metagenome(55, 30000, threads=2)
    flank = buf[offs[0]:offs[1]].tobytes()[:6000]
    a = flank[:3000] + b"ATG" * 14000 + b"TAA" + flank[3000:]
    b = flank[:2001] + b"TTA" + b"CAT" * 14000 + flank[2001:]
    seqs = [a, b, flank]
    res = engine.decode(seqs)
    st = engine.stats()
    assert st["n_starts"] > 4 * 2 * 14000 * 0.9
    d = tempfile.mkdtemp()
    paths = synth.write_model_set(d, kind="parametric")
    for mi, code in ((0, 11), (3, 101)):
        rows, totals = parity.oracle_genes(gmhmm.Model(paths[code]), seqs)
        diffs = parity.compare_genes(res.genes(mi), rows)
        assert not diffs, f"model {code}: " + "; ".join(diffs[:6])
        np.testing.assert_allclose(res.contig_totals(mi), totals, rtol=1e-9, atol=1e-9)
    # the engine keeps working at normal sizes afterwards
    res2 = engine.decode([flank])
    assert np.array_equal(res2.genes(0)[["left", "right"]], res.genes(0)[res.genes(0)["contig"] == 2][["left", "right"]])


def test_decode_large_streams_batches_and_merges(engine, tmp_path):
    """inputs beyond one device batch: whole contigs per batch, merged gene tables equal the single-batch decode, and the
    merged (host-only) result writes the same GFF"""
    from mgcod_b200.engine import synth_metagenome
    buf, offs, truth = synth_metagenome(41, 400000, threads=2)
    one = engine.decode_buffer(buf, offs)
    want = [one.genes(m) for m in range(4)]
    want_tot = [one.contig_totals(m) for m in range(4)]
    ids = [f"c{i}" for i in range(len(truth))]
    lens = np.diff(offs)
    one.write(2, ids, lens, "s.fa", "m.mod", 15, str(tmp_path / "one.gff"), "gff")
    many = engine.decode_large(buf, offs, max_bp=90000)
    for m in range(4):
        assert np.array_equal(many.genes(m), want[m])
        np.testing.assert_array_equal(many.contig_totals(m), want_tot[m])
    assert (many.model_cents() == one.model_cents()).all()
    many.write(2, ids, lens, "s.fa", "m.mod", 15, str(tmp_path / "many.gff"), "gff")
    strip = lambda t: "\n".join(l for l in t.split("\n") if not l.startswith("# output date"))
    assert strip(open(tmp_path / "one.gff").read()) == strip(open(tmp_path / "many.gff").read())


def test_results_do_not_depend_on_the_walk_tiling(engine, monkeypatch):
    """k_walk's items are ordered by (position tile, section pair, length class); a 200 Mbp batch has 4 tiles of 2 M plane
    words, its 40 Mbp slices one each.  The genes of a contig must not depend on the batch it was decoded in."""
    from mgcod_b200.engine import synth_metagenome
    buf, offs, truth = synth_metagenome(77, 200_000_000, threads=8)
    whole = engine.decode_buffer(buf, offs)
    want = [whole.genes(m).copy() for m in range(4)]
    tot = [whole.contig_totals(m).copy() for m in range(4)]
    parts = engine.decode_large(buf, offs, max_bp=40_000_000)
    for m in range(4):
        assert np.array_equal(parts.genes(m), want[m]), m
        np.testing.assert_array_equal(parts.contig_totals(m), tot[m])
    # and with tiles of 4096 words (32 tiles, the cap) on a small batch
    small_buf, small_offs = buf[:offs[40]], offs[:41]
    ref = engine.decode_buffer(small_buf, small_offs)
    ref_genes = [ref.genes(m).copy() for m in range(4)]
    monkeypatch.setenv("MGCOD_B200_WALK_TILE_WORDS", "4096")
    tiled = engine.decode_buffer(small_buf, small_offs)
    for m in range(4):
        assert np.array_equal(tiled.genes(m), ref_genes[m]), m
