"""GPU parity: CUDA engine through the C-ABI vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu
CODES = (11, 4, 15, 101)


@pytest.fixture(scope="module")
def engine(model_dir):
    from mgcod_b200.engine import Engine
    e = Engine([f"{model_dir}/mgm_{c}.mod" for c in CODES])
    yield e
    e.close()


@pytest.fixture(scope="module")
def oracles(model_dir):
    from oracle import gmhmm
    return [gmhmm.Model(f"{model_dir}/mgm_{c}.mod") for c in CODES]


def _synthetic(seed, total):
    from mgcod_b200.engine import synth_metagenome
    buf, offs, truth = synth_metagenome(seed, total, threads=4)
    return buf, offs, [buf[offs[i]:offs[i + 1]].tobytes() for i in range(len(truth))]


def test_candidates_and_chain_scores(engine, oracles):
    buf, offs, seqs = _synthetic(7, 120000)
    engine.decode_buffer(buf, offs)
    for mi, orc in enumerate(oracles):
        diffs = parity.compare_candidates(engine.candidates(mi), orc, seqs)
        assert not diffs, f"model {CODES[mi]}: " + "; ".join(diffs[:6])


def test_genes_match_oracle(engine, oracles):
    buf, offs, seqs = _synthetic(11, 300000)
    res = engine.decode_buffer(buf, offs)
    for mi, orc in enumerate(oracles):
        rows, totals = parity.oracle_genes(orc, seqs)
        diffs = parity.compare_genes(res.genes(mi), rows)
        assert not diffs, f"model {CODES[mi]}: " + "; ".join(diffs[:6])
        np.testing.assert_allclose(res.contig_totals(mi), totals, rtol=1e-9, atol=1e-9)
        cents = sum(int(round(r[4] * 100)) for r in rows)
        assert res.model_cents()[mi] == cents


def _sprinkle(seqs, seed):
    """ambiguous letters into clean contigs: singles, short runs, and runs of 39..300 (>= 40 splits the record)"""
    import random
    rng = random.Random(seed)
    out = []
    for s in seqs:
        s = bytearray(s)
        L = len(s)
        for _ in range(rng.choice([0, 1, 3, 10, 30])):
            a = rng.randrange(0, L)
            n = rng.choice([1, 1, 1, 2, 5, 20, 39, 40, 41, 100, 300])
            for i in range(a, min(L, a + n)):
                s[i] = rng.choice(b"NNNNRYKMSWBDHVnU")
        if rng.random() < 0.15:
            s[:50] = b"N" * min(50, L)
        if rng.random() < 0.15:
            s[-45:] = b"N" * min(45, L)
        out.append(bytes(s))
    return out


@pytest.mark.parametrize("seed", [3, 4])
def test_ambiguous_letters_and_gaps_match_oracle(engine, oracles, seed):
    buf, offs, seqs = _synthetic(30 + seed, 250000)
    seqs = _sprinkle(seqs, seed)
    res = engine.decode(seqs)
    assert res.n_contigs == len(seqs)
    for mi, orc in enumerate(oracles):
        rows, totals = parity.oracle_genes(orc, seqs)
        diffs = parity.compare_genes(res.genes(mi), rows)
        assert not diffs, f"model {CODES[mi]}: " + "; ".join(diffs[:6])
        np.testing.assert_allclose(res.contig_totals(mi), totals, rtol=1e-9, atol=1e-9)
        assert res.model_cents()[mi] == sum(int(round(r[4] * 100)) for r in rows)
    # the window kernel sees the caller's records, not the gap-free pieces
    lens = np.array([len(s) for s in seqs], np.int64)
    w = 5000
    offs_w, labels, cents = engine.window_labels(lens, np.full(len(lens), w, np.int32), w, 1, True)
    assert len(offs_w) == len(seqs) + 1
    for mi in range(len(CODES)):
        g = res.genes(mi)
        per_contig = np.bincount(g["contig"], weights=np.rint(g["logodd"] * 100), minlength=len(lens)).astype(np.int64)
        wsum = np.add.reduceat(cents[:, mi], offs_w[:-1])
        long_enough = lens > w
        # non-overlapping windows partition a record's genes (the spanning-gene fallback can only add to empty windows)
        assert (wsum[long_enough] >= per_contig[long_enough]).all()
        assert (wsum[long_enough] - per_contig[long_enough]).sum() <= 0.05 * max(per_contig.sum(), 1)
    # a second decode of the resident (re-slotted) batch gives the same genes
    res2 = engine.decode_resident()
    for mi in range(len(CODES)):
        assert np.array_equal(res.genes(mi), res2.genes(mi))


def test_long_contigs_chained_in_pieces_match_oracle(engine, oracles):
    """contigs far longer than one chain piece (8 kb): cut points, slot order and stitched totals.  The candidate dump
    re-adds the prefix score, so chain scores compare with the sequential oracle too."""
    from mgcod_b200.engine import synth_metagenome
    buf, offs, truth = synth_metagenome(77, 1_300_000, min_len=600_000, max_len=700_000, threads=4)
    seqs = [buf[offs[i]:offs[i + 1]].tobytes() for i in range(len(truth))]
    assert len(seqs) >= 2
    res = engine.decode_buffer(buf, offs)
    for mi in (0, 2):
        rows, totals = parity.oracle_genes(oracles[mi], seqs)
        diffs = parity.compare_genes(res.genes(mi), rows)
        assert not diffs, f"model {CODES[mi]}: " + "; ".join(diffs[:6])
        np.testing.assert_allclose(res.contig_totals(mi), totals, rtol=1e-9, atol=1e-9)
    diffs = parity.compare_candidates(engine.candidates(1), oracles[1], seqs)
    assert not diffs, "; ".join(diffs[:6])


def test_genes_longer_than_a_chain_piece(engine, oracles):
    """18 kb genes (both strands) span more bucket boundaries than the chain planner's look-ahead (2 x PIECE_BUCKETS at
    most): it must not cut inside them — they are counted through the difference array."""
    buf, offs, seqs = _synthetic(91, 200000)
    codons = []
    for donor in sorted(seqs, key=len, reverse=True)[:4]:
        genes, _ = oracles[0].decode(donor)
        for g in genes:
            if g["strand"] > 0 and not (g["partial_left"] or g["partial_right"]):
                body = donor[g["left"] - 1 + 3:g["right"] - 3]
                codons += [body[i:i + 3] for i in range(0, len(body), 3)]
    codons = [c for c in codons if c not in (b"TAA", b"TAG", b"TGA")]
    assert len(codons) >= 6000
    long_gene = b"ATG" + b"".join(codons[:6000]) + b"TAA"
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    hosts = sorted(seqs, key=len, reverse=True)[:2]
    ia, ib = len(hosts[0]) // 3 + 1, len(hosts[1]) // 2 + 2
    a = hosts[0][:ia] + long_gene + hosts[0][ia:]
    b = hosts[1][:ib] + long_gene.translate(comp)[::-1] + hosts[1][ib:]
    test_seqs = [a, b, long_gene, seqs[0]]
    res = engine.decode(test_seqs)
    found_long = 0
    for mi, orc in enumerate(oracles):
        rows, totals = parity.oracle_genes(orc, test_seqs)
        found_long += sum(1 for r in rows if r[2] - r[1] + 1 >= 17000)
        diffs = parity.compare_genes(res.genes(mi), rows)
        assert not diffs, f"model {CODES[mi]}: " + "; ".join(diffs[:6])
        np.testing.assert_allclose(res.contig_totals(mi), totals, rtol=1e-9, atol=1e-9)
        diffs = parity.compare_candidates(engine.candidates(mi), orc, test_seqs)
        assert not diffs, f"model {CODES[mi]}: " + "; ".join(diffs[:6])
    assert found_long >= 4          # the long gene is predicted, so candidates spanning > 128 boundaries existed


def test_many_short_and_odd_length_contigs_match_oracle(engine, oracles):
    """hundreds of contigs of 0..700 nt (lengths around the word, bucket and minimum-gene-length boundaries), cut from
    gene-dense sequence at random offsets, plus their reverse complements: slot, bucket, piece and edge-partial logic"""
    import random
    rng = random.Random(20260222)
    buf, offs, seqs = _synthetic(123, 150000)
    src = b"".join(seqs)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    lens = [0, 1, 2, 3, 5, 6, 31, 32, 33, 63, 64, 65, 89, 90, 91, 92, 93, 95, 96, 97, 127, 128, 129, 130, 191, 192, 255, 256, 257]
    test_seqs = []
    for L in lens * 4 + [rng.randrange(90, 700) for _ in range(250)]:
        a = rng.randrange(0, len(src) - L - 1)
        s = src[a:a + L]
        test_seqs.append(s if rng.random() < 0.5 else s.translate(comp)[::-1])
    rng.shuffle(test_seqs)
    res = engine.decode(test_seqs)
    total = 0
    for mi, orc in enumerate(oracles):
        rows, totals = parity.oracle_genes(orc, test_seqs)
        total += len(rows)
        diffs = parity.compare_genes(res.genes(mi), rows)
        assert not diffs, f"model {CODES[mi]}: " + "; ".join(diffs[:6])
        np.testing.assert_allclose(res.contig_totals(mi), totals, rtol=1e-9, atol=1e-9)
    assert total > 500
    diffs = parity.compare_candidates(engine.candidates(3), oracles[3], test_seqs)
    assert not diffs, "; ".join(diffs[:6])
