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
