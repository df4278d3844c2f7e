"""CPU: the C-ABI library loads, exports every symbol include/mgcod_b200.h declares, and its host-only entry
points behave (no compute calls without a GPU)."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

from mgcod_b200 import _lib, synth
from mgcod_b200.engine import Engine, MgcodError, ModelFile, synth_metagenome

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "mgcod_b200.h")).read()
    declared = set(re.findall(r"\b(mgc_[a-z_0-9]+)\s*\(", hdr))
    L = _lib.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"library lacks {missing}"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert b"mgcod_b200" in L.mgc_version()


def test_model_file_info_and_errors(tmp_path, model_dir):
    m = ModelFile(f"{model_dir}/mgm_4.mod")
    assert (m.gcode, m.cod_order, m.non_order, m.gene_min_length, m.n_bacterial, m.n_archaeal) == (4, 4, 2, 90, 41, 41)
    with pytest.raises(MgcodError) as e:
        ModelFile(str(tmp_path / "nope.mod"))
    assert e.value.code == _lib.MGC_E_IO
    bad = tmp_path / "bad.mod"
    bad.write_text("__B_GC_40\n$GCODE 11\n$COD_ORDER 1\n$NON_ORDER 0\n$NON_MAT\nA 0.25\nC 0.25\nG 0.25\nT 0.25\n")
    with pytest.raises(MgcodError) as e:
        ModelFile(str(bad))
    assert e.value.code == _lib.MGC_E_PARSE and "parameter not found" in str(e.value)
    bad.write_text("__B_GC_40\n$GCODE 7\n$GENE_MIN_LENGTH 90\n$COD_ORDER 0\n$NON_ORDER 0\n")
    with pytest.raises(MgcodError) as e:
        ModelFile(str(bad))
    assert "unsupported genetic code" in str(e.value)
    txt = open(f"{model_dir}/mgm_11.mod").read().replace("AAAAA ", "AAAA ", 1)
    bad.write_text(txt)
    with pytest.raises(MgcodError) as e:
        ModelFile(str(bad))
    assert "unexpected length" in str(e.value)


def test_engine_fails_loudly_without_gpu(model_dir):
    if _lib.lib().mgc_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(MgcodError) as e:
        Engine([f"{model_dir}/mgm_11.mod"])
    assert e.value.code == _lib.MGC_E_CUDA and "no CPU fallback" in str(e.value)


def test_synthetic_workload_is_deterministic():
    a, oa, ta = synth_metagenome(99, 200000, threads=1)
    b, ob, tb = synth_metagenome(99, 200000, threads=4)
    assert hashlib.sha256(a.tobytes()).hexdigest() == hashlib.sha256(b.tobytes()).hexdigest()
    assert (oa == ob).all() and (ta == tb).all()
    assert set(np.unique(a).tolist()) <= set(b"ACGT") and set(ta.tolist()) <= {11, 4, 15, 101}
    lens = np.diff(oa)
    assert lens.min() >= 5000 and lens.max() <= 500000 and oa[-1] >= 200000


def test_window_counts_rules():
    """window bookkeeping of bin/pipeline.py:194-275 (host part of mgc_window_labels)."""
    L = _lib.lib()
    lens = np.array([173747, 4000, 5000, 5001, 9000, 12000, 12001, 0], dtype=np.int64)
    stride = np.array([2500] * 8, dtype=np.int32)
    offs = np.zeros(9, dtype=np.int64)
    rc = L.mgc_window_counts(lens.ctypes.data_as(C.POINTER(C.c_int64)), stride.ctypes.data_as(C.POINTER(C.c_int32)), 8, 5000, 3, 0,
                             offs.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    # 173747/2500 -> 70 windows; <=5000 unclassified -> 1; 5001..10000 (w + 2*stride) -> ceil((L-w)/st)+1; longer -> ceil(L/st)
    assert np.diff(offs).tolist() == [70, 1, 1, 2, 3, 5, 5, 0]


def test_plane_index_layout():
    """the device tables are indexed (hi bits << n) | lo bits with the oldest letter at bit 0"""
    # natural index of "GAT" = 2*16 + 0*4 + 3 = 35 ; letters G(10) A(00) T(11): hi bits = 1,0,1 -> 0b101 ; lo = 0,0,1 -> 0b100
    nat, n = 35, 3
    hi = lo = 0
    for r in range(n):
        code = (nat >> (2 * (n - 1 - r))) & 3
        hi |= (code >> 1) << r
        lo |= (code & 1) << r
    assert (hi, lo) == (0b101, 0b100)


def test_native_fasta_reader_matches_python_reader(tmp_path):
    """mgc_fasta_read vs the plain Python reader: ids, deflines, sequences; CRLF, blank lines, spaces, lower case, an empty
    record, no trailing newline"""
    from mgcod_b200.data import Fasta
    from mgcod_b200.synth import read_fasta
    text = (">one first record\r\nACGTAC\r\nGGTT \r\n\r\n>two\n\nacgtnn\n  TTAA\n>empty\n>four|x  y\tz\nAC\nG")
    p = tmp_path / "x.fa"
    p.write_bytes(text.encode())
    f = Fasta(str(p), page_locked=False)
    want = read_fasta(str(p))
    assert f.ids == [w[0] for w in want] == ["one", "two", "empty", "four|x"]
    assert f.deflines == [w[1] for w in want]
    recs = f.records()
    assert [r.seq for r in recs] == [w[2] for w in want] == ["ACGTACGGTT", "acgtnnTTAA", "", "ACG"]
    assert [len(r) for r in recs] == [10, 10, 0, 3]
    assert f.offsets.tolist() == [0, 10, 20, 20, 23]
    # records cut their letters on first use: lengths are known at once, a record nobody asked for loses its letters with the
    # reader's buffer (Data.close detaches the records first), one that was read keeps its string
    g = Fasta(str(p), page_locked=False)
    lazy = g.records()
    assert [len(r) for r in lazy] == [10, 10, 0, 3] and all(r._seq is None for r in lazy)
    assert lazy[1].seq == "acgtnnTTAA"
    for r in lazy:
        r.detach()
    g.close()
    assert [r.seq for r in lazy] == ["", "acgtnnTTAA", "", ""] and [len(r) for r in lazy] == [10, 10, 0, 3]
    f.close()
    empty = tmp_path / "e.fa"
    empty.write_bytes(b"")
    assert Fasta(str(empty), page_locked=False).ids == []
    with pytest.raises(OSError):
        Fasta(str(tmp_path / "missing.fa"))


def test_gff_writer_formats_scores_like_printf(tmp_path):
    """the hand-rolled "%.2f" of the writer equals C printf on random and on boundary values (x.xx5, tiny, negative, huge)"""
    from mgcod_b200.engine import GENE_DTYPE, Result
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.uniform(0, 500, 60000), rng.integers(0, 100000, 20000) / 1000.0, rng.integers(0, 2000, 5000) / 200.0 + 0.005,
                           [0.0, 0.004999, 0.005, 0.015, 0.025, 1e-9, -0.001, -1.005, 1e15, 123456789.125, 2.675, 1.0049999999999999]])
    n = len(vals)
    g = np.zeros(n, dtype=GENE_DTYPE)
    g["contig"] = 0
    g["left"] = np.arange(n) * 3 + 1
    g["right"] = g["left"] + 299
    g["strand"] = 1
    g["logodd"] = vals
    g["gc"] = 40
    res = Result.from_genes([g], [np.array([1.5])])
    out = str(tmp_path / "f.gff")
    res.write(0, ["c"], [int(g["right"][-1]) + 10], "s", "m", 11, out, "gff")
    got = [l.split("\t")[5] for l in open(out) if "\tCDS\t" in l]
    assert got == ["%.2f" % v for v in vals]


def test_buffer_arguments_are_validated_before_crossing_the_abi():
    """wrong dtype, strided views and offsets that leave the buffer are refused in Python (no engine needed)"""
    chk = Engine._checked
    buf = np.frombuffer(b"ACGTACGTAC", dtype=np.uint8)
    assert chk(buf, [0, 4, 10]).tolist() == [0, 4, 10]
    with pytest.raises(TypeError):
        chk(buf.astype(np.int32), [0, 10])
    with pytest.raises(TypeError):
        chk(buf[::2], [0, 5])
    with pytest.raises(ValueError):
        chk(buf, [0, 11])
    with pytest.raises(ValueError):
        chk(buf, [0, 6, 4])
    with pytest.raises(ValueError):
        chk(buf, [3])
