"""Access to tests/golden (committed outputs of the reference; see tests/golden/make_golden.py)."""
from __future__ import annotations

import hashlib
import json
import lzma
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CODES = (11, 4, 15, 101)


def xz_text(name: str) -> str:
    with lzma.open(os.path.join(GOLDEN, name), "rt") as fh:
        return fh.read()


def write_xz(name: str, dst: str) -> str:
    with lzma.open(os.path.join(GOLDEN, name), "rb") as fh, open(dst, "wb") as out:
        out.write(fh.read())
    return dst


def parse_fasta_text(text: str):
    recs, cur = [], None
    for line in text.split("\n"):
        if line.startswith(">"):
            cur = [line[1:].split()[0] if line[1:].split() else "", []]
            recs.append(cur)
        elif cur is not None:
            cur[1].append(line.strip())
    return [(i, "".join(s)) for i, s in recs]


def phage2():
    return parse_fasta_text(xz_text("phage2.fa.xz"))


def trained_models(tmpdir: str) -> dict:
    return {c: write_xz(f"trained/mgm_{c}.mod.xz", os.path.join(tmpdir, f"mgm_{c}.mod")) for c in CODES}


def synthetic_case(tmpdir: str):
    """Regenerates the synthetic contigs + parametric models of the golden set and checks their hashes."""
    from mgcod_b200 import synth
    from mgcod_b200.engine import synth_metagenome
    meta = json.load(open(os.path.join(GOLDEN, "synth_parametric.json")))
    buf, offs, truth = synth_metagenome(meta["seed"], meta["total_bp"], threads=2)
    assert hashlib.sha256(buf.tobytes()).hexdigest() == meta["seq_sha256"], "synthetic generator drifted from the golden set"
    paths = synth.write_model_set(os.path.join(tmpdir, "par"), kind="parametric")
    for c in CODES:
        assert hashlib.sha256(open(paths[c], "rb").read()).hexdigest() == meta["model_sha256"][str(c)], \
            "parametric model generator drifted from the golden set"
    recs = [(f"contig_{i}", buf[offs[i]:offs[i + 1]].tobytes().decode()) for i in range(len(truth))]
    return recs, paths, truth


def gff_gene_lines(text: str):
    """[(contig, left, right, '%.2f', strand, attributes)] from GFF text."""
    out = []
    for line in text.split("\n"):
        if not line or line.startswith("#"):
            continue
        t = line.split("\t")
        out.append((t[0], int(t[3]), int(t[4]), t[5], t[6], t[8]))
    return out


def strip_dates(text: str) -> str:
    return "\n".join(l for l in text.split("\n") if not l.startswith("# output date"))


def genes_array_from_oracle(model, seqs):
    """Oracle decode of several contigs → (GENE_DTYPE array, totals)."""
    from mgcod_b200.engine import GENE_DTYPE
    rows, totals = [], []
    for ci, s in enumerate(seqs):
        genes, tot = model.decode(s)
        totals.append(tot)
        for g in genes:
            rows.append((ci, g["left"], g["right"], g["strand"], g["domain"], g["partial_left"], g["partial_right"], g["gc"], 0,
                         g["logodd"]))
    return np.array(rows, dtype=GENE_DTYPE), np.array(totals, dtype=np.float64)


def big_case(tmpdir: str):
    """The 24 Mbp / 82-section golden of tests/golden/make_golden_big.py: regenerates contigs + parameter files, checks their
    hashes → (meta, buffer, offsets, model paths)."""
    from mgcod_b200 import synth
    from mgcod_b200.engine import synth_metagenome
    with lzma.open(os.path.join(GOLDEN, "big_parametric.json.xz"), "rt") as fh:
        meta = json.load(fh)
    buf, offs, truth = synth_metagenome(meta["seed"], meta["total_bp"], threads=4)
    assert len(truth) == meta["n_contigs"]
    assert hashlib.sha256(buf.tobytes()).hexdigest() == meta["seq_sha256"], "synthetic generator drifted from the golden set"
    paths = synth.write_model_set(os.path.join(tmpdir, "bigpar"), kind="parametric")
    for c in CODES:
        assert hashlib.sha256(open(paths[c], "rb").read()).hexdigest() == meta["model_sha256"][str(c)]
    return meta, buf, offs, paths


def gene_attr_text(domain, pl, pr, gc, length) -> str:
    """attributes of a GFF gene line behind the gene_id (gmhmmp2 Output::PrintGenes_GFF)"""
    s = f"gene_type {'archaea' if domain else 'bacteria'};"
    if pl or pr:
        s += f" partial {int(pl)}{int(pr)};"
    return s + f" gc {int(gc)}; length {int(length)};"


def contig_digests(genes: np.ndarray, n_contigs: int):
    """per-contig digest of a GENE_DTYPE table in the form make_golden_big.py stores"""
    out = [hashlib.sha256() for _ in range(n_contigs)]
    for g in genes:
        left, right = int(g["left"]), int(g["right"])
        out[int(g["contig"])].update(("%d %d %s %s %s\n" % (
            left, right, "+" if g["strand"] > 0 else "-", "%.2f" % g["logodd"],
            gene_attr_text(g["domain"], g["partial_left"], g["partial_right"], g["gc"], right - left + 1))).encode())
    return [h.hexdigest()[:16] for h in out]
