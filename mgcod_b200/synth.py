"""Synthetic MetaGeneMark-style parameter files (``.mod``) for the four genetic codes.

The reference ships its parameters as ``dependencies/mgm_{11,4,15,101}.mod``
(consumed at reference ``bin/data.py:40-52``); those four files are absent from
the checkout (``.MISSING_LARGE_BLOBS``), so tests and the benchmark use
parameter sets written by this module in the grammar that GeneMark.hmm-2
accepts:

    __B_GC_<g> / __A_GC_<g>   section label (bacterial / archaeal, GC bin g)
    $KEY value                scalars
    $NON_MAT                  4^(NON_ORDER+1) lines  "KMER p"
    $COD_MAT                  4^(COD_ORDER+1) lines  "KMER p1 p2 p3"
                              (column c = k-mer whose first nt is codon position c)

Two ways to obtain tables:

* ``train_tables``      counts k-mers on a template sequence;
* ``parametric_tables`` closed-form tables from an i.i.d.-codon model at a GC
  level; needs no template, so it is what runs on a box without example data.

Everything here is deterministic: identical arguments give a byte-identical file.
"""
from __future__ import annotations

import hashlib
import os
from dataclasses import dataclass, field

import numpy as np

GCODES = (11, 4, 15, 101)
#: stop set implied by ``$GCODE``
STOP_CODONS = {
    11: ("TAA", "TAG", "TGA"),
    1: ("TAA", "TAG", "TGA"),
    4: ("TAA", "TAG"),
    25: ("TAA", "TAG"),
    15: ("TAA", "TGA"),
    101: ("TAG", "TGA"),
}
START_CODONS = ("ATG", "GTG", "TTG")
_NT = "ACGT"
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(_NT):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def encode(seq: str | bytes) -> np.ndarray:
    """ASCII → codes 0..3 (A,C,G,T); anything else → 4."""
    if isinstance(seq, str):
        seq = seq.encode()
    return _CODE[np.frombuffer(seq, dtype=np.uint8)]


def revcomp_codes(x: np.ndarray) -> np.ndarray:
    y = x[::-1].copy()
    ok = y < 4
    y[ok] = 3 - y[ok]
    return y


def kmer_labels(k: int) -> list[str]:
    out = [""]
    for _ in range(k):
        out = [p + c for p in out for c in _NT]
    return out


def _codon_index(c: str) -> int:
    return (_NT.index(c[0]) * 4 + _NT.index(c[1])) * 4 + _NT.index(c[2])


# ----------------------------------------------------------------------------------------
# tables
# ----------------------------------------------------------------------------------------
def _kmer_ids(x: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """ids of all k-mers of x (start positions 0..n-k) and a validity mask (no code 4 inside)."""
    n = x.shape[0] - k + 1
    if n <= 0:
        return np.zeros(0, np.int64), np.zeros(0, bool)
    ids = np.zeros(n, dtype=np.int64)
    bad = np.zeros(n, dtype=bool)
    for j in range(k):
        s = x[j:j + n]
        ids = ids * 4 + np.minimum(s, 3)
        bad |= s > 3
    return ids, ~bad


def train_tables(seq_codes: np.ndarray, gcode: int, cod_order: int, non_order: int,
                 min_orf: int = 900, pseudo: float = 1.0) -> tuple[np.ndarray, np.ndarray]:
    """Count frame-specific (k+1)-mers over long stop-free ORFs of both strands (coding) and
    all (k+1)-mers of both strands (non-coding); return joint probabilities
    ``cod[4^(kc+1), 3]``, ``non[4^(kn+1)]``."""
    kc, kn = cod_order + 1, non_order + 1
    cod = np.full((4 ** kc, 3), pseudo, dtype=np.float64)
    non = np.full(4 ** kn, pseudo, dtype=np.float64)
    stops = {_codon_index(c) for c in STOP_CODONS[gcode]}
    for strand in (seq_codes, revcomp_codes(seq_codes)):
        ids, ok = _kmer_ids(strand, kn)
        non += np.bincount(ids[ok], minlength=4 ** kn)
        cid, cok = _kmer_ids(strand, 3)
        is_stop = np.isin(cid, list(stops)) & cok
        L = strand.shape[0]
        for frame in range(3):
            pos = np.arange(frame, L - 2, 3)
            st = np.flatnonzero(is_stop[pos])
            bounds = np.concatenate([[-1], st, [pos.shape[0]]])
            for a, b in zip(bounds[:-1], bounds[1:]):
                ncod = b - a - 1
                if ncod * 3 < min_orf:
                    continue
                lo = pos[a + 1]
                hi = lo + ncod * 3  # exclusive
                sub = strand[lo:hi]
                kid, kok = _kmer_ids(sub, kc)
                ph = np.arange(kid.shape[0]) % 3
                for c in range(3):
                    m = kok & (ph == c)
                    cod[:, c] += np.bincount(kid[m], minlength=4 ** kc)
    cod /= cod.sum(axis=0, keepdims=True)
    non /= non.sum()
    return cod, non


def codon_usage(gc: float, gcode: int, skew: float = 0.35) -> np.ndarray:
    """A plausible codon distribution C[a,b,c] at GC level ``gc`` (fraction): position-specific
    base preferences, stops of the genetic code at ~0, reassigned stops at a low sense rate."""
    at, g_c = (1.0 - gc) / 2.0, gc / 2.0
    base = np.array([at, g_c, g_c, at])
    # codon-position preferences (purine-rich 1st position, weak 2nd, GC3 wobble follows GC strongly)
    p1 = base * np.array([1.15, 0.80, 1.45, 0.70])
    p2 = base * np.array([1.20, 0.95, 0.75, 1.10])
    w3 = np.array([1 - gc, gc, gc, 1 - gc]) ** (1.0 + 2.0 * skew)
    p3 = w3 * np.array([0.95, 1.05, 0.95, 1.05])
    p1, p2, p3 = p1 / p1.sum(), p2 / p2.sum(), p3 / p3.sum()
    C = p1[:, None, None] * p2[None, :, None] * p3[None, None, :]
    # deterministic codon-specific modulation so that higher orders carry information
    idx = np.arange(64).reshape(4, 4, 4)
    C = C * (1.0 + 0.45 * np.sin(1.7 * idx + 0.3) * np.cos(0.9 * idx))
    for c in ("TAA", "TAG", "TGA"):
        i = _codon_index(c)
        a, b, d = i // 16, (i // 4) % 4, i % 4
        if c in STOP_CODONS[gcode]:
            C[a, b, d] = 1e-7
        else:
            C[a, b, d] = 0.012  # reassigned stop read as a sense codon
    return C / C.sum()


def parametric_tables(gc: float, gcode: int, cod_order: int, non_order: int,
                      domain: str = "B") -> tuple[np.ndarray, np.ndarray]:
    """Closed-form joint tables for i.i.d. codons ``codon_usage(gc)`` (coding) and a first-order
    GC-skewed chain (non-coding)."""
    C = codon_usage(gc, gcode, skew=0.35 if domain == "B" else 0.20)
    kc, kn = cod_order + 1, non_order + 1
    digits = np.array(np.unravel_index(np.arange(4 ** kc), (4,) * kc)).T  # [n, kc]
    cod = np.zeros((4 ** kc, 3))
    for c in range(3):  # codon position (0-based) of the k-mer's first nt
        p = np.ones(4 ** kc)
        j = -c  # k-mer index where the first (possibly partial) codon starts
        while j < kc:
            sl = [slice(None)] * 3
            present = [(j + t) for t in range(3)]
            idxs = []
            marg = C
            # marginalise codon positions that fall outside the k-mer
            axes = tuple(t for t in range(3) if not (0 <= present[t] < kc))
            if axes:
                marg = C.sum(axis=axes)
            cols = [digits[:, present[t]] for t in range(3) if 0 <= present[t] < kc]
            p = p * marg[tuple(cols)]
            j += 3
        cod[:, c] = p
    cod /= cod.sum(axis=0, keepdims=True)
    # non-coding: order-1 chain with mild CpG depletion, expanded to the joint (kn)-mer table
    at, g_c = (1.0 - gc) / 2.0, gc / 2.0
    pi = np.array([at, g_c, g_c, at])
    T = np.tile(pi, (4, 1))
    T[1, 2] *= 0.7 if domain == "B" else 0.85
    T[3, 0] *= 1.1
    T /= T.sum(axis=1, keepdims=True)
    nd = np.array(np.unravel_index(np.arange(4 ** kn), (4,) * kn)).T
    non = pi[nd[:, 0]].copy()
    for j in range(1, kn):
        non = non * T[nd[:, j - 1], nd[:, j]]
    non /= non.sum()
    return cod, non


def tilt_tables(cod: np.ndarray, non: np.ndarray, w: float, mix: float = 0.0) -> tuple[np.ndarray, np.ndarray]:
    """Re-weight joint tables by ``w**(#G+#C in the k-mer)`` (moves the GC level of a trained
    table) and mix with uniform (``mix``); used to derive GC-bin / domain variants."""
    def gc_count(n):
        k = int(round(np.log(n) / np.log(4)))
        d = np.array(np.unravel_index(np.arange(n), (4,) * k)).T
        return ((d == 1) | (d == 2)).sum(axis=1)
    c2 = cod * (w ** gc_count(cod.shape[0]))[:, None]
    c2 /= c2.sum(axis=0, keepdims=True)
    c2 = (1 - mix) * c2 + mix / c2.shape[0]
    n2 = non * (w ** gc_count(non.shape[0]))
    n2 /= n2.sum()
    n2 = (1 - mix) * n2 + mix / n2.shape[0]
    return c2, n2


# ----------------------------------------------------------------------------------------
# file writer
# ----------------------------------------------------------------------------------------
@dataclass
class Section:
    label: str                      # e.g. "__B_GC_30"
    gcode: int
    cod_order: int
    non_order: int
    cod: np.ndarray                 # [4^(kc+1), 3]
    non: np.ndarray                 # [4^(kn+1)]
    gene_min_length: int = 90
    cod_p_n: float = 0.4
    non_p_n: float = 0.6
    non_duration_decay: float = 150.0
    cod_duration_decay: float = 300.0
    starts: dict = field(default_factory=lambda: {"ATG": 0.85, "GTG": 0.10, "TTG": 0.05})
    stops: dict | None = None       # default derived from gcode
    name: str = "synthetic"


_DEFAULT_STOP_P = {"TAA": 0.5, "TAG": 0.2, "TGA": 0.3}


def _fmt(p: float) -> str:
    return f"{p:.9e}"


def section_text(s: Section) -> str:
    stops = s.stops
    if stops is None:
        live = STOP_CODONS[s.gcode]
        tot = sum(_DEFAULT_STOP_P[c] for c in live)
        stops = {c: (_DEFAULT_STOP_P[c] / tot if c in live else 0.0) for c in _DEFAULT_STOP_P}
    out = [s.label,
           f"$NAME {s.name}",
           f"$GCODE {s.gcode}",
           f"$GENE_MIN_LENGTH {s.gene_min_length}",
           f"$COD_ORDER {s.cod_order}",
           f"$NON_ORDER {s.non_order}",
           f"$COD_P_N {s.cod_p_n:g}",
           f"$NON_P_N {s.non_p_n:g}",
           f"$NON_DURATION_DECAY {s.non_duration_decay:g}",
           f"$COD_DURATION_DECAY {s.cod_duration_decay:g}"]
    for c in START_CODONS:
        out.append(f"${c} {s.starts[c]:g}")
    for c in ("TAA", "TAG", "TGA"):
        out.append(f"${c} {stops[c]:g}")
    out.append("$NON_MAT")
    for lab, p in zip(kmer_labels(s.non_order + 1), s.non):
        out.append(f"{lab} {_fmt(p)}")
    out.append("$COD_MAT")
    for lab, row in zip(kmer_labels(s.cod_order + 1), s.cod):
        out.append(f"{lab} {_fmt(row[0])} {_fmt(row[1])} {_fmt(row[2])}")
    return "\n".join(out) + "\n"


def write_mod(path: str, sections: list[Section]) -> str:
    """Write ``sections`` to ``path``; returns the sha256 of the text."""
    text = "".join(section_text(s) for s in sections)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as fh:
        fh.write(text)
    return hashlib.sha256(text.encode()).hexdigest()


def parametric_sections(gcode: int, cod_order: int = 4, non_order: int = 2,
                        bins=range(30, 71), domains=("B", "A"), **kw) -> list[Section]:
    """The benchmark's model set: one section per (domain, GC bin); 41 + 41 by default
    (SURVEY.md §8d "Model sets for timing")."""
    secs = []
    for d in domains:
        for g in bins:
            cod, non = parametric_tables(g / 100.0, gcode, cod_order, non_order, d)
            secs.append(Section(f"__{d}_GC_{g}", gcode, cod_order, non_order, cod, non,
                                name=f"synth_{d}_{g}_gc{gcode}", **kw))
    return secs


def trained_sections(seq_codes: np.ndarray, gcode: int, cod_order: int = 4, non_order: int = 2,
                     bins=None, domains=("B",), center_gc: float | None = None, **kw) -> list[Section]:
    """Sections trained on a template; with ``bins`` the trained tables are GC-tilted per bin and
    the archaeal variant is additionally mixed with uniform so that B and A differ."""
    cod, non = train_tables(seq_codes, gcode, cod_order, non_order)
    if bins is None:
        return [Section("__B_GC_50", gcode, cod_order, non_order, cod, non, **kw)]
    if center_gc is None:
        ok = seq_codes < 4
        center_gc = float(((seq_codes == 1) | (seq_codes == 2)).sum() / max(1, ok.sum()))
    secs = []
    for d in domains:
        for g in bins:
            gf = min(max(g / 100.0, 0.02), 0.98)
            w = ((gf / (1 - gf)) / (center_gc / (1 - center_gc))) ** 0.5
            c2, n2 = tilt_tables(cod, non, w, mix=0.0 if d == "B" else 0.08)
            secs.append(Section(f"__{d}_GC_{g}", gcode, cod_order, non_order, c2, n2, **kw))
    return secs


def write_model_set(out_dir: str, kind: str = "parametric", template: np.ndarray | None = None,
                    gcodes=GCODES, **kw) -> dict[int, str]:
    """Write ``mgm_<code>.mod`` for every code into ``out_dir``; returns {code: path}."""
    paths = {}
    for gc in gcodes:
        secs = (parametric_sections(gc, **kw) if kind == "parametric"
                else trained_sections(template, gc, **kw))
        p = os.path.join(out_dir, f"mgm_{gc}.mod")
        write_mod(p, secs)
        paths[gc] = p
    return paths


def read_fasta(path: str) -> list[tuple[str, str, str]]:
    """Minimal FASTA reader → [(id, defline, sequence)]; id = first whitespace-delimited token
    (reference ``bin/data.py:117`` relies on Biopython's ``rec.id`` which is the same token)."""
    recs, name, desc, chunks = [], None, None, []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, desc, "".join(chunks)))
                desc = line[1:].rstrip("\r\n")
                name = desc.split()[0] if desc.split() else ""
                chunks = []
            else:
                chunks.append(line.strip())
    if name is not None:
        recs.append((name, desc, "".join(chunks)))
    return recs
