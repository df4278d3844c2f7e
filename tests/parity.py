"""Shared comparison helpers: CUDA engine (through the C-ABI) versus the CPU oracle."""
from __future__ import annotations

import numpy as np


def oracle_genes(model, seqs):
    """[(contig, left, right, strand, cents, domain, gc, pl, pr)] in (contig, left) order + per-contig totals."""
    rows, totals = [], []
    for ci, s in enumerate(seqs):
        genes, tot = model.decode(s)
        totals.append(tot)
        for g in genes:
            rows.append((ci, g["left"], g["right"], g["strand"], g["logodd"], g["domain"], g["gc"], g["partial_left"],
                         g["partial_right"]))
    return rows, totals


def compare_genes(gpu: np.ndarray, ref_rows, tol=1e-9):
    """Returns a list of human-readable differences (empty = parity)."""
    diffs = []
    a = [(int(g["contig"]), int(g["left"]), int(g["right"]), int(g["strand"])) for g in gpu]
    b = [(r[0], r[1], r[2], r[3]) for r in ref_rows]
    if a != b:
        sa, sb = set(a), set(b)
        diffs.append(f"gene sets differ: {len(a)} vs {len(b)}; only gpu {sorted(sa - sb)[:5]} only oracle {sorted(sb - sa)[:5]}")
        return diffs
    for g, r in zip(gpu, ref_rows):
        if abs(g["logodd"] - r[4]) > tol * max(1.0, abs(r[4])):
            diffs.append(f"logodd {tuple(r[:4])}: {g['logodd']} vs {r[4]}")
        if (int(g["domain"]), int(g["gc"]), int(g["partial_left"]), int(g["partial_right"])) != (r[5], r[6], r[7], r[8]):
            diffs.append(f"attributes {tuple(r[:4])}: gpu {(g['domain'], g['gc'], g['partial_left'], g['partial_right'])} vs {r[5:]}")
        if len(diffs) > 10:
            break
    return diffs


def compare_candidates(gpu: np.ndarray, model, seqs, tol=1e-9):
    """Candidate sets (total > 0) with totals and chain scores."""
    diffs = []
    for ci, s in enumerate(seqs):
        ref, _, _ = model.candidates(s)
        refd = {(c["left"], c["right"], c["strand"], c["partial_start"]): c for c in ref}
        mine = gpu[gpu["contig"] == ci]
        gd = {(int(c["left"]), int(c["right"]), int(c["strand"]), int(c["partial_start"])): c for c in mine}
        om, orf = sorted(set(gd) - set(refd)), sorted(set(refd) - set(gd))
        if om or orf:
            diffs.append(f"contig {ci}: candidates only gpu {om[:4]} ({len(om)}), only oracle {orf[:4]} ({len(orf)})")
        for k in set(gd) & set(refd):
            g, r = gd[k], refd[k]
            if abs(g["total"] - r["total"]) > tol * max(1.0, abs(r["total"])):
                diffs.append(f"contig {ci} cand {k}: total {g['total']} vs {r['total']}")
            elif abs(g["cum"] - r["cum"]) > tol * max(1.0, abs(r["cum"])):
                diffs.append(f"contig {ci} cand {k}: cum {g['cum']} vs {r['cum']}")
            elif (int(g["gc"]), int(g["domain"]), int(g["partial_stop"])) != (r["gc"], r["domain"], r["partial_stop"]):
                diffs.append(f"contig {ci} cand {k}: attrs")
            if len(diffs) > 12:
                return diffs
    return diffs
