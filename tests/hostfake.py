"""Host stand-in for the CUDA engine's aggregation calls — TEST INFRASTRUCTURE (CPU tests and tests/tools/host_profile.py).

Restates in numpy what `k_windows` / `mgc_window_counts` (reference bin/pipeline.py:157-385), `k_merge_labels`
(bin/pipeline.py:1064-1168, through mgcod_b200.labels) and `k_recoded_stops` (bin/results.py:253-294) compute, with the
engine's return layouts, so that the host layer above the C-ABI (Data / Pipeline / Results) can be exercised and profiled
in the build container, which has no GPU.  Never imported by the package.
"""
from __future__ import annotations

import numpy as np

from mgcod_b200 import labels as L
from mgcod_b200.engine import Engine, Result


def window_counts(lens, stride, window, consecutive, short_contigs):
    offs = np.zeros(len(lens) + 1, dtype=np.int64)
    for c, (n, st) in enumerate(zip(lens, stride)):
        n, st = int(n), int(st)
        if n <= 0:
            nw = 0
        elif n <= 5000 and not short_contigs:
            nw = 1
        elif n <= window + (consecutive - 1) * st:
            nw = (n - window + st - 1) // st + 1 if n > window else 1
        else:
            nw = (n + st - 1) // st
        offs[c + 1] = offs[c] + nw
    return offs


class HostEngine:
    MAX_BATCH_BP = Engine.MAX_BATCH_BP

    def __init__(self, genes_per_model, seqs=None):
        """genes_per_model[m]: GENE_DTYPE array ordered by (contig, left); seqs: contig strings in batch order (only for
        recoded_stops)"""
        self.genes = genes_per_model
        self.seqs = seqs
        self.n_models = len(genes_per_model)
        self._resident_owner = None

    def window_labels(self, contig_len, stride, window, consecutive, short_contigs):
        lens = np.asarray(contig_len, dtype=np.int64)
        st = np.asarray(stride, dtype=np.int64)
        offs = window_counts(lens, st, window, consecutive, short_contigs)
        nw = int(offs[-1])
        labels = np.zeros(nw, dtype=np.int32)
        cents = np.zeros((nw, self.n_models), dtype=np.int64)
        for m, g in enumerate(self.genes):
            gc = np.rint(g["logodd"] * 100.0).astype(np.int64)
            first = np.searchsorted(g["contig"], np.arange(len(lens) + 1))
            for c in range(len(lens)):
                a, b = int(first[c]), int(first[c + 1])
                n, s = int(lens[c]), int(st[c])
                w0, w1 = int(offs[c]), int(offs[c + 1])
                if n <= 5000 and not short_contigs:
                    continue
                if n <= window + (consecutive - 1) * s:
                    cents[w0:w1, m] = gc[a:b].sum()
                    continue
                left, right = g["left"][a:b].astype(np.int64), g["right"][a:b].astype(np.int64)
                pre = np.concatenate([[0], np.cumsum(gc[a:b])])
                starts = np.arange(w1 - w0, dtype=np.int64) * s
                lo = np.searchsorted(left, starts, side="left")
                hi = np.searchsorted(left, starts + window, side="left")
                sums = pre[hi] - pre[lo]
                prev = np.maximum(lo - 1, 0)
                span = (sums == 0) & (lo > 0) & (right[prev] > starts + window) if b > a else np.zeros(w1 - w0, bool)
                sums = np.where(span, gc[a:b][prev], sums) if b > a else sums
                cents[w0:w1, m] = sums
        if nw:
            labels[:] = np.argmax(cents, axis=1)
        for c in range(len(lens)):
            if int(lens[c]) <= 5000 and not short_contigs:
                labels[offs[c]:offs[c + 1]] = 0
        return offs, labels, cents

    def merge_labels_flat(self, flat, offs, consecutive):
        flat = np.asarray(flat, dtype=np.int32)
        out = flat.copy()
        status = np.zeros(len(offs) - 1, dtype=np.int32)
        for c in range(len(offs) - 1):
            l = flat[offs[c]:offs[c + 1]]
            if np.unique(l).shape[0] <= 1:
                continue
            try:
                out[offs[c]:offs[c + 1]] = L.find_best_sequence_of_labels(l, flat, consecutive)
            except IndexError:
                status[c] = 1
        return out, status

    merge_labels = Engine.merge_labels

    def isoform_labels(self, contig_len, stride, window, consecutive, short_contigs):
        offs, labels, cents = self.window_labels(contig_len, stride, window, consecutive, short_contigs)
        merged, status = self.merge_labels_flat(labels, offs, consecutive)
        return offs, labels, cents, merged, status

    def recoded_stops(self, genes):
        from mgcod_b200.results import Results, _COMP
        counts, offs, pos = np.zeros((len(genes), 3), dtype=np.int32), [0], []
        for g in range(len(genes)):
            s = self.seqs[int(genes["contig"][g])][int(genes["left"][g]) - 1:int(genes["right"][g])]
            if genes["strand"][g] < 0:
                s = s[::-1].translate(_COMP)
            s = s.lower()
            for t, stop in enumerate(("taa", "tag", "tga")):
                hits, _ = Results.find_recoded_stops(s, stop)
                counts[g, t] = len(hits)
                pos.extend(hits)
                offs.append(len(pos))
        return counts, np.array(offs, dtype=np.int64), np.array(pos, dtype=np.int32)


def oracle_result(model_paths, seqs, threads=8):
    """Decode ``seqs`` with the CPU oracle under every parameter file → (Result, genes per model)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import gmhmm
    from tests import fixtures as fx
    genes, totals = [], []
    for p in model_paths:
        model = gmhmm.Model(p)
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda s: fx.genes_array_from_oracle(model, [s]), seqs))
        for c, (g, _) in enumerate(parts):
            g["contig"] = c
        genes.append(np.concatenate([g for g, _ in parts]))
        totals.append(np.concatenate([t for _, t in parts]))
    return Result.from_genes(genes, totals), genes
