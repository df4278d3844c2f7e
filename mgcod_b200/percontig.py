"""Per-contig result containers that cost nothing for the contigs nothing happens on.

The reference's Pipeline exposes its results as Python lists and dicts with one entry per contig (``genome_label``,
``updated_genome_labels``, ``logodd_scores``; ``gene_label_to_gcode``, ``confident_1/2``, ``segment_labels``,
``predicted_switch_regions``, bin/pipeline.py:56-137).  For a metagenome of 40 000 contigs, of which a handful carry a
genetic-code switch, building those entries one by one was most of the command line's time.  These two classes keep the API —
indexing, iteration, ``items()``, ``len()`` — over flat arrays: an entry is made when somebody asks for it.
"""
from __future__ import annotations

from collections.abc import Mapping, Sequence

import numpy as np


class Ragged(Sequence):
    """list of arrays = slices ``flat[lo[k]:hi[k]]`` of one flat array (a list of per-contig label / score arrays)"""

    def __init__(self, flat: np.ndarray, lo, hi):
        self.flat = flat
        self.lo = np.asarray(lo, dtype=np.int64)
        self.hi = np.asarray(hi, dtype=np.int64)

    def __len__(self):
        return self.lo.shape[0]

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        return self.flat[self.lo[k]:self.hi[k]]

    def __iter__(self):
        flat = self.flat
        for a, b in zip(self.lo.tolist(), self.hi.tolist()):
            yield flat[a:b]

    def concatenated(self) -> np.ndarray:
        """all entries in order, as one array"""
        if len(self) and bool((self.lo[1:] == self.hi[:-1]).all()):
            return self.flat[self.lo[0]:self.hi[-1]]
        return np.concatenate(list(self)) if len(self) else self.flat[:0]


class PerContig(Mapping):
    """contig id -> value: explicit entries for the few contigs that need one, ``rule(n)`` (n = position of the contig in
    ``ids``) for all others.  Behaves like the reference's dict, in ``ids`` order."""

    def __init__(self, ids, explicit: dict, rule):
        self._ids = ids
        self._explicit = explicit
        self._rule = rule
        self._pos = None

    def _position(self, key):
        if self._pos is None:
            self._pos = {c: n for n, c in enumerate(self._ids)}
        return self._pos[key]

    def __getitem__(self, key):
        if key in self._explicit:
            return self._explicit[key]
        return self._rule(self._position(key))

    def at(self, n: int):
        """value of the n-th contig (no dictionary of positions needed)"""
        key = self._ids[n]
        return self._explicit[key] if key in self._explicit else self._rule(n)

    def __iter__(self):
        return iter(self._ids)

    def __len__(self):
        return len(self._ids)

    def __contains__(self, key):
        if key in self._explicit:
            return True
        try:
            self._position(key)
            return True
        except KeyError:
            return False

    def items(self):
        ex, rule = self._explicit, self._rule
        return [(c, ex[c] if c in ex else rule(n)) for n, c in enumerate(self._ids)]

    def values(self):
        return [v for _, v in self.items()]

    def explicit_keys(self):
        return self._explicit.keys()
