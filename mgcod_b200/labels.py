"""Label smoothing and segmentation of window labels — host logic of the hot path's last stage.

Restates two methods of the reference on plain numpy arrays (no DataFrames):

* :func:`find_best_sequence_of_labels`   reference ``Pipeline.find_best_sequence_of_labels``
  (bin/pipeline.py:1064-1168): seed-and-extend merging of runs of equal window labels.
* :func:`segment_contig`                 reference ``Pipeline.segment_genome_based_on_window_labels``
  (bin/pipeline.py:387-492) for one contig: switch coordinates between blocks of different genetic code.

Both are tiny, strictly sequential per contig and integer-valued; they run on the host on the label
arrays the GPU produced (DESIGN.md §"Aggregation").
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def resolve_stride(stride: int | None, length: int) -> int:
    """Reference default: 2500 bp up to 100 kb, else 5000 bp (bin/pipeline.py:172-183)."""
    if stride is not None:
        return int(stride)
    return 2500 if length <= 100000 else 5000


def noise_labels(all_labels: np.ndarray) -> np.ndarray:
    """Labels other than the two most frequent ones over the whole input (bin/pipeline.py:1084-1086).
    Frequency ties are resolved the way ``np.argsort(counts)[::-1]`` resolves them."""
    uniq, counts = np.unique(all_labels, return_counts=True)
    return uniq[np.argsort(counts)[::-1]][2:]


def find_best_sequence_of_labels(labels: Sequence[int], genome_label: Sequence[int] | None, consecutive_labels: int,
                                 allow_multiple_labels: bool = False) -> np.ndarray:
    """Merge short runs into their neighbours.

    Every position carries the length of the run of equal labels ending there (noise labels extend the
    run they follow).  Repeatedly: take the position with the largest remaining run length as seed, take
    its label as reference, extend the block to the left and to the right until ``consecutive_labels``
    consecutive positions disagree with the reference, and overwrite the block with the reference.  A
    block that ends up shorter than ``consecutive_labels`` is given to the longer of its two neighbouring
    runs (right wins ties) or to their common label.
    """
    labels = np.asarray(labels)
    n = labels.shape[0]
    out = labels.copy()
    if n == 0:
        return out
    noise = [] if allow_multiple_labels else list(noise_labels(np.asarray(genome_label)))
    run = np.zeros(n, dtype=np.int64)
    run[0] = 1
    for i in range(1, n):
        run[i] = run[i - 1] + 1 if (labels[i] == labels[i - 1] or labels[i] in noise) else 1
    k = int(consecutive_labels)
    while run.any():
        seed = int(np.argmax(run))
        lo = seed - int(run[seed])     # one position left of the run
        hi = seed + 1                  # one position right of the run
        ref_pos = seed
        ref = out[ref_pos]
        while ref in noise:            # a noise label never names a block
            ref_pos -= 1
            ref = out[ref_pos]
        miss = 0
        while lo >= 0 and miss < k:
            miss = miss + 1 if out[lo] != ref else 0
            lo -= 1
        miss_left_exhausted = lo < 0
        miss = 0
        while hi <= n - 1 and miss < k:
            miss = miss + 1 if out[hi] != ref else 0
            hi += 1
        lo = 0 if miss_left_exhausted else lo + k + 1   # drop the k mismatches (lo itself is one step beyond)
        hi = n if hi > n - 1 else hi - k
        if hi - lo < k:
            left_label = out[lo - 1]                    # note: index -1 wraps to the last element, as in the reference
            right_label = out[hi] if hi < n else left_label
            if left_label == right_label:
                ref = left_label
            else:
                i, left_len = lo - 1, 0
                while i >= 0 and out[i] == left_label:
                    i -= 1
                    left_len += 1
                j, right_len = hi, 0
                while j < n and out[j] == right_label:
                    j += 1
                    right_len += 1
                ref = left_label if left_len > right_len else right_label
        run[lo:hi] = 0
        out[lo:hi] = ref
    return out


def segment_contig(updated_labels: np.ndarray, contig_len: int, stride: int, window: int, genes_by_label: dict,
                   label_to_gcode: dict) -> tuple[list, list, list]:
    """Switch coordinates for one contig.

    ``genes_by_label[label]`` = (left[], right[], strand[]) of that model's genes on this contig (GFF
    coordinates, strand as +1/-1, ordered as in the GFF).  Returns (confident_1, confident_2, codes):
    the blocks (start, end) carrying the smaller / larger label value and the genetic code per block.
    """
    labs = np.asarray(updated_labels)
    uniq = np.unique(labs)
    label_1 = uniq[0]
    label_2 = uniq[1] if uniq.shape[0] > 1 else None
    conf_1, conf_2, codes = [], [], []
    ref = labs[0]
    seg_start = 1

    def close(seg_end):
        if ref == label_1:
            conf_1.append((seg_start, seg_end))
            codes.append(label_to_gcode[int(label_1)])
        else:
            conf_2.append((seg_start, seg_end))
            codes.append(label_to_gcode[int(label_2)])

    for i in range(1, labs.shape[0]):
        if labs[i] == ref:
            continue
        start = (i - 1) * stride
        end = i * stride + window
        o_left, o_right, o_strand = genes_by_label[int(labs[i - 1])]
        n_left, _, _ = genes_by_label[int(labs[i])]
        inwin = (o_left >= start) & (o_left < end)
        strands = o_strand[inwin]
        ends = o_right[inwin].astype(np.float64)
        if ends.shape[0] == 0:
            ends = np.array([(start + end) / 2])
        if np.unique(strands).shape[0] > 1:
            # the coding strand flips inside the two windows: the switch sits at the last gene before the flip
            flip = np.where(np.diff(np.where(strands > 0, 1, 0)) != 0)[0][0]
            sp1 = ends[flip]
        else:
            cand = o_right[o_left <= ends[0]]
            sp1 = cand[-1] if cand.shape[0] else ends[0]
        nxt = n_left[n_left >= sp1]
        sp2 = nxt[0] if nxt.shape[0] else sp1 + 2
        seg_end = int((sp1 + sp2) / 2)
        close(seg_end)
        seg_start = seg_end + 1
        ref = labs[i]
    close(int(contig_len))
    return conf_1, conf_2, codes
