"""Multi-GPU sharding of the hot path: contigs are independent units (the reference parallelises over
FASTA files with a process pool, bin/multiprocess_mgcod.py:140-145), so ranks take disjoint sets of contigs
and the only exchange is the final gather of the (small) gene tables.  No data-path collective.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def partition_contigs(lengths: Sequence[int], n_ranks: int) -> list[np.ndarray]:
    """Greedy longest-first bin packing by bp: returns, per rank, the (ascending) contig indices it owns.
    All models of a contig stay on the same GPU, so each sequence is read from host memory once."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(n_ranks, dtype=np.int64)
    owner = np.zeros(lengths.shape[0], dtype=np.int64)
    for c in order:
        r = int(np.argmin(load))
        owner[c] = r
        load[r] += lengths[c]
    return [np.flatnonzero(owner == r) for r in range(n_ranks)]


def merge_rank_genes(per_rank_genes: Sequence[np.ndarray], per_rank_contigs: Sequence[np.ndarray]) -> np.ndarray:
    """Rank-local gene tables (contig index local to the rank's shard) → one table in global contig order."""
    parts = []
    for genes, contigs in zip(per_rank_genes, per_rank_contigs):
        g = genes.copy()
        if g.shape[0]:
            g["contig"] = np.asarray(contigs)[g["contig"]]
        parts.append(g)
    allg = np.concatenate(parts) if parts else np.zeros(0)
    return allg[np.lexsort((allg["left"], allg["contig"]))] if allg.shape[0] else allg


def gather_genes(local_genes: np.ndarray, local_contigs: np.ndarray, group=None) -> np.ndarray | None:
    """torch.distributed gather of the rank-local results on rank 0 (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    payload = (local_genes.tobytes(), np.asarray(local_contigs, dtype=np.int64).tobytes(), local_genes.dtype.descr)
    gathered = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if gathered is None:
        return None
    genes = [np.frombuffer(g, dtype=np.dtype(d)) for g, _, d in gathered]
    contigs = [np.frombuffer(c, dtype=np.int64) for _, c, _ in gathered]
    return merge_rank_genes(genes, contigs)
