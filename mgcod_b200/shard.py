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


def partition_ranges(lengths: Sequence[int], n_ranks: int) -> np.ndarray:
    """Contiguous ranges of contigs of about equal bp: → cut[n_ranks + 1]; rank r owns contigs cut[r] .. cut[r+1]-1.
    What the GPU path uses: a rank's contigs are one slice of the host letter buffer (``offsets[cut[r]:cut[r+1]+1]`` goes to
    the engine as it is, nothing is copied together), the shares differ by at most one contig, and the rank-local gene tables
    concatenated in rank order are already in global (contig, left) order."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    ends = np.cumsum(lengths)
    total = int(ends[-1]) if n else 0
    cut = np.zeros(n_ranks + 1, dtype=np.int64)
    for r in range(1, n_ranks):
        target = total * r / n_ranks
        c = int(np.searchsorted(ends, target, side="left"))          # ends[c] >= target > ends[c - 1]
        if c < n and (c == 0 or ends[c] - target <= target - ends[c - 1]):
            c += 1                                                    # the cut behind contig c is the nearer one
        cut[r] = max(min(c, n), cut[r - 1])
    cut[n_ranks] = n
    return cut


def gather_genes_device(engine, first_contig: int, group=None, dst: int = 0):
    """Final gather of a sharded decode, straight from the engines' device memory: every rank contributes the gene tables of
    its last decode (all models, 32-byte rows), rank ``dst`` returns them merged per model in global contig order (other ranks
    return None).  ``first_contig`` = global index of the rank's first contig (``partition_ranges``).
    Replaces the result collection of the reference's process pool (bin/multiprocess_mgcod.py:140-145).

    Ranks of ONE node (the usual case: torchrun sets LOCAL_WORLD_SIZE == WORLD_SIZE) write their rows into one host segment
    that all of them map (:func:`gather_gene_tables_shared`): N device-to-host copies over N PCIe links side by side, where
    the NCCL gather to one GPU (:func:`gather_gene_tables`; several nodes, or ``MGCOD_B200_GATHER=nccl``) ends in one copy of
    everything over rank ``dst``'s link."""
    import os
    import torch
    tables = [engine.device_genes(m) for m in range(engine.n_models)]
    device = torch.device("cuda", torch.cuda.current_device())
    one_node = os.environ.get("LOCAL_WORLD_SIZE", "1") == os.environ.get("WORLD_SIZE", "1")
    if one_node and os.environ.get("MGCOD_B200_GATHER", "shared") != "nccl":
        return gather_gene_tables_shared(tables, first_contig, device, group, dst)
    return gather_gene_tables(tables, first_contig, device, group, dst)


_SHARED = {}      # the host segment of this process: capacity, generation, mapping, tensor view


def _shared_segment(total: int, rank: int, dst: int, group, cuda: bool):
    """A host memory segment of at least ``total`` bytes mapped by every rank of the node (a file under /dev/shm, created by
    rank ``dst``, unlinked as soon as everybody has mapped it) and, with CUDA, page-locked in every process.  All ranks call
    this with the same ``total``, so they agree on capacity and generation without talking."""
    import mmap
    import os
    import torch
    import torch.distributed as dist
    st = _SHARED
    if st.get("cap", 0) >= total and st.get("tensor") is not None:
        return st["tensor"]
    if st.get("tensor") is not None:
        if st["registered"]:
            torch.cuda.cudart().cudaHostUnregister(st["tensor"].data_ptr())
        st["tensor"] = None
        st["map"] = None          # unmapped when the last view dies
    cap = (int(total * 1.25) + (1 << 20) + 4095) // 4096 * 4096
    gen = st.get("gen", 0) + 1
    st.update(cap=0, gen=gen)
    path = f"/dev/shm/mgcod_b200_{os.getuid()}_{os.environ.get('MASTER_PORT', '0')}_{gen}"
    ok = torch.zeros(1, dtype=torch.int64, device="cuda" if cuda else "cpu")
    if rank == dst:
        # a container's /dev/shm may be 64 MB: writing past its free space is a SIGBUS, not an error code
        try:
            fs = os.statvfs("/dev/shm")
            if fs.f_bavail * fs.f_frsize >= cap + (64 << 20):
                fd = os.open(path, os.O_CREAT | os.O_TRUNC | os.O_RDWR, 0o600)
                os.ftruncate(fd, cap)
                ok += 1
        except OSError:
            pass
    dist.broadcast(ok, src=dst, group=group)      # also: the file exists before the others open it
    if int(ok.item()) == 0:
        return None                               # the caller falls back to the collective
    if rank != dst:
        fd = os.open(path, os.O_RDWR)
    mm = mmap.mmap(fd, cap)
    os.close(fd)
    dist.barrier(group)
    if rank == dst:
        os.unlink(path)           # the mappings keep the memory; nothing is left behind if a rank dies
    t = torch.frombuffer(mm, dtype=torch.uint8)
    registered = False
    if cuda:
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), cap, 0)
        registered = int(rc) == 0        # not page-locked: the copies still work, through the driver's staging buffer
    st.update(cap=cap, gen=gen, map=mm, tensor=t, registered=registered)
    return t


def gather_gene_tables_shared(tables, first_contig: int, device, group=None, dst: int = 0):
    """Gather through host memory shared by the ranks of one node.  ``tables`` as for :func:`gather_gene_tables`.  Ranks
    exchange their per-model counts (one small all_gather: also the point after which rank ``dst`` no longer reads the arrays
    of the previous call), every rank copies its rows of model m — contig field already global — to
    ``offset(m) + rows of lower ranks`` of the segment, a barrier closes the call.  The arrays rank ``dst`` returns are views of
    the segment, valid until the next call."""
    import torch
    import torch.distributed as dist
    from .engine import GENE_DTYPE
    import os
    import time
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    M = len(tables)
    cuda = device.type == "cuda"
    trace = os.environ.get("MGCOD_TRACE") and rank == dst
    t0 = time.perf_counter()
    counts = torch.tensor([len(t) for t in tables], dtype=torch.int64, device=device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = torch.stack(all_counts).cpu().numpy()            # [world, M]
    per_model = all_counts.sum(axis=0)
    total = int(per_model.sum()) * 32
    t1 = time.perf_counter()
    seg = _shared_segment(max(total, 32), rank, dst, group, cuda)
    if seg is None:       # no room under /dev/shm
        return gather_gene_tables(tables, first_contig, device, group, dst)
    t2 = time.perf_counter()
    base = np.concatenate(([0], np.cumsum(per_model)))[:M] * 32
    for m, t in enumerate(tables):
        n = len(t) * 32
        if not n:
            continue
        at = int(base[m]) + int(all_counts[:rank, m].sum()) * 32
        if isinstance(t, np.ndarray):
            rows = np.ascontiguousarray(t, dtype=GENE_DTYPE).copy()
            rows["contig"] += int(first_contig)
            seg[at:at + n].copy_(torch.from_numpy(rows.view(np.uint8)))
        else:
            rows = torch.as_tensor(t, device=device)
            if first_contig:
                rows = rows.clone()
                rows.view(torch.int32).view(-1, 8)[:, 0] += int(first_contig)    # contig = first field of the 32-byte row
            seg[at:at + n].copy_(rows, non_blocking=True)
    if cuda:
        torch.cuda.current_stream().synchronize()
    t3 = time.perf_counter()
    dist.barrier(group)
    if trace:
        import sys
        print(f"gather (shared segment, registered {_SHARED.get('registered')}): counts {1e3 * (t1 - t0):.2f} ms, segment {1e3 * (t2 - t1):.2f} ms, "
              f"copies {1e3 * (t3 - t2):.2f} ms, barrier {1e3 * (time.perf_counter() - t3):.2f} ms", file=sys.stderr)
    if rank != dst:
        return None
    host = seg.numpy()
    return [host[int(base[m]):int(base[m]) + int(per_model[m]) * 32].view(GENE_DTYPE) for m in range(M)]


_PINNED = {}      # device index -> page-locked receive buffer of rank dst (grown on demand, kept between calls)


def _pinned_bytes(device, n: int):
    import torch
    key = str(device)
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < n:
        buf = torch.empty(int(n * 1.25) + 4096, dtype=torch.uint8, pin_memory=device.type == "cuda")
        _PINNED[key] = buf
    return buf


def gather_gene_tables(tables, first_contig: int, device, group=None, dst: int = 0):
    """The collective under :func:`gather_genes_device`.  ``tables``: one gene table per model, each either a
    :class:`mgcod_b200.engine.DeviceGenes` (device memory, NCCL) or a GENE_DTYPE numpy array (host memory; the gloo tests).

    On the wire: per rank one uint8 tensor holding model 0's rows, then model 1's, ... with the rank's first contig already
    added to the rows' contig field (a strided int32 add on the send buffer); ranks exchange their per-model counts first
    (one small all_gather) and pad to the largest total, because a gather takes equal sizes.  Rank ``dst`` then strings the
    ranks' pieces of each model together ON THE DEVICE and brings every model's table to the host with one copy into a
    page-locked buffer; the returned arrays are views of it (valid until the next call)."""
    import torch
    import torch.distributed as dist
    from .engine import GENE_DTYPE
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    M = len(tables)
    counts = torch.tensor([len(t) for t in tables], dtype=torch.int64, device=device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = torch.stack(all_counts).cpu().numpy()            # [world, M]
    width = int(all_counts.sum(axis=1).max()) * 32
    send = torch.empty(max(width, 32), dtype=torch.uint8, device=device)
    at = 0
    for t in tables:
        if len(t):
            src = (torch.from_numpy(np.ascontiguousarray(t, dtype=GENE_DTYPE).view(np.uint8).copy()) if isinstance(t, np.ndarray)
                   else torch.as_tensor(t, device=device))
            send[at:at + len(t) * 32].copy_(src)
        at += len(t) * 32
    if at and first_contig:
        send[:at].view(torch.int32).view(-1, 8)[:, 0] += int(first_contig)    # contig = first field of the 32-byte row
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    total = int(all_counts.sum()) * 32
    host = _pinned_bytes(device, max(total, 32))
    merged, at = [], 0
    for m in range(M):
        pieces = []
        for r in range(world):
            a = int(all_counts[r, :m].sum()) * 32
            pieces.append(recv[r][a:a + int(all_counts[r, m]) * 32])
        n = int(all_counts[:, m].sum()) * 32
        if n:
            host[at:at + n].copy_(torch.cat(pieces), non_blocking=True)
        merged.append(host[at:at + n].numpy().view(GENE_DTYPE))
        at += n
    if device.type == "cuda":
        torch.cuda.current_stream().synchronize()
    return merged
