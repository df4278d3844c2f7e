"""CPU: the N>1 path — contigs sharded over ranks, rank-local decode, gather on rank 0 — with world_size 2 on gloo.
The per-rank decode is the oracle here (no GPU in this container); the sharding and gather logic is the product's."""
import os
import socket

import numpy as np
import pytest

from mgcod_b200 import shard


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(1)
    lens = rng.integers(5000, 500000, 300)
    parts = shard.partition_contigs(lens, 8)
    allidx = np.sort(np.concatenate(parts))
    assert (allidx == np.arange(300)).all()
    loads = np.array([lens[p].sum() for p in parts])
    assert loads.max() - loads.min() <= lens.max()
    assert all((np.diff(p) > 0).all() for p in parts)


def _worker(rank, world, port, model_path, ret):
    import torch.distributed as dist
    from mgcod_b200.engine import synth_metagenome
    from oracle import gmhmm
    from tests import fixtures as fx
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf, offs, _ = synth_metagenome(5, 120000, threads=1)
    lens = np.diff(offs)
    mine = shard.partition_contigs(lens, world)[rank]
    genes, _ = fx.genes_array_from_oracle(gmhmm.Model(model_path), [buf[offs[c]:offs[c + 1]].tobytes() for c in mine])
    merged = shard.gather_genes(genes, mine)
    if rank == 0:
        full, _ = fx.genes_array_from_oracle(gmhmm.Model(model_path), [buf[offs[c]:offs[c + 1]].tobytes() for c in range(len(lens))])
        ret["ok"] = bool(merged.shape == full.shape and (merged == full).all())
        ret["n"] = int(full.shape[0])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank(model_dir):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as man:
        ret = man.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, f"{model_dir}/mgm_11.mod", ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert ret["ok"] and ret["n"] > 50


def test_range_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(2)
    lens = rng.integers(5000, 500000, 1000)
    for n in (1, 2, 3, 8):
        cut = shard.partition_ranges(lens, n)
        assert cut[0] == 0 and cut[-1] == 1000 and (np.diff(cut) >= 0).all()
        loads = np.array([lens[cut[r]:cut[r + 1]].sum() for r in range(n)])
        assert loads.max() - loads.min() <= lens.max()
    assert shard.partition_ranges([10, 10], 4).tolist() == [0, 1, 1, 2, 2]       # more ranks than contigs: empty shares
    assert shard.partition_ranges([], 3).tolist() == [0, 0, 0, 0]


def _range_worker(rank, world, port, model_paths, ret, how="collective"):
    import torch
    import torch.distributed as dist
    from mgcod_b200.engine import synth_metagenome
    from oracle import gmhmm
    from tests import fixtures as fx
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf, offs, _ = synth_metagenome(6, 150000, threads=1)
    cut = shard.partition_ranges(np.diff(offs), world)
    models = [gmhmm.Model(p) for p in model_paths]
    mine = range(int(cut[rank]), int(cut[rank + 1]))
    tables = [fx.genes_array_from_oracle(m, [buf[offs[c]:offs[c + 1]].tobytes() for c in mine])[0] for m in models]
    gather = shard.gather_gene_tables_shared if how == "shared" else shard.gather_gene_tables
    merged = gather(tables, int(cut[rank]), torch.device("cpu"))
    if how == "shared":      # a second call re-uses the segment; the first call's arrays are no longer valid
        merged = gather(tables, int(cut[rank]), torch.device("cpu"))
    if rank == 0:
        full = [fx.genes_array_from_oracle(m, [buf[offs[c]:offs[c + 1]].tobytes() for c in range(len(offs) - 1)])[0] for m in models]
        ret["ok"] = all(a.shape == b.shape and (a == b).all() for a, b in zip(merged, full))
        ret["n"] = int(sum(len(f) for f in full))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("how", ["collective", "shared"])
def test_two_rank_table_gather_equals_single_rank(model_dir, how):
    """the two gathers of the GPU path on gloo: the collective (counts all_gather + padded gather of the 32-byte gene rows of all
    models) and, for ranks of one node, the host segment every rank maps and writes its rows into"""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as man:
        ret = man.dict()
        paths = [f"{model_dir}/mgm_11.mod", f"{model_dir}/mgm_4.mod"]
        procs = [ctx.Process(target=_range_worker, args=(r, 2, port, paths, ret, how)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert ret["ok"] and ret["n"] > 100


@pytest.mark.gpu
def test_device_gather_on_one_gpu(model_dir):
    """both paths of the multi-GPU gather with a process group of one rank: the engine's device gene tables wrapped as
    torch tensors without a copy, counts exchanged, rows gathered — equal to the host-side result"""
    import torch
    import torch.distributed as dist
    from mgcod_b200.engine import Engine, synth_metagenome
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        buf, offs, _ = synth_metagenome(8, 400000, threads=2)
        eng = Engine([f"{model_dir}/mgm_{c}.mod" for c in (11, 4, 15, 101)])
        cut = shard.partition_ranges(np.diff(offs), 3)         # pretend to be the middle rank of three
        mine = offs[int(cut[1]):int(cut[2]) + 1]
        res = eng.decode_buffer(buf, mine)
        for how in ("shared", "nccl"):         # the host segment mapped by the ranks of a node, then the NCCL gather
            os.environ["MGCOD_B200_GATHER"] = how
            merged = shard.gather_genes_device(eng, int(cut[1]))
            for m in range(4):
                want = res.genes(m)
                want["contig"] += int(cut[1])
                assert len(want) > 20 and merged[m].tobytes() == want.tobytes(), how
        os.environ.pop("MGCOD_B200_GATHER", None)
        eng.close()
    finally:
        dist.destroy_process_group()
