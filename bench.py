#!/usr/bin/env python
"""Benchmark of the hot path: bp x models / s decoded (BASELINE.json metric) on synthetic contigs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--total-bp B]

A "step" decodes one batch of synthetic contigs (BASELINE.json config 5: lengths ~ L^-2 on [5 kb, 500 kb], gene-like
content, four genetic codes) under the four parameter files (parametric 41 bacterial + 41 archaeal GC-bin
sections each, order 4/2).  At N = 1 the batch is the full 1 Gbp; under torchrun (N > 1) every rank decodes its own 1 Gbp
(`value`, `e2e`: WEAK scaling, no data-path collective; the only exchange is the max-over-ranks of the timings).

  value     device-resident: ASCII contigs already in HBM, K x (pack -> ORFs -> scores -> chain -> genes to host)
  e2e       the public call Engine.decode_stream() on pinned HOST buffers: every step copies its sequences H2D (on the copy
            stream, overlapping the previous batch's decode) and reads its genes back, all inside the timed region;
            e2e.single_call_* is Engine.decode_buffer() per step, nothing overlapped
  strong    ONE 1 Gbp metagenome (the same on every rank) split over the N ranks by mgcod_b200.shard.partition_ranges, each
            rank decoding its slice from host buffers, and the gene tables gathered on rank 0 over NCCL straight from device
            memory (shard.gather_genes_device) inside the timed region — north_star's target configuration; rank 0 checks
            the merged tables against a single-GPU decode of the whole input
  isoforms  windows/s of the --isoforms labelling (window sums + labels + noise labels + smoothing, one device call)
  file_e2e  100 Mbp FASTA file -> four GFF files, and the whole command line in --isoforms mode on its two-code half
  roofline  the dominant kernel (the pair walk) against the HBM roof of SURVEY.md §8(d) (0.40 B per bp x model); `roofs`
            adds the builder's compulsory-bytes figure and the L1 data-pipe (shared-memory wavefront) roof from the
            committed ncu capture, which is the one that binds
  cpu_baseline      B2 of BASELINE.md: the reference's gene finder (oracle/_ref/gmhmmp2) on all host cores, bounded sample
  cpu_baseline_b1   B1 of BASELINE.md: the reference's own `multiprocess_mgcod.py --threads $(nproc)` (unmodified scripts staged
                    under oracle/_ref/mgcod, Biopython etc. replaced by oracle/shims) on a subsample, extrapolation labelled

`--impl reference` times B2 on the box's host cores on a bounded sample of the same workload; that process loads nothing of
the product (the workload generator lives in workload/, the model writer is pure Python).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODES = (11, 4, 15, 101)
SEED = 20260101
METRIC = "bp_x_models_per_s_decoded"
UNIT = "bp*models/s"
# One launch of the dominant kernel (k_walk_pair) on the default workload (1 Gbp x 4 models), from the committed ncu capture
# profiles/r3_k_walk_pair_1Gbp.txt (ncu --set full): DRAM bytes moved and L1 data-pipe wavefronts (shared 4.51e9 + global:
# 73.8 % of one wavefront per SM and clock over the launch's 50.1 M cycles).
WALK_PROFILE = "profiles/r3_k_walk_pair_1Gbp.txt"
WALK_DRAM_BYTES_1GBP = 22.78e9
WALK_L1_WAVEFRONTS_1GBP = 5.47e9
SURVEY_BYTES_PER_UNIT = 0.40          # algorithmic bytes per bp x model


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def write_models(dirname: str) -> list[str]:
    from mgcod_b200 import synth          # pure Python: does not load the product library
    paths = synth.write_model_set(dirname, kind="parametric")
    return [paths[c] for c in CODES]


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baselines: the reference's own code on host cores, bounded samples
# --------------------------------------------------------------------------------------------------
def _ref_binary() -> str | None:
    p = os.path.join(ROOT, "oracle", "_ref", "gmhmmp2")
    if os.path.isfile(p) and os.access(p, os.X_OK):
        try:
            r = subprocess.run([p], capture_output=True, text=True, timeout=20)
            if "GeneMark.hmm-2" in (r.stdout + r.stderr):
                return p
        except (OSError, subprocess.TimeoutExpired):
            pass
    return None


def bind_to_gpu_numa_node(device: int):
    """Pins this rank to the CPUs next to its GPU (sysfs local_cpulist), so that the page-locked staging buffers it
    allocates afterwards are on the GPU's NUMA node.  Best effort: returns the cpu list used, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        text = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in text.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return text
    except Exception:
        pass
    return None


def cpu_reference_run(buf: np.ndarray, offs: np.ndarray, model_paths: list[str], sample_bp: int, workdir: str, cores: int) -> dict:
    """B2: decode the first contigs of the workload (about sample_bp) under the 4 models on `cores` host threads."""
    n = int(np.searchsorted(offs, sample_bp, side="left"))
    n = max(1, min(n, len(offs) - 1))
    bp = int(offs[n])
    binary = _ref_binary()
    nshard = max(1, min(cores, n))
    bounds = [int(round(i * n / nshard)) for i in range(nshard + 1)]
    t0 = time.perf_counter()
    if binary:
        shards = []
        for s in range(nshard):
            fa = os.path.join(workdir, f"shard{s}.fa")
            with open(fa, "wb") as fh:
                for c in range(bounds[s], bounds[s + 1]):
                    fh.write(b">c%d\n" % c)
                    fh.write(buf[offs[c]:offs[c + 1]].tobytes())
                    fh.write(b"\n")
            shards.append(fa)
        t0 = time.perf_counter()  # FASTA writing is not part of the reference's decode time
        jobs = [(fa, mp, os.path.join(workdir, f"out_{si}_{mi}.gff")) for si, fa in enumerate(shards) for mi, mp in enumerate(model_paths)]
        idx = 0
        running = []
        while idx < len(jobs) or running:
            while idx < len(jobs) and len(running) < cores:
                fa, mp, out = jobs[idx]
                running.append(subprocess.Popen([binary, "-M", mp, "-s", fa, "-f", "gff", "-o", out], stdout=subprocess.DEVNULL,
                                                stderr=subprocess.DEVNULL))
                idx += 1
            running = [p for p in running if p.poll() is None]
            time.sleep(0.002)
        kind = "reference"
        used = min(cores, len(jobs))
    else:
        from concurrent.futures import ThreadPoolExecutor
        from oracle import gmhmm
        models = [gmhmm.Model(p) for p in model_paths]

        def work(s):
            for c in range(bounds[s], bounds[s + 1]):
                seq = buf[offs[c]:offs[c + 1]].tobytes()
                for m in models:
                    m.decode(seq)
        with ThreadPoolExecutor(nshard) as ex:  # ctypes releases the GIL inside the C calls
            list(ex.map(work, range(nshard)))
        kind = "port"
        used = nshard
    dt = time.perf_counter() - t0
    out = {"value": bp * len(model_paths) / dt, "unit": UNIT, "cores": used, "kind": kind,
           "sample": f"first {n} contigs = {bp} bp x {len(model_paths)} models of the step's workload, {dt:.2f} s wall"}
    if binary:
        out["_gff"] = [[os.path.join(workdir, f"out_{si}_{mi}.gff") for si in range(nshard)] for mi in range(len(model_paths))]
        out["_contigs"] = n
    return out


def parity_against_baseline(engine, buf, offs, base: dict) -> dict:
    """The gene lines the reference binary has just printed for the CPU-baseline sample against the engine's genes for the same
    contigs: left, right, strand, %.2f score and the attributes behind the gene id (gene_type, partial, gc, length)."""
    import pandas as pd
    from mgcod_b200.data import GFF_COLUMNS, GeneFrames
    n = base["_contigs"]
    res = engine.decode_buffer(buf, offs[:n + 1])
    frames = GeneFrames([res.genes(m) for m in range(engine.n_models)], [f"c{i}" for i in range(n)])
    genes = differing = 0
    for mi, files in enumerate(base["_gff"]):
        ref = pd.concat([pd.read_csv(f, sep="\t", header=None, comment="#", names=GFF_COLUMNS) for f in files], ignore_index=True)
        mine = frames[mi]
        genes += len(ref)
        if len(ref) != len(mine):
            differing += abs(len(ref) - len(mine))
            continue
        same = ((ref["sequence"].astype(str).values == mine["sequence"].values) & (ref["start"].values == mine["start"].values)
                & (ref["end"].values == mine["end"].values) & (ref["strand"].values == mine["strand"].values)
                & (ref["logodd"].values == mine["logodd"].values)
                & (ref["attributes"].str.split(";", n=1).str[1].values == mine["attributes"].str.split(";", n=1).str[1].values))
        differing += int((~same).sum())
    return {"against": "the gene lines oracle/_ref/gmhmmp2 printed in the cpu_baseline leg of this run", "contigs": n, "genes": genes,
            "differing": differing}


def cpu_reference_b1(buf: np.ndarray, offs: np.ndarray, model_paths: list[str], n_contigs: int, workdir: str, cores: int) -> dict | None:
    """B1 of BASELINE.md §3: the reference's own fan-out, `bin/multiprocess_mgcod.py --genomes list.txt --threads $(nproc)
    --isoforms` (bin/multiprocess_mgcod.py:19-54, 140-145), unmodified, over one-contig-per-file FASTAs of the first
    `n_contigs` contigs of the workload.  Needs the staged tree oracle/_ref/mgcod (oracle/stage_reference.py)."""
    tree = os.path.join(ROOT, "oracle", "_ref", "mgcod")
    if not os.path.isfile(os.path.join(tree, "bin", "multiprocess_mgcod.py")) or not _ref_binary():
        return None
    n = max(1, min(n_contigs, len(offs) - 1))
    for code, mp in zip(CODES, model_paths):
        shutil.copyfile(mp, os.path.join(tree, "dependencies", f"mgm_{code}.mod"))
    gdir = os.path.join(workdir, "b1")
    os.makedirs(gdir, exist_ok=True)
    lst = os.path.join(gdir, "list.txt")
    with open(lst, "w") as fh:
        for c in range(n):
            fa = os.path.join(gdir, f"contig_{c}.fasta")
            with open(fa, "wb") as out:
                out.write(b">contig_%d\n" % c + buf[offs[c]:offs[c + 1]].tobytes() + b"\n")
            fh.write(fa + "\n")
    lens = np.diff(offs[:n + 1])
    windows = int(np.sum(np.ceil(lens / np.where(lens > 100000, 5000, 2500))))
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle", "shims"), MGCOD_REFERENCE_SHIMS="1")
    threads = max(1, min(cores, n))
    # the reference hands its pool chunks of --batch_size genomes (default 256): one chunk per process keeps every core busy
    cmd = [sys.executable, os.path.join(tree, "bin", "multiprocess_mgcod.py"), "--genomes", lst, "--threads", str(threads),
           "--batch_size", str((n + threads - 1) // threads), "--isoforms", "-o", os.path.join(gdir, "out") + "/",
           "-p", os.path.join(gdir, "pred") + "/", "--logfile", os.path.join(gdir, "log", "b1.log")]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=env, cwd=gdir, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    done = len([f for f in os.listdir(os.path.join(gdir, "out")) if f.endswith(".gff")]) if os.path.isdir(os.path.join(gdir, "out")) else 0
    bp = int(offs[n])
    if r.returncode != 0 or done == 0:
        return {"error": (r.stdout + r.stderr)[-300:], "rc": r.returncode}
    rate = bp * len(model_paths) / dt
    return {"value": rate, "unit": UNIT, "windows_per_s": windows / dt, "cores": threads, "kind": "reference",
            "path": "bin/multiprocess_mgcod.py --isoforms (unmodified, staged under oracle/_ref/mgcod; Biopython / matplotlib / "
                    "multiprocessing_logging replaced by oracle/shims)",
            "sample": f"first {n} contigs ({bp} bp, {windows} windows), one FASTA per contig, {dt:.1f} s wall, {done} of {n} finished",
            "extrapolated_1gbp_s": 1e9 * len(model_paths) / rate,
            "note": "measured rate on the subsample; the 1 Gbp time is an EXTRAPOLATION (about 3 s of interpreter start-up per contig)"}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--total-bp", type=int, default=int(os.environ.get("MGCOD_BENCH_BP", 1_000_000_000)),
                    help="bp per GPU per step (default: 1 Gbp, BASELINE.json config 5)")
    ap.add_argument("--cpu-sample-bp", type=int, default=128_000_000)
    ap.add_argument("--b1-contigs", type=int, default=int(os.environ.get("MGCOD_BENCH_B1_CONTIGS", 96)),
                    help="contigs of the B1 (multiprocess_mgcod.py) subsample; BASELINE.md asks for >= 2000 (minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workload = (f"synthetic metagenome seed {SEED}: contigs 5-500 kb (p(L)~L^-2), {args.total_bp} bp per GPU x 4 genetic codes "
                f"(11/4/15/101), parametric 82-section order-4/2 models")
    sample_bp = min(args.total_bp, max(args.cpu_sample_bp, 1))
    config = {"workload": workload, "bp_per_gpu": args.total_bp, "models": 4,
              "l2": "inputs (1 B/nt ASCII + 0.5 B/nt planes) far exceed the 126 MB L2; no explicit flush",
              # the CPU arms (cpu_baseline here, the whole line of --impl reference) decode a bounded SAMPLE of this workload —
              # its first contigs, same seed — and report the rate measured on it
              "reference_sample_bp": sample_bp}

    workdir = tempfile.mkdtemp(prefix=f"mgcod_bench_{rank}_")
    try:
        model_paths = write_models(os.path.join(workdir, "models"))
        if args.impl == "reference":
            if rank != 0:
                return 0
            from workload import synth_metagenome
            # the sample = the first contigs of the very workload the other arm decodes (same seed, same generator): contigs
            # are generated independently, so the prefix of a larger draw equals a smaller draw
            buf, offs, _ = synth_metagenome(SEED, sample_bp, threads=min(cores, 16))
            vals = []
            for it in range(args.warmup + args.steps):
                r = cpu_reference_run(buf, offs, model_paths, sample_bp, workdir, cores)
                if it >= args.warmup:
                    vals.append(r)
            v = float(np.mean([r["value"] for r in vals]))
            bp = int(offs[-1])
            last = {k: x for k, x in vals[-1].items() if not k.startswith("_")}
            last["value"] = v
            print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": 1e3 * bp * 4 / v, "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": last,
                              "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0,
                              "note": f"each step decodes the {bp} bp sample named in config.reference_sample_bp and cpu_baseline.sample; "
                                      f"value is the rate on it, ms_per_step the time of one sample"}))
            return 0

        import torch
        import torch.distributed as dist
        from mgcod_b200 import shard
        from mgcod_b200.engine import Engine
        from workload import synth_metagenome

        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(local)
        numa = bind_to_gpu_numa_node(local) if world > 1 else None   # before any page-locked allocation (first touch)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # workload in pinned host memory
        t0 = time.perf_counter()
        pinned = torch.empty(args.total_bp + 600_000, dtype=torch.uint8, pin_memory=True)
        buf, offs, truth = synth_metagenome(SEED + rank, args.total_bp, max_len=int(os.environ.get('MGCOD_BENCH_MAXLEN', 500000)),
                                            threads=min(cores, 32), out=pinned.numpy())
        gen_s = time.perf_counter() - t0
        total_bp = int(offs[-1])
        engine = Engine(model_paths, device=local)
        units = total_bp * engine.n_models

        # ---- warm-up (e2e path: allocations, first-touch) ----
        for _ in range(args.warmup):
            res = engine.decode_buffer(buf, offs)
            del res
        for res in engine.decode_stream((buf, offs) for _ in range(max(2, args.warmup))):   # the timed e2e path
            del res
        free_b, total_b = torch.cuda.mem_get_info()
        hbm_used_gb = (total_b - free_b) / 1e9       # buffers of one 1 Gbp x 4 models batch incl. the prefetch buffer
        # ---- device-resident timing ----
        engine.upload(buf, offs)
        res = engine.decode_resident()
        del res
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = engine.decode_resident()      # returns when the kernels are done; the gene tables travel on the copy-out stream
            del res                             # while the next step computes (the closing barrier waits for the last copy)
        barrier()
        dt_res = time.perf_counter() - t0
        clocks = sampler.stop()
        # phase split (CUDA events inside the engine) from untimed steps: reading the copy-out time waits for that copy, which
        # the timed loop above must not do
        phase = {}
        launches = 0
        n_phase_steps = min(args.steps, 3)
        for _ in range(n_phase_steps):
            res = engine.decode_resident()
            st = engine.stats()
            for k in ("pack", "orf_find", "orf_prep", "orf_score", "orf_final", "order", "chain", "aggregate", "d2h", "total"):
                phase[k] = phase.get(k, 0.0) + st[k]
            launches += st["kernel_launches"]
            n_genes = st["n_genes"]
            d2h_bytes = n_genes * 32 + engine.n_contigs * 8 * engine.n_models
            counts = {k: st[k] for k in ("n_orfs", "n_starts", "n_walks", "n_candidates", "n_genes")}
            del res
        launches = launches * args.steps // n_phase_steps      # launches inside the timed region (the same every step)
        phase = {k: v * args.steps / n_phase_steps for k, v in phase.items()}
        # ---- end-to-end timing through the public call with host buffers ----
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = engine.decode_buffer(buf, offs)
            cents = res.model_cents()
            del res
        barrier()
        dt_single = time.perf_counter() - t0
        # the same through Engine.decode_stream(): a stream of host batches, batch i+1 copied while batch i is decoded.
        # Every step still moves its own H2D bytes (adopted from the copy stream) and reads its genes back.
        barrier()
        t0 = time.perf_counter()
        for res in engine.decode_stream((buf, offs) for _ in range(args.steps)):
            cents = res.model_cents()
            del res
        barrier()
        dt_e2e = time.perf_counter() - t0

        # ---- --isoforms labelling of the resident result (BASELINE.json configs[1] parameters: -w 5000 -st 2500 -n 3) ----
        engine.upload(buf, offs)
        res = engine.decode_resident()
        lens = np.diff(offs)
        strides = np.full(len(lens), 2500, np.int32)
        for _ in range(2):      # warm-up: device / host buffers of the labelling call
            woffs, wlab, wcents, wmerged, wstatus = engine.isoform_labels(lens, strides, 5000, 3, False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            woffs, wlab, wcents, wmerged, wstatus = engine.isoform_labels(lens, strides, 5000, 3, False)
        torch.cuda.synchronize()
        dt_win = (time.perf_counter() - t0) / args.steps
        n_windows = int(woffs[-1])
        del res

        # ---- strong scaling: ONE 1 Gbp metagenome over the N ranks through shard.py, gene tables gathered over NCCL ----
        strong = None
        if not os.environ.get("MGCOD_BENCH_SKIP_STRONG"):
            if world > 1 and rank != 0:   # every rank needs the same input: rank r > 0 generated seed + r above
                sbuf_t = torch.empty(args.total_bp + 600_000, dtype=torch.uint8, pin_memory=True)
                sbuf, soffs, _ = synth_metagenome(SEED, args.total_bp, max_len=int(os.environ.get('MGCOD_BENCH_MAXLEN', 500000)),
                                                  threads=min(cores, 32), out=sbuf_t.numpy())
            else:
                sbuf, soffs = buf, offs
            tp0 = time.perf_counter()
            cut = shard.partition_ranges(np.diff(soffs), world)
            partition_ms = 1e3 * (time.perf_counter() - tp0)
            mine = soffs[int(cut[rank]):int(cut[rank + 1]) + 1]

            for _ in range(2):
                r_ = engine.decode_buffer(sbuf, mine)          # H2D of my slice + decode + genes to my host
                merged = shard.gather_genes_device(engine, int(cut[rank])) if world > 1 else None
                del r_
            r_ = None
            barrier()
            t0 = time.perf_counter()
            t_gather = 0.0
            for _ in range(args.steps):
                r_ = None                                      # the previous result goes back to the page-locked pool first
                r_ = engine.decode_buffer(sbuf, mine)
                tg = time.perf_counter()
                # N = 1: the result of the one rank IS the merged table, already in page-locked host memory
                merged = shard.gather_genes_device(engine, int(cut[rank])) if world > 1 else None
                t_gather += time.perf_counter() - tg
                if world > 1:
                    del r_
            barrier()
            dt_strong = time.perf_counter() - t0
            equal = None
            if rank == 0:
                # the merged tables against one GPU decoding the whole input (the N = 1 result)
                if world == 1:
                    merged = [r_.genes(m) for m in range(engine.n_models)]
                full = engine.decode_buffer(sbuf, soffs)
                equal = all(len(merged[m]) == full.num_genes(m) and merged[m].tobytes() == full.genes(m).tobytes()
                            for m in range(engine.n_models))
                del full
            strong = {"dt": dt_strong, "gather_s": t_gather, "partition_ms": partition_ms, "equal": equal,
                      "bp": int(soffs[-1]), "my_bp": int(mine[-1] - mine[0]), "my_contigs": len(mine) - 1}

        # ---- second shape of SURVEY.md 8(d): 200k contigs of 5 kb (launch/overhead behaviour) ----
        shape2 = None
        if world == 1 and not os.environ.get("MGCOD_BENCH_SKIP_SHAPE2"):
            # its own host buffer: `buf` (a view of the pinned block) is still needed by the file and CPU-baseline legs below
            buf2, offs2, _ = synth_metagenome(SEED + 1000, args.total_bp, min_len=5000, max_len=5000, threads=min(cores, 32))
            engine.upload(buf2, offs2)
            for _ in range(2):
                del_me = engine.decode_resident()
                del del_me
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                del_me = engine.decode_resident()
                del del_me
            torch.cuda.synchronize()
            dt2 = (time.perf_counter() - t0) / args.steps
            shape2 = {"workload": f"{len(offs2) - 1} contigs x 5 kb, same generator", "ms_per_step": 1e3 * dt2,
                      "value": int(offs2[-1]) * engine.n_models / dt2, "unit": UNIT, "resident": True}
            del buf2

        # ---- file to files on a 100 Mbp slice: native FASTA read -> decode -> four GFF files (SURVEY.md 8d "also report") ----
        file_e2e = None
        if world == 1 and not os.environ.get("MGCOD_BENCH_SKIP_FILES"):
            file_e2e = file_legs(engine, buf, offs, truth, model_paths, workdir, local, torch)

        # ---- latency of the reference's own example (BASELINE.json configs[0]/[1]: phage2, 173 747 bp, one contig) ----
        small = None
        golden = os.path.join(ROOT, "tests", "golden")
        if world == 1 and os.path.exists(os.path.join(golden, "phage2.fa.xz")) and not os.environ.get("MGCOD_BENCH_SKIP_SMALL"):
            small = small_input_leg(golden, workdir, local, torch, Engine)

        strong_t = [strong["dt"], strong["gather_s"]] if strong else [0.0, 0.0]
        times = torch.tensor([dt_res, dt_e2e, dt_single, dt_win] + strong_t, dtype=torch.float64, device="cuda")
        tot_units = torch.tensor([float(units)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot_units, op=dist.ReduceOp.SUM)
        dt_res, dt_e2e, dt_single, dt_win, dt_strong, dt_gather = (float(x) for x in times)
        all_units = float(tot_units[0])

        if rank == 0:
            peak, peak_src = _peaks()
            default_workload = args.total_bp == 1_000_000_000 and "MGCOD_BENCH_MAXLEN" not in os.environ
            score_ms = phase["orf_score"] / args.steps
            units_launch = total_bp * engine.n_models
            # (1) the contract's roof: SURVEY.md §8(d) algorithmic bytes, 0.40 B per bp x model, through the dominant kernel
            survey_bytes = SURVEY_BYTES_PER_UNIT * units_launch
            achieved = survey_bytes / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
            # (2) what this implementation must move at least in that launch (all 4 models): both strands' planes once per model
            # (0.5 B/nt), per walk its 32-byte item, per start its record read and its four sums written (16 + 32 B); the
            # section-pair tables (224 KB each) live in shared memory and are not counted
            alg_bytes = 4 * 0.5 * total_bp + counts["n_walks"] * 32 + counts["n_starts"] * 48
            mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
            wave_peak = 148 * mhz * 1e6                 # one L1 data-pipe wavefront per SM per clock
            roofs = {
                "hbm_survey_8d": {"bytes_per_unit": SURVEY_BYTES_PER_UNIT, "achieved_gbs": achieved, "peak_gbs": peak, "frac": achieved / peak},
                "hbm_compulsory": {"bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0,
                                   "peak_gbs": peak, "frac": alg_bytes / (score_ms * 1e-3) / 1e9 / peak if score_ms > 0 else 0.0},
                "l1_data_pipe_wavefronts": ({"wavefronts_per_launch": WALK_L1_WAVEFRONTS_1GBP,
                                             "achieved_per_s": WALK_L1_WAVEFRONTS_1GBP / (score_ms * 1e-3), "peak_per_s": wave_peak,
                                             "frac": WALK_L1_WAVEFRONTS_1GBP / (score_ms * 1e-3) / wave_peak, "source": WALK_PROFILE,
                                             "note": "random 16-byte table lookups in shared memory: the roof that binds this kernel"}
                                            if default_workload else None),
            }
            out = {
                "metric": METRIC, "value": all_units * args.steps / dt_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dt_res / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "contigs_per_gpu": int(len(truth)),
                "e2e": {"value": all_units * args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(total_bp + 8 * (len(offs))),
                        "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": 1e3 * dt_e2e / args.steps,
                        "api": "Engine.decode_stream (H2D of batch i+1 overlaps the decode of batch i)",
                        "single_call_ms_per_step": 1e3 * dt_single / args.steps,
                        "single_call_value": all_units * args.steps / dt_single},
                "isoforms": {"metric": "windows_per_s_labelled", "windows_per_gpu": n_windows, "w": 5000, "stride": 2500, "n": 3,
                             "labelling_ms": 1e3 * dt_win, "api": "Engine.isoform_labels (window sums, labels, noise labels, smoothing in one device call)",
                             "value_labelling_only": world * n_windows / dt_win,
                             "value_with_decode": world * n_windows / (dt_win + dt_res / args.steps), "unit": "windows/s"},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"kernel": "k_walk_pair", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": WALK_DRAM_BYTES_1GBP if default_workload else None,
                             "traffic_source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch on this "
                                               f"workload ({WALK_PROFILE})" if default_workload else None,
                             "peak_source": peak_src, "kernel_ms_per_launch": score_ms,
                             "algorithmic_bytes_per_launch": survey_bytes,
                             "algorithmic_bytes_source": "SURVEY.md 8(d): 0.40 B per bp x model x the 4e9 units of one launch",
                             "note": "the path is not HBM-bound (see roofs.l1_data_pipe_wavefronts, DESIGN.md section 4)"},
                "roofs": roofs,
                "phases_ms_per_step": {k: v / args.steps for k, v in phase.items()},
                "counts_per_step": counts,
                "host": {"cores": cores, "workload_generation_s": gen_s, "numa_cpus": numa},
                "hbm_used_gb": hbm_used_gb,
            }
            if strong:
                sunits = strong["bp"] * engine.n_models
                out["strong"] = {"workload": f"one {strong['bp']} bp metagenome (seed {SEED}) x 4 models over {world} GPU(s)", "scaling": "strong",
                                 "ms_per_step": 1e3 * dt_strong / args.steps, "value": sunits * args.steps / dt_strong, "unit": UNIT,
                                 "gather_ms_per_step": 1e3 * dt_gather / args.steps, "partition_ms": strong["partition_ms"],
                                 "rank0_bp": strong["my_bp"], "rank0_contigs": strong["my_contigs"],
                                 "merged_equals_single_gpu_decode": strong["equal"],
                                 "api": "shard.partition_ranges -> Engine.decode_buffer(host slice) per rank -> shard.gather_genes_device (ranks of one node: rows copied into a host segment all ranks map; else NCCL "
                                        "gather of the 32-byte gene rows from device memory), max over ranks"}
            if shape2:
                out["other_shapes"] = [shape2]
            if file_e2e:
                out["file_e2e"] = file_e2e
            if small:
                out["small_input_latency"] = small
            if world == 1 and not args.no_cpu_baseline:
                base = cpu_reference_run(buf, offs, model_paths, args.cpu_sample_bp, workdir, cores)
                if "_gff" in base:       # the binary's output is at hand: compare it with the engine's genes, line by line
                    try:
                        out["parity_sample"] = parity_against_baseline(engine, buf, offs, base)
                    except Exception as ex:
                        out["parity_sample"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                out["cpu_baseline"] = {k: v for k, v in base.items() if not k.startswith("_")}
                if args.b1_contigs > 0:
                    b1 = cpu_reference_b1(buf, offs, model_paths, args.b1_contigs, workdir, cores)
                    if b1:
                        out["cpu_baseline_b1"] = b1
            print(json.dumps(out))
        engine.close()
        if world > 1:
            dist.destroy_process_group()
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
    return 0


def file_legs(engine, buf, offs, truth, model_paths, workdir, local, torch):
    from concurrent.futures import ThreadPoolExecutor
    from mgcod_b200.data import Fasta
    nrec = int(np.searchsorted(offs, 100_000_000, side="left"))
    fa = os.path.join(workdir, "sample.fasta")
    with open(fa, "wb") as fh:
        for i in range(nrec):
            fh.write(b">contig_%d synthetic\n" % i)
            rec = buf[offs[i]:offs[i + 1]]
            k = len(rec) // 80
            rows = np.empty((k, 81), np.uint8)          # 80 letters + newline per line
            rows[:, :80] = rec[:k * 80].reshape(k, 80)
            rows[:, 80] = 10
            fh.write(rows.tobytes())
            if len(rec) > k * 80:
                fh.write(rec[k * 80:].tobytes() + b"\n")
    Fasta(fa).close()            # the reader keeps its page-locked letter buffer for the next file of the process: the first
    #                              read of a process pays cudaHostAlloc (~0.3 ms per MB), reported separately below
    t00 = time.perf_counter()
    f = Fasta(fa)
    f.close()
    t0 = time.perf_counter()
    f = Fasta(fa)
    t1 = time.perf_counter()
    r = engine.decode_buffer(f.buffer, f.offsets)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    lens_f = np.diff(f.offsets)

    def write_four():
        with ThreadPoolExecutor(max_workers=4) as tp:      # as Data.run_mgm does: the C writer runs without the GIL
            for j in [tp.submit(r.write, mi, f.ids, lens_f, fa, model_paths[mi], code, os.path.join(workdir, f"mgm_{code}_sample.gff"), "gff")
                      for mi, code in enumerate((11, 4, 15, 101))]:
                j.result()

    write_four()
    t3 = time.perf_counter()
    assert int(f.offsets[-1]) == int(offs[nrec]), "FASTA round trip lost letters"
    file_e2e = {"bp": int(f.offsets[-1]), "contigs": nrec, "file_bytes": os.path.getsize(fa), "read_fasta_s": t1 - t0,
                "read_fasta_gbs": os.path.getsize(fa) / (t1 - t0) / 1e9, "read_fasta_previous_s": t0 - t00,
                "decode_s": t2 - t1, "write_4_gff_s": t3 - t2,
                "value": int(f.offsets[-1]) * engine.n_models / (t3 - t0), "unit": UNIT,
                "genes_written": int(sum(r.num_genes(mi) for mi in range(engine.n_models)))}
    del r
    f.close()
    # The whole command line in --isoforms mode (python -m mgcod_b200.cli: FASTA -> decode -> four GFFs -> window labels ->
    # smoothing -> isoform comparison of contigs with a switch -> .gff / .tab / _stats.tab), on two inputs: the two-code half of
    # the 100 Mbp file (round 1's number) and the two-code half of the WHOLE workload (fixed costs amortised).  Input = the
    # contigs whose true code is 11 or 15: Mgcod models at most two codes per input (labels other than the two most frequent
    # ones are noise, bin/pipeline.py:1084), and a contig made of noise labels only stops the reference's smoothing (and this
    # mirror) with an IndexError.
    import datetime
    from mgcod_b200.pipeline import Pipeline
    from mgcod_b200.results import Results

    def cli_run(tag, last_rec):
        """twice: the first call of a process on a new input size pays device and page-locked allocations (and, for the very
        first, the engine set-up: CUDA context, four parameter files); the second is what `python -m mgcod_b200.multi` sees for
        every further genome"""
        first = cli_once(tag, last_rec)
        again = cli_once(tag + "_again", last_rec)
        if "error" in first or "error" in again:
            return first if "error" in first else again
        again["first_call_seconds"] = first["seconds"]
        again["first_call_parts_s"] = first["pipeline_parts_s"]
        return again

    def cli_once(tag, last_rec):
        try:
            cli_out = os.path.join(workdir, f"cli_out_{tag}") + "/"
            pred_dir = os.path.join(workdir, f"cli_pred_{tag}") + "/"
            path = os.path.join(workdir, f"{tag}.fasta")
            two = [i for i in range(last_rec) if int(truth[i]) in (11, 15)]
            with open(path, "wb") as fh:
                for i in two:
                    fh.write(b">contig_%d synthetic\n" % i)
                    rec = buf[offs[i]:offs[i + 1]]
                    for k in range(0, len(rec), 1 << 20):
                        fh.write(rec[k:k + (1 << 20)].tobytes())
                    fh.write(b"\n")
            lens2 = np.array([offs[i + 1] - offs[i] for i in two], dtype=np.int64)
            os.makedirs(pred_dir, exist_ok=True)
            os.makedirs(cli_out, exist_ok=True)
            t4 = time.perf_counter()
            P = Pipeline(path, pred_dir, tag, True, False, 3, 5, 5000, None, 30, False, False, workdir,
                         model_dir=os.path.dirname(model_paths[0]), device=local)
            P()
            t5 = time.perf_counter()
            Results(P, tag, pred_dir, path, cli_out, datetime.datetime.now(), False, False, False, False, workdir)()
            P.close()                      # as mgcod_b200.cli does: buffers back to their pools
            t6 = time.perf_counter()
            tab = open(os.path.join(cli_out, f"{tag}.tab")).read().split("\n")
            windows = float(np.sum(np.ceil(lens2 / np.where(lens2 > 100000, 5000, 2500))))
            return {"seconds": t6 - t4, "pipeline_s": t5 - t4, "results_s": t6 - t5, "pipeline_parts_s": getattr(P, "timings", None), "tab_rows": len(tab), "contigs": len(two),
                    "bp": int(lens2.sum()), "windows": int(windows), "windows_per_s": windows / (t6 - t4),
                    "bp_x_models_per_s": int(lens2.sum()) * 4 / (t6 - t4),
                    "genes_in_final_gff": sum(1 for l in open(os.path.join(cli_out, f"{tag}.gff")) if not l.startswith("#") and l.strip())}
        except Exception as ex:      # a host-side failure here must not take the headline numbers down with it
            return {"error": f"{type(ex).__name__}: {ex}"[:900]}
    file_e2e["cli_isoforms"] = cli_run("sample2", nrec)
    if not os.environ.get("MGCOD_BENCH_SKIP_CLI_FULL"):
        file_e2e["cli_isoforms_full"] = cli_run("full2", len(truth))
    return file_e2e


def small_input_leg(golden, workdir, local, torch, Engine):
    import lzma
    text = lzma.open(os.path.join(golden, "phage2.fa.xz"), "rt").read()
    pseq = np.frombuffer("".join(l.strip() for l in text.split("\n") if not l.startswith(">")).encode(), dtype=np.uint8)
    ppin = torch.empty(len(pseq), dtype=torch.uint8, pin_memory=True)
    ppin.numpy()[:] = pseq
    poffs = np.array([0, len(pseq)], dtype=np.int64)
    tdir = os.path.join(workdir, "trained")
    os.makedirs(tdir, exist_ok=True)
    tpaths = []
    for c in (11, 4, 15, 101):
        tp = os.path.join(tdir, f"mgm_{c}.mod")
        open(tp, "wb").write(lzma.open(os.path.join(golden, "trained", f"mgm_{c}.mod.xz")).read())
        tpaths.append(tp)
    eng2 = Engine(tpaths, device=local)
    plens, pstr = np.array([len(pseq)], np.int64), np.array([2500], np.int32)
    for _ in range(5):
        r = eng2.decode_buffer(ppin.numpy(), poffs)
        eng2.isoform_labels(plens, pstr, 5000, 3, False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        r = eng2.decode_buffer(ppin.numpy(), poffs)
        label = int(np.argmax(r.model_cents()))
    t1 = time.perf_counter()
    for _ in range(20):
        r = eng2.decode_buffer(ppin.numpy(), poffs)
        wo, wl, wc, wm, ws = eng2.isoform_labels(plens, pstr, 5000, 3, False)
    t2 = time.perf_counter()
    small = {"input": "phage2 (173747 bp, 1 contig) x 4 trained order-4/2 models, host buffers in, genes/labels out",
             "default_mode_ms": 1e3 * (t1 - t0) / 20, "isoforms_w5000_st2500_n3_ms": 1e3 * (t2 - t1) / 20,
             "windows": int(wo[-1]), "genetic_code": (11, 4, 15, 101)[label]}
    eng2.close()
    return small


if __name__ == "__main__":
    sys.exit(main())
