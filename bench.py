#!/usr/bin/env python
"""Benchmark of the hot path: bp x models / s decoded (BASELINE.json metric) on synthetic contigs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--total-bp B]

A "step" decodes one batch of synthetic contigs (BASELINE.json config 5: lengths ~ L^-2 on
[5 kb, 500 kb], gene-like content, four genetic codes) under the four parameter files
(parametric 41 bacterial + 41 archaeal GC-bin sections each, order 4/2).  At N = 1 the batch is the full
1 Gbp; under torchrun (N > 1) every rank decodes its own 1 Gbp shard (weak scaling, no data-path
collective; the only exchange is the final max-over-ranks of the timings).

  value   device-resident: ASCII contigs already in HBM, K x (pack -> ORFs -> scores -> chain -> genes to host)
  e2e     the public call Engine.decode_stream() on pinned HOST buffers: every step copies its sequences H2D (on the
          copy stream, overlapping the previous batch's decode) and reads its genes back, all inside the timed
          region; e2e.single_call_* is Engine.decode_buffer() per step, nothing overlapped
  isoforms        windows/s of the --isoforms labelling (window sums + labels + label merging) on the resident result
  other_shapes    the "200k contigs x 5 kb" shape of SURVEY.md 8(d), device-resident

`--impl reference` times the reference's own gene finder (oracle/_ref/gmhmmp2, the binary Mgcod shells out
to; else the C port in oracle/) on the box's host cores on a bounded sample of the same workload.
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
# DRAM bytes of ONE k_walk launch on the default 1 Gbp x 4 models workload, from the committed ncu capture
# profiles/r1j_k_walk_1Gbp.txt (15.19 GB read + 4.64 GB written; the algorithmic bytes are 12.4 GB, the rest is 32-byte
# sector granularity on 16-byte records and plane groups -- DESIGN.md section 4; 48.5 GB before the tile-major item order).
WALK_DRAM_BYTES_1GBP = 19.83e9
METRIC = "bp_x_models_per_s_decoded"
UNIT = "bp*models/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _alu_roofline(units_per_launch, kernel_ms, clocks):
    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    peak = 148 * 128 * mhz * 1e6 / 1e12
    achieved = 26.0 * units_per_launch / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
    return {"kernel": "k_walk", "bound": "alu", "achieved": achieved, "peak": peak, "unit": "Top/s", "frac": achieved / peak,
            "ops_per_unit": 26, "sm_mhz": mhz,
            "note": "algorithmic op count of SURVEY.md 8(d); the kernel's measured limiter is the shared-memory data pipe "
                    "(profiles/r1f_top_kernels_1Gbp.txt)"}


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
    from mgcod_b200 import synth
    paths = synth.write_model_set(dirname, kind="parametric")
    return [paths[c] for c in CODES]


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own gene finder on host cores, bounded sample
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
    """Decode the first contigs of the workload (about sample_bp) under the 4 models on `cores` host threads."""
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
        procs, idx = [], 0
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
    return {"value": bp * len(model_paths) / dt, "unit": UNIT, "cores": used, "kind": kind,
            "sample": f"first {n} contigs = {bp} bp x {len(model_paths)} models of the step's workload, {dt:.2f} s wall"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--total-bp", type=int, default=int(os.environ.get("MGCOD_BENCH_BP", 1_000_000_000)),
                    help="bp per GPU per step (default: 1 Gbp, BASELINE.json config 5)")
    ap.add_argument("--cpu-sample-bp", type=int, default=128_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workload = (f"synthetic metagenome seed {SEED}: contigs 5-500 kb (p(L)~L^-2), {args.total_bp} bp per GPU x 4 genetic codes "
                f"(11/4/15/101), parametric 82-section order-4/2 models")
    config = {"workload": workload, "bp_per_gpu": args.total_bp, "models": 4,
              "l2": "inputs (1 B/nt ASCII + 0.5 B/nt planes) far exceed the 126 MB L2; no explicit flush"}

    workdir = tempfile.mkdtemp(prefix=f"mgcod_bench_{rank}_")
    try:
        model_paths = write_models(os.path.join(workdir, "models"))
        if args.impl == "reference":
            if rank != 0:
                return 0
            from mgcod_b200.engine import synth_metagenome
            sample = min(args.total_bp, max(args.cpu_sample_bp, 1))
            buf, offs, _ = synth_metagenome(SEED, sample, threads=min(cores, 16))
            vals = []
            for it in range(args.warmup + args.steps):
                r = cpu_reference_run(buf, offs, model_paths, sample, workdir, cores)
                if it >= args.warmup:
                    vals.append(r)
            v = float(np.mean([r["value"] for r in vals]))
            bp = int(offs[-1])
            last = vals[-1]
            last["value"] = v
            print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": 1e3 * bp * 4 / v, "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": last,
                              "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0}))
            return 0

        import torch
        import torch.distributed as dist
        from mgcod_b200.engine import Engine, synth_metagenome

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
        phase = {}
        launches = 0
        for _ in range(args.steps):
            res = engine.decode_resident()
            st = engine.stats()
            for k in ("pack", "orf_find", "orf_prep", "orf_score", "orf_final", "order", "chain", "aggregate", "d2h", "total"):
                phase[k] = phase.get(k, 0.0) + st[k]
            launches += st["kernel_launches"]
            n_genes = st["n_genes"]
            d2h_bytes = n_genes * 32 + engine.n_contigs * 8 * engine.n_models
            counts = {k: st[k] for k in ("n_orfs", "n_starts", "n_walks", "n_candidates", "n_genes")}
            del res
        barrier()
        dt_res = time.perf_counter() - t0
        clocks = sampler.stop()
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
        for _ in range(2):      # warm-up: device / host buffers of the labelling calls
            woffs, wlab, wcents = engine.window_labels(lens, strides, 5000, 3, False)
            merged, mstatus = engine.merge_labels_flat(wlab, woffs, 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t_parts = [0.0, 0.0]
        for _ in range(args.steps):
            ta = time.perf_counter()
            woffs, wlab, wcents = engine.window_labels(lens, strides, 5000, 3, False)
            tb = time.perf_counter()
            merged, mstatus = engine.merge_labels_flat(wlab, woffs, 3)
            tc_ = time.perf_counter()
            t_parts[0] += tb - ta
            t_parts[1] += tc_ - tb
        torch.cuda.synchronize()
        dt_win = (time.perf_counter() - t0) / args.steps
        n_windows = int(woffs[-1])
        del res

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
            t0 = time.perf_counter()
            f = Fasta(fa)
            t1 = time.perf_counter()
            r = engine.decode_buffer(f.buffer, f.offsets)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            lens_f = np.diff(f.offsets)
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=4) as tp:      # as Data.run_mgm does: the C writer runs without the GIL
                for j in [tp.submit(r.write, mi, f.ids, lens_f, fa, model_paths[mi], code, os.path.join(workdir, f"mgm_{code}_sample.gff"), "gff")
                          for mi, code in enumerate((11, 4, 15, 101))]:
                    j.result()
            t3 = time.perf_counter()
            assert int(f.offsets[-1]) == int(offs[nrec]), "FASTA round trip lost letters"
            file_e2e = {"bp": int(f.offsets[-1]), "contigs": nrec, "read_fasta_s": t1 - t0, "decode_s": t2 - t1, "write_4_gff_s": t3 - t2,
                        "value": int(f.offsets[-1]) * engine.n_models / (t3 - t0), "unit": UNIT,
                        "genes_written": int(sum(r.num_genes(mi) for mi in range(engine.n_models)))}
            del r
            f.close()
            # the whole command line on the same file, isoform mode: FASTA -> decode -> four GFFs -> window labels -> label
            # smoothing -> isoform comparison of contigs with a switch -> .gff / .tab / _stats.tab  (python -m mgcod_b200.cli)
            # Input: the contigs of that sample whose true code is 11 or 15 — Mgcod models at most two codes per input (labels
            # other than the two most frequent ones are noise, bin/pipeline.py:1084), and a contig made of noise labels only
            # stops the reference's smoothing (and this mirror) with an IndexError.
            try:
                cli_out = os.path.join(workdir, "cli_out") + "/"
                fa = os.path.join(workdir, "sample2.fasta")
                two = [i for i in range(nrec) if int(truth[i]) in (11, 15)]
                with open(fa, "wb") as fh:
                    for i in two:
                        fh.write(b">contig_%d synthetic\n" % i)
                        rec = buf[offs[i]:offs[i + 1]]
                        for k in range(0, len(rec), 1 << 20):
                            fh.write(rec[k:k + (1 << 20)].tobytes())
                        fh.write(b"\n")
                lens_f = np.array([offs[i + 1] - offs[i] for i in two], dtype=np.int64)
                # what python -m mgcod_b200.cli does (mgcod_b200/cli.py main), kept in reach for the diagnostics below
                import datetime
                from mgcod_b200.pipeline import Pipeline
                from mgcod_b200.results import Results
                pred_dir = os.path.join(workdir, "cli_pred") + "/"
                os.makedirs(pred_dir, exist_ok=True)
                os.makedirs(cli_out, exist_ok=True)
                t4 = time.perf_counter()
                P = Pipeline(fa, pred_dir, "sample2", True, False, 3, 5, 5000, None, 30, False, False, workdir,
                             model_dir=os.path.dirname(model_paths[0]), device=local)
                try:
                    P()
                except IndexError as ex:
                    labs = P.genome_label if isinstance(P.genome_label, list) else []
                    allab = np.concatenate(labs) if labs else np.zeros(0, dtype=np.int64)
                    hist_l = np.bincount(allab, minlength=4)
                    noise = set(np.argsort(hist_l)[::-1][2:].tolist())
                    bad = [(k, P.sequence[k].id, len(P.sequence[k]), l.tolist()) for k, l in enumerate(labs)
                           if len(l) and all(int(x) in noise for x in l)]
                    raise IndexError(f"{ex}; label histogram {hist_l.tolist()}, contigs of noise labels only: {bad[:3]}")
                Results(P, "sample2", pred_dir, fa, cli_out, datetime.datetime.now(), False, False, False, False, workdir)()
                rc = 0
                t5 = time.perf_counter()
                tab = open(os.path.join(cli_out, "sample2.tab")).read().split("\n") if rc == 0 else []
                file_e2e["cli_isoforms"] = {"seconds": t5 - t4, "rc": rc, "tab_rows": len(tab), "contigs": len(two), "bp": int(lens_f.sum()),
                                            "windows_per_s": float(np.sum(np.ceil(lens_f / np.where(lens_f > 100000, 5000, 2500)))) / (t5 - t4),
                                            "genes_in_final_gff": sum(1 for l in open(os.path.join(cli_out, "sample2.gff")) if not l.startswith("#") and l.strip())}
            except Exception as ex:      # a host-side failure here must not take the headline numbers down with it
                file_e2e["cli_isoforms"] = {"error": f"{type(ex).__name__}: {ex}"[:900]}

        # ---- latency of the reference's own example (BASELINE.json configs[0]/[1]: phage2, 173 747 bp, one contig) ----
        small = None
        golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden")
        if world == 1 and os.path.exists(os.path.join(golden, "phage2.fa.xz")) and not os.environ.get("MGCOD_BENCH_SKIP_SMALL"):
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
                eng2.window_labels(plens, pstr, 5000, 3, False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                r = eng2.decode_buffer(ppin.numpy(), poffs)
                label = int(np.argmax(r.model_cents()))
            t1 = time.perf_counter()
            for _ in range(20):
                r = eng2.decode_buffer(ppin.numpy(), poffs)
                wo, wl, wc = eng2.window_labels(plens, pstr, 5000, 3, False)
                wm, _ = eng2.merge_labels_flat(wl, wo, 3)
            t2 = time.perf_counter()
            small = {"input": "phage2 (173747 bp, 1 contig) x 4 trained order-4/2 models, host buffers in, genes/labels out",
                     "default_mode_ms": 1e3 * (t1 - t0) / 20, "isoforms_w5000_st2500_n3_ms": 1e3 * (t2 - t1) / 20,
                     "windows": int(wo[-1]), "genetic_code": (11, 4, 15, 101)[label]}
            eng2.close()

        times = torch.tensor([dt_res, dt_e2e, dt_single, dt_win], dtype=torch.float64, device="cuda")
        tot_units = torch.tensor([float(units)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot_units, op=dist.ReduceOp.SUM)
        dt_res, dt_e2e, dt_single, dt_win = (float(x) for x in times)
        all_units = float(tot_units[0])

        if rank == 0:
            peak, peak_src = _peaks()
            default_workload = args.total_bp == 1_000_000_000 and "MGCOD_BENCH_MAXLEN" not in os.environ
            # dominant kernel = k_walk (phase "orf_score").  Algorithmic bytes of one launch (all 4 models): both strands'
            # planes once per model (0.5 B/nt), per walk its item + ORF record (16 + 24 B), per start its record read and its
            # four sums written (16 + 32 B).  The section-pair tables (104 KB each) live in shared memory and are not counted.
            score_ms = phase["orf_score"] / args.steps
            alg_bytes = 4 * 0.5 * total_bp + counts["n_walks"] * 40 + counts["n_starts"] * 48
            achieved = alg_bytes / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
            out = {
                "metric": METRIC, "value": all_units * args.steps / dt_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dt_res / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(config, contigs_per_gpu=int(len(truth))),
                "e2e": {"value": all_units * args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(total_bp + 8 * (len(offs))),
                        "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": 1e3 * dt_e2e / args.steps,
                        "api": "Engine.decode_stream (H2D of batch i+1 overlaps the decode of batch i)",
                        "single_call_ms_per_step": 1e3 * dt_single / args.steps,
                        "single_call_value": all_units * args.steps / dt_single},
                "isoforms": {"metric": "windows_per_s_labelled", "windows_per_gpu": n_windows, "w": 5000, "stride": 2500, "n": 3,
                             "labelling_ms": 1e3 * dt_win, "window_labels_ms": 1e3 * t_parts[0] / args.steps, "merge_labels_ms": 1e3 * t_parts[1] / args.steps, "value_labelling_only": world * n_windows / dt_win,
                             "value_with_decode": world * n_windows / (dt_win + dt_res / args.steps), "unit": "windows/s"},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"kernel": "k_walk", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": WALK_DRAM_BYTES_1GBP if default_workload else None,
                             "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one k_walk launch on this "
                                               "workload (profiles/r1j_k_walk_1Gbp.txt)" if default_workload else None,
                             "peak_source": peak_src,
                             "kernel_ms_per_launch": score_ms, "algorithmic_bytes_per_launch": alg_bytes},
                # SURVEY.md section 8(d) asks for the second candidate roof as well: ~26 simple ops (k-mer index updates, table
                # lookups' adds, prefix sums) per bp x model against 148 SMs x 128 lanes x the SM clock seen during the run
                "roofline_alu": _alu_roofline(total_bp * engine.n_models, score_ms, clocks),
                "phases_ms_per_step": {k: v / args.steps for k, v in phase.items()},
                "counts_per_step": counts,
                "host": {"cores": cores, "workload_generation_s": gen_s, "numa_cpus": numa},
                "hbm_used_gb": hbm_used_gb,
            }
            if shape2:
                out["other_shapes"] = [shape2]
            if file_e2e:
                out["file_e2e"] = file_e2e
            if small:
                out["small_input_latency"] = small
            if world == 1 and not args.no_cpu_baseline:
                out["cpu_baseline"] = cpu_reference_run(buf, offs, model_paths, args.cpu_sample_bp, workdir, cores)
            print(json.dumps(out))
        engine.close()
        if world > 1:
            dist.destroy_process_group()
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
