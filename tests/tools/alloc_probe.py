#!/usr/bin/env python
"""Where the time of a FIRST decode goes (engine set-up, device and page-locked allocations) against a repeated one.
    MGCOD_B200_TRACE=1 python tests/tools/alloc_probe.py"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from mgcod_b200.data import shared_engine  # noqa: E402
from workload import synth_metagenome  # noqa: E402


def main():
    work = tempfile.mkdtemp()
    paths = bench.write_models(os.path.join(work, "models"))
    t0 = time.perf_counter()
    eng = shared_engine(os.path.dirname(paths[0]))
    print(f"engine set-up (4 parameter files parsed + uploaded, CUDA context): {time.perf_counter() - t0:.3f} s", flush=True)
    for bp in (50_000_000, 50_000_000, 500_000_000, 500_000_000, 50_000_000):
        buf, offs, _ = synth_metagenome(bench.SEED, bp, threads=16)
        t0 = time.perf_counter()
        r = eng.decode_buffer(buf, offs)
        t1 = time.perf_counter()
        g = [r.genes(m) for m in range(4)]
        t2 = time.perf_counter()
        print(f"{bp} bp: decode_buffer {t1 - t0:.3f} s (device {eng.stats()['total']:.1f} ms), genes to numpy {t2 - t1:.3f} s", flush=True)
        del r


def cli_flow():
    """the command line's flow (Pipeline + Results, --isoforms) twice on the two-code half of a 1 Gbp workload"""
    import datetime
    import numpy as np
    from mgcod_b200.pipeline import Pipeline
    from mgcod_b200.results import Results
    work = tempfile.mkdtemp()
    paths = bench.write_models(os.path.join(work, "models"))
    buf, offs, truth = synth_metagenome(bench.SEED, 1_000_000_000, threads=16)
    two = [i for i in range(len(truth)) if int(truth[i]) in (11, 15)]
    fa = os.path.join(work, "full2.fasta")
    with open(fa, "wb") as fh:
        for i in two:
            fh.write(b">contig_%d synthetic\n" % i + buf[offs[i]:offs[i + 1]].tobytes() + b"\n")
    for it in range(3):
        pred, out = os.path.join(work, f"pred{it}") + "/", os.path.join(work, f"out{it}") + "/"
        os.makedirs(pred), os.makedirs(out)
        t0 = time.perf_counter()
        P = Pipeline(fa, pred, "full2", True, False, 3, 5, 5000, None, 30, False, False, work, model_dir=os.path.dirname(paths[0]))
        P()
        t1 = time.perf_counter()
        Results(P, "full2", pred, fa, out, datetime.datetime.now(), False, False, False, False, work)()
        t2 = time.perf_counter()
        print(f"run {it}: pipeline {t1 - t0:.3f} s {P.timings}, results {t2 - t1:.3f} s", flush=True)


if __name__ == "__main__":
    cli_flow() if "--cli" in sys.argv else main()
