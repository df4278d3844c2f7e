#!/usr/bin/env python
"""Profile of the HOST layer above the C-ABI (Data / Pipeline / Results in --isoforms mode) in the build container, which has
no GPU: the decode comes from the CPU oracle and the device aggregation calls from tests/hostfake.py, so that what remains
under the profiler is exactly the Python that runs on the GPU box around the kernels.

    python tests/tools/host_profile.py [--bp 20000000] [--top 25] [--no-profile]

Input: the contigs of true code 11 or 15 among the first ``bp`` letters of the bench workload (what bench.py's
file_e2e.cli_isoforms leg runs).
"""
from __future__ import annotations

import argparse
import cProfile
import datetime
import os
import pstats
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=20_000_000)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--no-profile", action="store_true")
    a = ap.parse_args()
    import bench
    from mgcod_b200 import data as data_mod
    from mgcod_b200.engine import synth_metagenome
    from mgcod_b200.pipeline import Pipeline
    from mgcod_b200.results import Results
    from tests import hostfake
    work = tempfile.mkdtemp()
    model_paths = bench.write_models(os.path.join(work, "models"))
    buf, offs, truth = synth_metagenome(bench.SEED, a.bp, threads=8)
    two = [i for i in range(len(truth)) if int(truth[i]) in (11, 15)]
    fa = os.path.join(work, "sample2.fasta")
    seqs = []
    with open(fa, "wb") as fh:
        for i in two:
            s = buf[offs[i]:offs[i + 1]].tobytes()
            seqs.append(s.decode())
            fh.write(b">contig_%d synthetic\n" % i + s + b"\n")
    t0 = time.perf_counter()
    result, genes = hostfake.oracle_result(model_paths, seqs)
    print(f"{len(two)} contigs, {sum(map(len, seqs))} bp; oracle decode {time.perf_counter() - t0:.1f} s", flush=True)
    eng = hostfake.HostEngine(genes, seqs)

    class FakeData(data_mod.Data):
        """Data with the decode replaced: the engine is the host stand-in, the result the oracle's"""

        def _decode(self):
            self.engine = eng
            eng._resident_owner = self
            return result

        def ensure_resident(self):
            return eng, [r.id for r in self._file_order]

    import mgcod_b200.pipeline as pm
    pm.Data = FakeData
    pred, out = os.path.join(work, "pred") + "/", os.path.join(work, "out") + "/"
    os.makedirs(pred), os.makedirs(out)

    def run():
        P = Pipeline(fa, pred, "sample2", True, False, 3, 5, 5000, None, 30, False, False, work, model_dir=os.path.dirname(model_paths[0]))
        P()
        Results(P, "sample2", pred, fa, out, datetime.datetime.now(), False, False, False, False, work)()
        return P

    pr = cProfile.Profile()
    t0 = time.perf_counter()
    if a.no_profile:
        P = run()
    else:
        pr.enable()
        P = run()
        pr.disable()
    dt = time.perf_counter() - t0
    nwin = sum(len(l) for l in P.genome_label) if not isinstance(P.genome_label, (int, np.integer)) else 0
    print(f"host layer: {dt:.3f} s for {nwin} windows -> {nwin / dt:.0f} windows/s"
          f"{' (under cProfile)' if not a.no_profile else ''}")
    if not a.no_profile:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(a.top)


if __name__ == "__main__":
    main()
