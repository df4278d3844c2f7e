#!/usr/bin/env python
"""GPU box: cProfile of the whole command line in --isoforms mode on the two-code half of the bench workload (bench.py's
file_e2e.cli_isoforms_full), second call in the process.   python tests/tools/cli_profile.py [bp] [top]"""
import cProfile
import datetime
import os
import pstats
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from mgcod_b200.engine import synth_metagenome  # noqa: E402
from mgcod_b200.pipeline import Pipeline  # noqa: E402
from mgcod_b200.results import Results  # noqa: E402


def main():
    bp = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    work = tempfile.mkdtemp()
    model_paths = bench.write_models(os.path.join(work, "models"))
    buf, offs, truth = synth_metagenome(bench.SEED, bp, threads=16)
    two = [i for i in range(len(truth)) if int(truth[i]) in (11, 15)]
    fa = os.path.join(work, "full2.fasta")
    with open(fa, "wb") as fh:
        for i in two:
            fh.write(b">contig_%d synthetic\n" % i + buf[offs[i]:offs[i + 1]].tobytes() + b"\n")

    def run(tag):
        pred, out = os.path.join(work, f"pred_{tag}") + "/", os.path.join(work, f"out_{tag}") + "/"
        os.makedirs(pred), os.makedirs(out)
        t0 = time.perf_counter()
        P = Pipeline(fa, pred, tag, True, False, 3, 5, 5000, None, 30, False, False, work, model_dir=os.path.dirname(model_paths[0]))
        P()
        t1 = time.perf_counter()
        Results(P, tag, pred, fa, out, datetime.datetime.now(), False, False, False, False, work)()
        P.close()
        return t1 - t0, time.perf_counter() - t1, getattr(P, "timings", None)

    print("first call", run("a"))
    print("second call", run("b"))
    pr = cProfile.Profile()
    pr.enable()
    r = run("c")
    pr.disable()
    print("third call under cProfile", r)
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(top)
    st.sort_stats("cumulative").print_stats("mgcod_b200", top)


if __name__ == "__main__":
    main()
