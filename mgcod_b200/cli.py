"""Command line of the reference (bin/mgcod.py:22-181) on top of the CUDA engine.

    python -m mgcod_b200.cli -i genome.fasta [-p PRED_DIR] [-o OUT_DIR] [--isoforms] [-n 3] [-g 5] [-w 5000] [-st 5000]
                             [-t 30] [-r] [-d] [-AA] [-NT] [--short_contigs] [--logfile F] [-v]

Same flags, defaults and output files.  Two additions: ``--model_dir`` (where mgm_11/4/15/101.mod live; default
``<package parent>/dependencies`` like the reference, or $MGCOD_MODEL_DIR) and ``--device``.  ``-m/--path_to_plots`` is
accepted and ignored with a warning (plots are outside the hot path).
"""
from __future__ import annotations

import argparse
import datetime
import logging
import os
import re
import sys
import time
import warnings


def build_parser() -> argparse.ArgumentParser:
    cwd = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = argparse.ArgumentParser(prog="mgcod.py", description="Mgcod segments contigs based on genetic code usage and performs "
                                "genetic-code-informed annotation of coding regions (GPU engine)")
    req = p.add_argument_group("required arguments")
    req.add_argument("-i", "--path_to_genome", required=True, help="Path to input file, FASTA format")
    opt = p.add_argument_group("optional arguments")
    opt.add_argument("-p", "--path_to_mgm_predictions", dest="path_to_predictions", default=f"{cwd}/mgm_results/",
                     help="Directory where to save MGM predictions so that they can be re-used")
    opt.add_argument("-o", "--path_to_output", dest="output", default=f"{cwd}/results/",
                     help="Directory where to save final gene annotations and supporting outputs")
    opt.add_argument("-m", "--path_to_plots", default=None, help="(ignored) directory for plots")
    opt.add_argument("-r", "--circular", action="store_true", default=False, help="Set if sequence is circular")
    opt.add_argument("--isoforms", action="store_true", default=False, help="Enable prediction of isoforms")
    opt.add_argument("-n", "--consecutive_windows", type=int, default=3)
    opt.add_argument("-g", "--consecutive_gene_labels", type=int, default=5)
    opt.add_argument("-w", "--window_size", type=int, default=5000)
    opt.add_argument("-st", "--stride", type=int, default=None)
    opt.add_argument("-t", "--tolerance", type=int, default=30)
    opt.add_argument("-d", "--delete", action="store_true", default=False, help="Delete intermediary files")
    opt.add_argument("-AA", "--amino_acids", action="store_true", default=False)
    opt.add_argument("-NT", "--nucleotides", action="store_true", default=False)
    opt.add_argument("--short_contigs", action="store_true", default=False)
    opt.add_argument("--logfile", default=None)
    opt.add_argument("-v", "--verbose", action="store_true", default=False)
    opt.add_argument("--version", action="version", version="%(prog)s 1.0.1 (mgcod-b200)")
    opt.add_argument("--model_dir", default=None, help="directory holding mgm_11.mod, mgm_4.mod, mgm_15.mod, mgm_101.mod")
    opt.add_argument("--device", type=int, default=0, help="CUDA device index")
    p.set_defaults(cwd=cwd)
    return p


def main(argv=None) -> int:
    from .pipeline import Pipeline
    from .results import Results

    date = datetime.datetime.now()
    a = build_parser().parse_args(argv)
    logfile = a.logfile
    if not logfile:
        t = time.localtime()
        logfile = "log/mgcod_{}{}{}{}{}{}.log".format(t.tm_year, t.tm_mon, t.tm_mday, t.tm_hour, t.tm_min, t.tm_sec)
    logdir = "/".join(logfile.split("/")[:-1])
    if logdir and not os.path.isdir(logdir):
        os.makedirs(logdir)
    logging.basicConfig(filename=logfile, level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s", datefmt="%H:%M:%S")
    if not os.path.isfile(a.path_to_genome):
        sys.exit("{} does not exist".format(a.path_to_genome))
    pred, out = a.path_to_predictions, a.output
    for d in (pred, out):
        if not os.path.isdir(d):
            os.makedirs(d)
    if not pred.endswith("/"):
        pred += "/"
    if not out.endswith("/"):
        out += "/"
    if a.path_to_plots is not None:
        warnings.warn("-m/--path_to_plots is ignored: plotting is not part of this build")
    suffix = "." + a.path_to_genome.split(".")[-1]
    species = a.path_to_genome.split("/")[-1].split(suffix)[0]
    for ch in "'()":
        species = re.sub("[" + ch + "]", "\\\\" + ch, species)
    pipeline = Pipeline(a.path_to_genome, pred, species, a.isoforms, a.circular, a.consecutive_windows, a.consecutive_gene_labels,
                        a.window_size, a.stride, a.tolerance, a.short_contigs, a.verbose, a.cwd, model_dir=a.model_dir, device=a.device)
    pipeline()
    results = Results(pipeline, species, pred, a.path_to_genome, out, date, a.delete, a.amino_acids, a.nucleotides, a.verbose, a.cwd)
    results()
    pipeline.close()
    if a.verbose:
        print("----------------------------Done----------------------------")
    return 0


if __name__ == "__main__":
    sys.exit(main())
