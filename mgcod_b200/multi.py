"""Many genomes — the reference's bin/multiprocess_mgcod.py (process pool over genome files, :140-145) on GPUs.

    python -m mgcod_b200.multi --genomes list.txt -o results/ -p mgm_results/ [mgcod.py flags]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m mgcod_b200.multi --genomes list.txt ...

``list.txt`` holds one FASTA path per line (as for the reference).  Every process takes the genomes ``rank::world`` (one
process per GPU under torchrun; RANK / WORLD_SIZE / LOCAL_RANK from the environment), keeps ONE engine with the four parameter
files resident, and runs the same per-genome pipeline as ``python -m mgcod_b200.cli``.  ``--threads`` and ``--batch_size`` are
accepted for compatibility and ignored: there is no process pool, the parallelism is inside the GPU.  A genome that fails is
logged and skipped, like a failing worker command in the reference (:46-54).
"""
from __future__ import annotations

import logging
import os
import sys

from . import cli

DRIVER_FLAGS = ("--genomes", "--threads", "--batch_size", "-i", "--path_to_genome", "--device")   # not part of the per-genome command


def main(argv=None) -> int:
    p = cli.build_parser()
    p.prog = "multiprocess_mgcod.py"
    for a in p._actions:
        if a.dest == "path_to_genome":
            a.required = False
    p.add_argument("--genomes", required=True, help="path to txt file with genomes (one FASTA path per line)")
    p.add_argument("--threads", type=int, default=8, help="(ignored) number of processes of the reference")
    p.add_argument("--batch_size", type=int, default=256, help="(ignored) batch size of the reference's pool")
    args, _ = p.parse_known_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    with open(args.genomes) as fh:
        genomes = [l.strip() for l in fh if l.strip()]
    mine = genomes[rank::world]
    base = []
    skip_next = False
    for tok in (sys.argv[1:] if argv is None else list(argv)):          # per-genome command = our flags minus the driver's own
        if skip_next:
            skip_next = False
            continue
        if tok in DRIVER_FLAGS:
            skip_next = True
            continue
        if tok.split("=", 1)[0] in DRIVER_FLAGS and "=" in tok:      # --genomes=list.txt, --threads=8, ...
            continue
        base.append(tok)
    failed = 0
    for g in mine:
        try:
            cli.main(["-i", g, "--device", str(local)] + base)
        except SystemExit as e:                    # cli exits like the reference ("<path> does not exist")
            if e.code not in (0, None):
                logging.error("mgcod failed for %s: %s", g, e.code)
                failed += 1
        except Exception as e:                     # noqa: BLE001 — one bad genome must not stop the others
            logging.error("mgcod failed for %s: %s", g, e)
            failed += 1
    if world > 1:
        print(f"rank {rank}/{world}: {len(mine) - failed} of {len(mine)} genomes done")
    return 1 if failed and failed == len(mine) else 0


if __name__ == "__main__":
    sys.exit(main())
