#!/usr/bin/env python
"""Stages the reference's own Python layer for the B1 baseline (`multiprocess_mgcod.py --threads $(nproc)`).

    python oracle/stage_reference.py            (build container; needs /root/reference)

Copies /root/reference/bin/*.py verbatim into oracle/_ref/mgcod/bin/ and the gene finder into oracle/_ref/mgcod/dependencies/
(the layout mgcod.py expects: <root>/bin, <root>/dependencies).  oracle/_ref/ is git-ignored — no reference source enters the
repository's history — but travels to the GPU box with the snapshot, like the staged binary.  The parameter files
(mgm_<code>.mod) are written next to the binary by whoever runs the baseline (bench.py writes the synthetic sets there).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MGCOD_REFERENCE", "/root/reference")


def main():
    if not os.path.isdir(f"{REF}/bin"):
        print(f"reference checkout absent ({REF}): keeping existing oracle/_ref/mgcod (if any)")
        return 0
    root = os.path.join(HERE, "_ref", "mgcod")
    os.makedirs(f"{root}/bin", exist_ok=True)
    os.makedirs(f"{root}/dependencies", exist_ok=True)
    for name in os.listdir(f"{REF}/bin"):
        if name.endswith(".py"):
            shutil.copyfile(f"{REF}/bin/{name}", f"{root}/bin/{name}")
            os.chmod(f"{root}/bin/{name}", 0o755)
    shutil.copyfile(f"{REF}/dependencies/gmhmmp2", f"{root}/dependencies/gmhmmp2")
    os.chmod(f"{root}/dependencies/gmhmmp2", 0o755)
    print(f"staged {root}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
