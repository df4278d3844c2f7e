#!/usr/bin/env python
"""Golden vectors for ambiguous letters and gaps (IUPAC codes, U, runs of >= 40 non-ACGT letters).

Runs the reference's closed binary (/root/reference/dependencies/gmhmmp2) in the build container on inputs derived
from the committed fixtures; writes amb_input.fa.xz and amb_<model>.gff.xz next to this file.

    python tests/golden/make_golden_amb.py
"""
import lzma
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import run_gm, strip_dates, xz_write  # noqa: E402

MODELS = {"code4": "odd_model_code4.mod.xz", "trained11": "trained/mgm_11.mod.xz", "trained101": "trained/mgm_101.mod.xz"}


def build_records():
    seq = "".join(l.strip() for l in lzma.open(f"{HERE}/phage2.fa.xz", "rt") if not l.startswith(">")).upper()
    rng = random.Random(4040)
    recs = []

    def mutate(s, runs):
        s = list(s)
        for a, n, alphabet in runs:
            for i in range(a, min(len(s), a + n)):
                s[i] = rng.choice(alphabet)
        return "".join(s)

    base = seq[30000:42000]
    recs.append(("singles", mutate(base, [(rng.randrange(12000), 1, "NRYKMSWBDHV") for _ in range(40)])))
    recs.append(("short_runs", mutate(base, [(rng.randrange(12000), rng.choice([2, 5, 20, 39]), "N") for _ in range(12)])))
    recs.append(("gap40", mutate(base, [(3000, 40, "N"), (7000, 41, "NRY"), (9000, 39, "N")])))
    recs.append(("gap_long_lower", mutate(base, [(5000, 700, "n")]).replace("T", "U", 50)))
    recs.append(("gap_at_start", mutate(base, [(0, 120, "N"), (6000, 3, "N")])))
    recs.append(("gap_at_end", mutate(base, [(11900, 100, "N"), (2000, 45, "N")])))
    recs.append(("two_gaps_close", mutate(base, [(4000, 60, "N"), (4100, 60, "N")])))
    recs.append(("all_gap", "N" * 300))
    # ambiguous letters right behind start codons (the order-ramp positions) of the genes gmhmmp2 finds in the clean slice
    import subprocess
    with tempfile.TemporaryDirectory() as tmp:
        mod = os.path.join(tmp, "m.mod")
        open(mod, "wb").write(lzma.open(f"{HERE}/trained/mgm_11.mod.xz").read())
        fa = os.path.join(tmp, "b.fa")
        open(fa, "w").write(">b\n" + base + "\n")
        ramp = list(base)
        for line in run_gm(mod, fa).split("\n"):
            f = line.split("\t")
            if len(f) > 8 and f[2] == "CDS" and rng.random() < 0.7:
                left, right = int(f[3]), int(f[4])
                for o in rng.sample(range(3, 11), rng.choice([1, 2, 3])):
                    ramp[left - 1 + o if f[6] == "+" else right - 1 - o] = rng.choice("NNRY")
    recs.append(("ramp_ns", "".join(ramp)))
    recs.append(("clean", seq[60000:66000]))
    return recs


def main():
    recs = build_records()
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "amb.fa")
        with open(fa, "w") as fh:
            for name, s in recs:
                fh.write(f">{name}\n")
                fh.write("\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")
        xz_write(f"{HERE}/amb_input.fa.xz", open(fa, "rb").read())
        for tag, rel in MODELS.items():
            mod = os.path.join(tmp, tag + ".mod")
            open(mod, "wb").write(lzma.open(f"{HERE}/{rel}").read())
            xz_write(f"{HERE}/amb_{tag}.gff.xz", strip_dates(run_gm(mod, fa)).encode())
            print(tag, "ok")


if __name__ == "__main__":
    main()
