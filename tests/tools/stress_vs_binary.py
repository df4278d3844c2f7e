#!/usr/bin/env python
"""Stress harness: the LIVE reference binary (gmhmmp2 --debug 1 and its GFF) against the CPU oracle over random parameter
files and inputs.  Build container only — it needs /root/reference (binary + the two example genomes).

    python tests/tools/stress_vs_binary.py models [seed] [cases]     random orders 0-5 / 0-3, 1-10 GC bins, one or two
                                                                      domains, codes 11/4/15/101/25/1, slices / reverse
                                                                      complements / splices of both example genomes;
                                                                      compares candidate sets, totals, gc, domain, every
                                                                      chain value and the gene lines
    python tests/tools/stress_vs_binary.py ambiguous [seed] [cases]  runs of 1..100 IUPAC letters sprinkled over slices
    python tests/tools/stress_vs_binary.py ramp [seed] [cases]       ambiguous letters placed 3..12 nt behind start codons
                                                                      and at the record edges (order ramp, virtual starts)
    python tests/tools/stress_vs_binary.py keys [seed] [cases]       parameter files with the optional keys of SURVEY §A.2

This is the harness DESIGN.md §2 refers to (round 1 kept it in a scratch directory).  Exit status 1 on any failure; the
failing parameter file and input are left in the work directory that is printed.
"""
from __future__ import annotations

import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("MGCOD_REFERENCE", "/root/reference")
GM = f"{REF}/dependencies/gmhmmp2"

from mgcod_b200 import synth  # noqa: E402
from oracle import gmhmm  # noqa: E402


def rc(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


def run_gm(mod, fasta, out, extra=()):
    if os.path.exists(out):
        os.remove(out)
    r = subprocess.run([GM, "-M", mod, "-s", fasta, "-f", "gff", "-o", out, *extra], capture_output=True, text=True,
                       errors="replace")
    return r.returncode, r.stdout + r.stderr


def parse_debug(text):
    """`--debug 1` map: one dict per point."""
    pts = []
    for line in text.split("\n"):
        if " ; " not in line:
            continue
        parts = line.strip().split(" ; ")
        if len(parts) != 3:
            continue
        a, b, c = (p.split() for p in parts)
        if len(a) != 9:
            continue
        pts.append(dict(gc=int(a[2]), pos=int(a[3]), stop=int(a[4]), len=int(a[5]), type=a[6], best=float(b[0]),
                        cum=float(c[0]), ptr=int(c[1]), dom=c[2] if len(c) > 2 else ""))
    return pts


def gff_rows(path):
    out = []
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        t = line.rstrip("\n").split("\t")
        out.append((int(t[3]), int(t[4]), t[6], t[5], t[8].split("gene_type ")[1]))
    return out


def oracle_rows(model, seq):
    out = []
    for g in model.decode(seq)[0]:
        attr = ("archaea" if g["domain"] else "bacteria") + "; "
        if g["partial_left"] or g["partial_right"]:
            attr += f"partial {g['partial_left']}{g['partial_right']}; "
        attr += f"gc {g['gc']}; length {g['right'] - g['left'] + 1};"
        out.append((g["left"], g["right"], "+" if g["strand"] > 0 else "-", "%.2f" % g["logodd"], attr))
    return out


def compare_debug(work, modpath, seq):
    """candidate level: sets, totals, gc, domain, chain values; then gene lines."""
    fa, out = os.path.join(work, "c.fa"), os.path.join(work, "c.gff")
    open(fa, "w").write(f">t\n{seq}\n")
    _, dbg = run_gm(modpath, fa, out, ["--debug", "1"])
    pts = parse_debug(dbg)
    M = gmhmm.Model(modpath)
    cands, _, _ = M.candidates(seq)
    refc = {}
    for p in pts:
        t = p["type"]
        if "sta" in t:
            part = t.endswith("C_") or "ini" in t or "ter" in t
            refc[(p["pos"], p["stop"], 1, part) if t.startswith("D") else (p["stop"], p["pos"], -1, part)] = p
    minec = {(c["left"], c["right"], c["strand"], bool(c["partial_start"])): c for c in cands}
    msgs = []
    om, orf = sorted(set(minec) - set(refc)), sorted(set(refc) - set(minec))
    if om or orf:
        msgs.append(f"candidates only oracle {om[:4]} only binary {orf[:4]}")
    common = set(minec) & set(refc)
    for k in common:
        m, r = minec[k], refc[k]
        if abs(m["total"] - r["best"]) > 6e-4 + 1e-5 * abs(r["best"]):
            msgs.append(f"total {k}: {m['total']} vs {r['best']}")
        if m["gc"] != r["gc"]:
            msgs.append(f"gc {k}: {m['gc']} vs {r['gc']}")
        if ("arc" if m["domain"] else "bac") != r["dom"]:
            msgs.append(f"domain {k}")
        if k[2] < 0 and abs(m["cum"] - r["cum"]) > 6e-4 + 1e-5 * abs(r["cum"]):
            msgs.append(f"chain value {k}: {m['cum']} vs {r['cum']}")
    ends = {p["pos"]: p for p in pts if p["type"].startswith("D-end")}
    best_d = {}
    for c in cands:
        if c["strand"] > 0:
            best_d[c["right"]] = max(best_d.get(c["right"], -1e300), c["cum"])
    for r, b in best_d.items():
        if r in ends and abs(b - ends[r]["cum"]) > 6e-4 + 1e-5 * abs(b):
            msgs.append(f"chain value at D-end {r}: {b} vs {ends[r]['cum']}")
    if oracle_rows(M, seq) != gff_rows(out):
        a, b = set(oracle_rows(M, seq)), set(gff_rows(out))
        msgs.append(f"genes only oracle {sorted(a - b)[:3]} only binary {sorted(b - a)[:3]}")
    return msgs, len(cands)


def compare_gff(work, modpath, seq):
    fa, out = os.path.join(work, "c.fa"), os.path.join(work, "c.gff")
    open(fa, "w").write(f">t\n{seq}\n")
    code, text = run_gm(modpath, fa, out)
    if code != 0 or not os.path.exists(out):       # the binary reports some errors with exit status 0 and no output
        try:
            gmhmm.Model(modpath)
        except RuntimeError:
            return [], 0      # both refuse the file
        return [f"binary failed ({text.strip()[:80]}), oracle accepted"], 0
    ref = gff_rows(out)
    try:
        model = gmhmm.Model(modpath)
    except RuntimeError as ex:
        return [f"oracle refused ({ex}), binary accepted"], len(ref)
    mine = oracle_rows(model, seq)
    if mine != ref:
        a, b = set(mine), set(ref)
        return [f"genes only oracle {sorted(a - b)[:3]} only binary {sorted(b - a)[:3]}"], len(ref)
    return [], len(ref)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "models"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ncase = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rng = random.Random(seed)
    work = tempfile.mkdtemp(prefix=f"stress_{mode}_")
    ph = synth.read_fasta(f"{REF}/examples/phage2.fasta")[0][2]
    gca = synth.read_fasta(f"{REF}/examples/GCA_002924505.1_ASM292450v1_genomic.fasta")[0][2]
    xph, xg = synth.encode(ph), synth.encode(gca)
    fails = 0
    for it in range(ncase):
        mp = os.path.join(work, "m.mod")
        if mode == "models":
            code = rng.choice([11, 4, 15, 101, 11, 25, 1])
            kc = rng.choice([0, 1, 2, 3, 4, 5])
            kn = rng.randint(0, min(kc, 3))
            nb = rng.choice([1, 1, 2, 3, 5, 10])
            doms = rng.choice([("B",), ("B", "A"), ("B", "A")])
            tmpl = rng.choice([xph, xg])
            if nb == 1 and doms == ("B",):
                secs = synth.trained_sections(tmpl, code, kc, kn)
                secs[0].label = f"__B_GC_{rng.choice([0, 30, 50, 100])}"
            else:
                secs = synth.trained_sections(tmpl, code, kc, kn, bins=sorted(rng.sample(range(15, 60), nb)), domains=doms)
            minl = rng.choice([90, 90, 60, 120, 300])
            for s in secs:
                s.gene_min_length = minl
                s.cod_duration_decay = rng.choice([300.0, 250.0])
                s.non_duration_decay = rng.choice([150.0, 120.0])
            synth.write_mod(mp, secs)
            src = rng.choice([ph, gca])
            L = rng.choice([50, 200, 1000, 3000, 10000, 30000])
            lo = rng.randrange(0, len(src) - L)
            seq = src[lo:lo + L]
            if rng.random() < 0.3:
                seq = rc(seq)
            if rng.random() < 0.2:
                lo2 = rng.randrange(0, len(src) - L)
                seq = seq + rc(src[lo2:lo2 + L])
            msgs, n = compare_debug(work, mp, seq)
            tag = f"code {code} orders {kc}/{kn} sections {len(secs)} L {len(seq)} min {minl}"
        else:
            code = rng.choice([11, 4, 15, 101])
            kc = rng.choice([1, 2, 3, 4, 5])
            kn = min(rng.choice([0, 1, 2, 3]), kc)
            doms = rng.choice([("B",), ("B", "A")])
            bins = sorted(rng.sample(range(20, 50), rng.choice([1, 3])))
            if mode == "keys":      # labels whose lexicographic and numeric orders differ (9 / 35 / 100)
                bins = sorted(rng.sample([5, 9, 20, 28, 35, 42, 100], rng.choice([2, 3, 4])))
            secs = synth.trained_sections(xph, code, kc, kn, bins=bins, domains=doms)
            for s in secs:
                if s.label.startswith("__A"):
                    s.cod_p_n, s.non_p_n = 0.3, 0.7
            synth.write_mod(mp, secs)
            src = rng.choice([ph, gca])
            L = rng.choice([3000, 10000, 30000]) if mode == "ambiguous" else rng.choice([5000, 12000])
            lo = rng.randrange(0, len(src) - L)
            s = list(src[lo:lo + L])
            if mode == "ambiguous":
                for _ in range(rng.choice([1, 3, 10, 30])):
                    a = rng.randrange(0, L)
                    for i in range(a, min(L, a + rng.choice([1, 1, 2, 5, 20, 39, 40, 41, 100]))):
                        s[i] = rng.choice("NNNNRYKMSWBDHVn")
            elif mode == "ramp":
                fa, out = os.path.join(work, "b.fa"), os.path.join(work, "b.gff")
                open(fa, "w").write(">t\n" + "".join(s) + "\n")
                run_gm(mp, fa, out)
                for left, right, strand, _, _ in gff_rows(out):
                    if rng.random() < 0.6:
                        for o in rng.sample(range(3, 3 + kc + 4), rng.choice([1, 1, 2, 3])):
                            p = (left - 1 + o) if strand == "+" else (right - 1 - o)
                            if 0 <= p < L:
                                s[p] = rng.choice("NNRYn")
                for _ in range(rng.choice([0, 1, 2])):
                    s[rng.choice([rng.randrange(0, 8), L - 1 - rng.randrange(0, 8)])] = "N"
            elif mode == "keys":
                # optional keys in front of the matrices of randomly chosen sections (the binary lets the section that comes
                # last in the lexicographic order of the labels decide)
                text = open(mp).read().split("\n")
                ats = [i for i, l in enumerate(text) if l.startswith("$NON_MAT")]
                for at in sorted(rng.sample(ats, rng.randint(1, len(ats))), reverse=True):
                    text[at:at] = [k.format(v=rng.choice(["0.3", "0.7", "0.05", "0.5", "1", "0.9"]))
                                   for k in rng.sample(KEY_LINES, rng.choice([1, 1, 2, 3]))]
                open(mp, "w").write("\n".join(text))
            msgs, n = compare_gff(work, mp, "".join(s))
            tag = f"code {code} orders {kc}/{kn} sections {len(secs)} L {L}"
        print(it, tag, "items", n, "OK" if not msgs else "FAIL", flush=True)
        if msgs:
            fails += 1
            for m in msgs[:6]:
                print("    ", m)
            os.replace(mp, os.path.join(work, f"fail_{it}.mod"))
            os.replace(os.path.join(work, "c.fa"), os.path.join(work, f"fail_{it}.fa"))
    print(f"{mode}: {ncase - fails}/{ncase} cases equal to the binary; work directory {work}")
    return 1 if fails else 0


# optional keys of SURVEY.md §A.2 with a value the parser accepts (see DESIGN.md §2 for what each one does)
KEY_LINES = ["$TO_MGM {v}", "$TO_NATIVE {v}", "$TO_ATYPICAL_FIRST_BACTERIA {v}", "$TO_ATYPICAL_SECOND_ARCHAEA {v}",
             "$BUILD test", "$RBS 0", "$PROMOTER 0", "$GENE_GC_DIST 0", "$TO_MGM -1"]

if __name__ == "__main__":
    sys.exit(main())
