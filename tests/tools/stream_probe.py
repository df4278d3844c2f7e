"""Timing probe for Engine.decode_stream: wall time of each call and the device phases per step."""
import sys, time, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mgcod_b200 import synth
from mgcod_b200.engine import Engine, synth_metagenome

d = tempfile.mkdtemp()
synth.write_model_set(d)
bp = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
pinned = torch.empty(bp + 600_000, dtype=torch.uint8, pin_memory=True)
buf, offs, _ = synth_metagenome(20260101, bp, threads=16, out=pinned.numpy())
eng = Engine([f"{d}/mgm_{c}.mod" for c in (11, 4, 15, 101)])
for _ in range(3):
    eng.decode_buffer(buf, offs)
def t(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    a = t(); eng.upload(buf, offs); b = t(); r = eng.decode_resident(); c = t()
    print(f"single: upload {1e3*(b-a):.1f} decode {1e3*(c-b):.1f} dev_total {eng.stats()['total']:.1f}")
eng.upload(buf, offs)
for it in range(4):
    a = time.perf_counter(); eng.prefetch(buf, offs); b = time.perf_counter(); r = eng.decode_resident(); c = time.perf_counter()
    st = eng.stats()
    eng.upload(buf, offs); dd = time.perf_counter()
    print(f"stream: prefetch {1e3*(b-a):.1f} decode {1e3*(c-b):.1f} adopt {1e3*(dd-c):.1f} dev_total {st['total']:.1f} "
          + " ".join(f"{k}={st[k]:.1f}" for k in ("pack", "orf_find", "orf_prep", "orf_score", "orf_final", "chain", "d2h")))
# the loop bench.py times as e2e: the generator, the scores read, the result dropped at once
marks = []
def batches(n):
    for _ in range(n):
        yield buf, offs
torch.cuda.synchronize()
t0 = time.perf_counter()
for res in eng.decode_stream(batches(5)):
    a = time.perf_counter(); cents = res.model_cents(); b = time.perf_counter(); del res; c = time.perf_counter()
    marks.append((a - t0, b - a, c - b))
    t0 = time.perf_counter()
torch.cuda.synchronize()
print("generator loop: " + "; ".join(f"step {1e3*x:.1f} cents {1e3*y:.2f} free {1e3*z:.2f}" for x, y, z in marks), f"tail {1e3*(time.perf_counter()-t0):.1f}")
# device-resident loop as bench.py times it
eng.upload(buf, offs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    a = time.perf_counter(); res = eng.decode_resident(); b = time.perf_counter(); del res
    print(f"resident: decode_resident {1e3*(b-a):.2f} ms, free {1e3*(time.perf_counter()-b):.2f} ms", file=sys.stderr)
torch.cuda.synchronize()
print(f"resident loop: {1e3*(time.perf_counter()-t0)/4:.2f} ms per step")
