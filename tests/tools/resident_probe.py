import sys, time, os, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mgcod_b200 import synth
from mgcod_b200.engine import Engine, synth_metagenome
d = tempfile.mkdtemp()
synth.write_model_set(d)
bp = 1_000_000_000
pinned = torch.empty(bp + 600_000, dtype=torch.uint8, pin_memory=True)
buf, offs, _ = synth_metagenome(20260101, bp, threads=16, out=pinned.numpy())
eng = Engine([f"{d}/mgm_{c}.mod" for c in (11, 4, 15, 101)])
for _ in range(3):
    eng.decode_buffer(buf, offs)
eng.upload(buf, offs)
for sleep in (0.0, 0.006):
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        a = time.perf_counter(); res = eng.decode_resident(); b = time.perf_counter(); del res
        ts.append(1e3 * (b - a))
        if sleep: time.sleep(sleep)
    torch.cuda.synchronize()
    print(f"sleep {sleep}: decode_resident " + " ".join(f"{t:.2f}" for t in ts))
