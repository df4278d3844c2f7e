"""Synthetic metagenome of BASELINE.json config 5 — a workload generator, not product code.

ctypes face of workload/libmgc_workload.so (built from synth_workload.cpp by workload/Makefile; host only, no CUDA).  Kept out
of the product library so that bench.py's reference arm, the golden-file scripts and the CPU tests can make their inputs
without loading mgcod_b200/libmgcod_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so, src = os.path.join(_HERE, "libmgc_workload.so"), os.path.join(_HERE, "synth_workload.cpp")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
            subprocess.check_call(["make", "-C", _HERE, "libmgc_workload.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        i64, i32 = C.c_int64, C.c_int32
        L.wl_synth_lengths.argtypes = [C.c_uint64, i64, i32, i32, C.POINTER(i64), i64]
        L.wl_synth_lengths.restype = i64
        L.wl_synth_contigs.argtypes = [C.c_uint64, C.POINTER(i64), i64, C.c_void_p, C.POINTER(i32), C.c_int]
        _LIB = L
    return _LIB


def synth_metagenome(seed: int, total_bp: int, min_len: int = 5000, max_len: int = 500000, threads: int = 8,
                     out: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Lengths ~ L^-2 on [min_len, max_len] until ``total_bp``; gene-like content with a per-contig GC regime and true
    genetic code.  → (ascii buffer uint8, offsets int64[n+1], truth_code int32[n]).  ``out``: write into this buffer
    (e.g. the numpy view of a page-locked torch tensor)."""
    L = lib()
    cap = total_bp // min_len + 16
    lens = np.zeros(cap, dtype=np.int64)
    n = L.wl_synth_lengths(seed, total_bp, min_len, max_len, lens.ctypes.data_as(C.POINTER(C.c_int64)), cap)
    if n < 0:
        raise ValueError("bad arguments for the synthetic workload")
    lens = lens[:n].copy()
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    buf = out if out is not None else np.empty(int(offs[-1]), dtype=np.uint8)
    if buf.size < offs[-1]:
        raise ValueError("output buffer too small")
    truth = np.zeros(n, dtype=np.int32)
    if L.wl_synth_contigs(seed, lens.ctypes.data_as(C.POINTER(C.c_int64)), n, C.c_void_p(buf.ctypes.data),
                          truth.ctypes.data_as(C.POINTER(C.c_int32)), threads) != 0:
        raise ValueError("bad arguments for the synthetic workload")
    return buf[:int(offs[-1])], offs, truth
