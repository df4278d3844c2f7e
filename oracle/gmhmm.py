"""ctypes binding of oracle/liboracle.so (plain-C restatement of the reference gene finder).

TEST INFRASTRUCTURE — see the header of gmhmm_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Cand(C.Structure):
    _fields_ = [("left", C.c_long), ("right", C.c_long), ("strand", C.c_int), ("len", C.c_int),
                ("gc", C.c_int), ("domain", C.c_int), ("partial_start", C.c_int), ("partial_stop", C.c_int),
                ("total", C.c_double), ("cum", C.c_double), ("prev", C.c_long)]


class Gene(C.Structure):
    _fields_ = [("left", C.c_long), ("right", C.c_long), ("strand", C.c_int), ("domain", C.c_int),
                ("gc", C.c_int), ("partial_left", C.c_int), ("partial_right", C.c_int), ("logodd", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gmhmm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_model_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_model_gcode.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_candidates.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.POINTER(C.POINTER(Cand))]
        L.orc_candidates.restype = C.c_long
        L.orc_chain.argtypes = [C.POINTER(Cand), C.c_long, C.POINTER(C.c_long), C.POINTER(C.c_double)]
        L.orc_chain.restype = C.c_long
        L.orc_decode.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.POINTER(C.POINTER(Gene)), C.POINTER(C.c_double)]
        L.orc_decode.restype = C.c_long
        L.orc_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class Model:
    def __init__(self, path: str):
        self._h = C.c_void_p()
        if lib().orc_model_load(path.encode(), C.byref(self._h)) != 0:
            raise RuntimeError(lib().orc_last_error().decode())
        self.gcode = lib().orc_model_gcode(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_model_free(self._h)
            self._h = None

    def candidates(self, seq: str | bytes, with_chain: bool = True):
        """All candidates with total > 0 as dicts (0-based coordinates); ``cum``/``prev`` filled."""
        if isinstance(seq, str):
            seq = seq.encode()
        p = C.POINTER(Cand)()
        n = lib().orc_candidates(self._h, seq, len(seq), C.byref(p))
        chain = (C.c_long * (n + 1))()
        tot = C.c_double()
        ng = lib().orc_chain(p, n, chain, C.byref(tot)) if with_chain else 0
        out = [{f: getattr(p[i], f) for f, _ in Cand._fields_} for i in range(n)]
        lib().orc_free(p)
        return out, [chain[i] for i in range(ng)], tot.value

    def decode(self, seq: str | bytes):
        """→ (genes, total_logodd); genes as dicts with 1-based GFF coordinates."""
        if isinstance(seq, str):
            seq = seq.encode()
        p = C.POINTER(Gene)()
        tot = C.c_double()
        n = lib().orc_decode(self._h, seq, len(seq), C.byref(p), C.byref(tot))
        out = [{f: getattr(p[i], f) for f, _ in Gene._fields_} for i in range(n)]
        lib().orc_free(p)
        return out, tot.value
