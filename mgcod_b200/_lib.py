"""ctypes binding of the in-tree shared library ``libmgcod_b200.so`` (C-ABI: include/mgcod_b200.h).

The library is built for sm_100a by ``mgcod_b200/csrc/Makefile`` (``__graft_entry__.build()``).  There is
no CPU fallback: if the library is missing the import fails loudly; if no B200 is visible,
``mgc_engine_create`` fails with MGC_E_CUDA and :class:`~mgcod_b200.engine.Engine` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MGCOD_B200_LIB") or os.path.join(_HERE, "libmgcod_b200.so")   # override: A/B runs of two builds

MGC_MAX_MODELS = 8
MGC_E_ARG, MGC_E_IO, MGC_E_PARSE, MGC_E_CUDA, MGC_E_ALPHABET, MGC_E_NOMEM = -1, -2, -3, -4, -5, -6


class Gene(C.Structure):
    _fields_ = [("contig", C.c_int32), ("left", C.c_int32), ("right", C.c_int32), ("strand", C.c_int8),
                ("domain", C.c_int8), ("partial_left", C.c_int8), ("partial_right", C.c_int8), ("gc", C.c_int32),
                ("logodd", C.c_double)]


class Candidate(C.Structure):
    _fields_ = [("contig", C.c_int32), ("left", C.c_int32), ("right", C.c_int32), ("strand", C.c_int32),
                ("gc", C.c_int32), ("domain", C.c_int32), ("partial_start", C.c_int32), ("partial_stop", C.c_int32),
                ("total", C.c_double), ("cum", C.c_double)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d", "pack", "orf_find", "orf_prep", "orf_score", "orf_final", "order", "chain", "aggregate", "d2h", "total")] + \
               [(n, C.c_int64) for n in ("n_orfs", "n_starts", "n_walks", "n_candidates", "n_genes")] + [("kernel_launches", C.c_int32)]


#: every symbol include/mgcod_b200.h declares
EXPORTS = (
    "mgc_version", "mgc_last_error", "mgc_device_count", "mgc_model_load", "mgc_model_free", "mgc_model_info",
    "mgc_engine_create", "mgc_engine_destroy", "mgc_decode_batch", "mgc_upload_batch", "mgc_prefetch_batch", "mgc_decode_resident",
    "mgc_engine_stats", "mgc_result_num_genes", "mgc_result_genes", "mgc_result_contig_totals",
    "mgc_result_model_cents", "mgc_result_min_gene_length", "mgc_result_set_min_gene_length", "mgc_result_free", "mgc_result_create", "mgc_recoded_stops", "mgc_window_counts", "mgc_window_labels", "mgc_isoform_labels", "mgc_merge_labels", "mgc_debug_candidates",
    "mgc_write_gff", "mgc_engine_device_genes",
    "mgc_fasta_read", "mgc_fasta_num_records", "mgc_fasta_id", "mgc_fasta_defline", "mgc_fasta_ids_joined", "mgc_fasta_id_array", "mgc_fasta_sequence", "mgc_fasta_offsets", "mgc_fasta_free",
)

_LIB = None


def build(force: bool = False) -> str:
    """Compile the library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.check_call(["make", "-C", src_dir], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(mgcod_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, cp = C.c_void_p, C.c_int32, C.c_int64, C.c_char_p
    L.mgc_version.restype = cp
    L.mgc_last_error.restype = cp
    L.mgc_device_count.restype = C.c_int
    L.mgc_model_load.argtypes = [cp, C.POINTER(vp)]
    L.mgc_model_free.argtypes = [vp]
    L.mgc_model_info.argtypes = [vp, C.POINTER(i32)]
    L.mgc_engine_create.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]
    L.mgc_engine_destroy.argtypes = [vp]
    L.mgc_decode_batch.argtypes = [vp, vp, C.POINTER(i64), i32, C.POINTER(vp)]
    L.mgc_upload_batch.argtypes = [vp, vp, C.POINTER(i64), i32]
    L.mgc_prefetch_batch.argtypes = [vp, vp, C.POINTER(i64), i32]
    L.mgc_decode_resident.argtypes = [vp, C.POINTER(vp)]
    L.mgc_engine_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mgc_result_num_genes.argtypes = [vp, C.c_int]
    L.mgc_result_num_genes.restype = i64
    L.mgc_result_genes.argtypes = [vp, C.c_int]
    L.mgc_result_genes.restype = C.POINTER(Gene)
    L.mgc_result_contig_totals.argtypes = [vp, C.c_int]
    L.mgc_result_contig_totals.restype = C.POINTER(C.c_double)
    L.mgc_result_model_cents.argtypes = [vp, C.POINTER(i64)]
    L.mgc_result_min_gene_length.argtypes = [vp, C.c_int]
    L.mgc_result_min_gene_length.restype = i32
    L.mgc_result_set_min_gene_length.argtypes = [vp, C.c_int, i32]
    L.mgc_result_free.argtypes = [vp]
    L.mgc_result_create.argtypes = [C.c_int, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(vp)]
    L.mgc_recoded_stops.argtypes = [vp, vp, i64, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), i64]
    L.mgc_recoded_stops.restype = i64
    L.mgc_window_counts.argtypes = [C.POINTER(i64), C.POINTER(i32), i32, i32, i32, C.c_int, C.POINTER(i64)]
    L.mgc_window_labels.argtypes = [vp, C.POINTER(i32), i32, i32, i32, C.c_int, C.POINTER(i64), C.POINTER(i32), C.POINTER(i64)]
    L.mgc_isoform_labels.argtypes = [vp, C.POINTER(i32), i32, i32, i32, C.c_int, C.POINTER(i64), C.POINTER(i32), C.POINTER(i64),
                                     C.POINTER(i32), C.POINTER(i32)]
    L.mgc_merge_labels.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), i32, i32, C.c_uint32, C.POINTER(i32), C.POINTER(i32)]
    L.mgc_debug_candidates.argtypes = [vp, C.c_int, C.POINTER(Candidate), i64]
    L.mgc_debug_candidates.restype = i64
    L.mgc_write_gff.argtypes = [vp, C.c_int, C.POINTER(cp), C.POINTER(i64), i32, cp, cp, C.c_int, cp, cp]
    L.mgc_engine_device_genes.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    L.mgc_fasta_read.argtypes = [cp, C.c_int, C.POINTER(vp)]
    L.mgc_fasta_num_records.argtypes = [vp]
    L.mgc_fasta_num_records.restype = i32
    L.mgc_fasta_id.argtypes = [vp, i32]
    L.mgc_fasta_id.restype = cp
    L.mgc_fasta_defline.argtypes = [vp, i32]
    L.mgc_fasta_defline.restype = cp
    L.mgc_fasta_ids_joined.argtypes = [vp, C.c_int, C.POINTER(i64)]
    L.mgc_fasta_ids_joined.restype = vp
    L.mgc_fasta_id_array.argtypes = [vp]
    L.mgc_fasta_id_array.restype = vp
    L.mgc_fasta_sequence.argtypes = [vp]
    L.mgc_fasta_sequence.restype = vp
    L.mgc_fasta_offsets.argtypes = [vp]
    L.mgc_fasta_offsets.restype = C.POINTER(i64)
    L.mgc_fasta_free.argtypes = [vp]
    _LIB = L
    return L


def last_error() -> str:
    return lib().mgc_last_error().decode(errors="replace")
