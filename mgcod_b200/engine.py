"""Python face of the C-ABI (include/mgcod_b200.h): model files, the per-GPU engine and decode results.

This is the in-process replacement for the reference's process boundary
``subprocess.check_call([gmhmmp2, '-M', mgm_<code>.mod, '-s', fasta, '-f', 'gff', '-o', ...])``
(reference bin/data.py:39-54): one :class:`Engine` holds the four parameter files on the GPU and decodes
whole batches of contigs under all of them in one call.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Iterator, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Candidate, Gene, Stats

GENE_DTYPE = np.dtype([("contig", "<i4"), ("left", "<i4"), ("right", "<i4"), ("strand", "i1"), ("domain", "i1"),
                       ("partial_left", "i1"), ("partial_right", "i1"), ("gc", "<i4"), ("_pad", "<i4"), ("logodd", "<f8")])
assert GENE_DTYPE.itemsize == C.sizeof(Gene) == 32
CAND_DTYPE = np.dtype([("contig", "<i4"), ("left", "<i4"), ("right", "<i4"), ("strand", "<i4"), ("gc", "<i4"),
                       ("domain", "<i4"), ("partial_start", "<i4"), ("partial_stop", "<i4"), ("total", "<f8"), ("cum", "<f8")])
assert CAND_DTYPE.itemsize == C.sizeof(Candidate)

#: label order used everywhere in Mgcod (reference bin/pipeline.py:53-54)
MGCOD_CODES = (11, 4, 15, 101)


class MgcodError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _check(rc: int) -> None:
    if rc != 0:
        raise MgcodError(rc, _lib.last_error())


class ModelFile:
    """A parsed ``.mod`` parameter file (host side; no GPU needed)."""

    def __init__(self, path: str):
        self.path = str(path)
        self._h = C.c_void_p()
        _check(_lib.lib().mgc_model_load(self.path.encode(), C.byref(self._h)))
        info = (C.c_int32 * 6)()
        _check(_lib.lib().mgc_model_info(self._h, info))
        self.gcode, self.cod_order, self.non_order, self.gene_min_length, self.n_bacterial, self.n_archaeal = list(info)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "lib", None):   # None during interpreter shutdown
            _lib.lib().mgc_model_free(self._h)
            self._h = None


def pack_sequences(seqs: Sequence[bytes | str]) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate sequences into one uint8 buffer + int64 offsets (the layout mgc_decode_batch takes)."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bs], out=offs[1:])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8) if offs[-1] else np.zeros(0, np.uint8)
    return buf, offs


class Result:
    """Genes of one decoded batch for every model of the engine."""

    def __init__(self, handle: C.c_void_p, n_models: int, n_contigs: int):
        self._h = handle
        self.n_models = n_models
        self.n_contigs = n_contigs

    @classmethod
    def from_genes(cls, genes_per_model: Sequence[np.ndarray], totals_per_model: Sequence[np.ndarray],
                   min_gene_length: Optional[Sequence[int]] = None) -> "Result":
        """Host-only result (no GPU), e.g. from cached GFF files; arrays of GENE_DTYPE ordered by (contig, left).
        min_gene_length: $GENE_MIN_LENGTH per model (default 90), only used by write() for gene-less records."""
        M = len(genes_per_model)
        gs = [np.ascontiguousarray(g, dtype=GENE_DTYPE) for g in genes_per_model]
        ts = [np.ascontiguousarray(t, dtype=np.float64) for t in totals_per_model]
        n_contigs = len(ts[0])
        gp = (C.c_void_p * M)(*[g.ctypes.data for g in gs])
        tp = (C.c_void_p * M)(*[t.ctypes.data for t in ts])
        counts = (C.c_int64 * M)(*[len(g) for g in gs])
        h = C.c_void_p()
        _check(_lib.lib().mgc_result_create(M, n_contigs, gp, counts, tp, C.byref(h)))
        for m, v in enumerate(min_gene_length or ()):
            _check(_lib.lib().mgc_result_set_min_gene_length(h, m, int(v)))
        return cls(h, M, n_contigs)

    def num_genes(self, model: int) -> int:
        return int(_lib.lib().mgc_result_num_genes(self._h, model))

    def genes(self, model: int) -> np.ndarray:
        """Structured array (GENE_DTYPE), ordered by (contig, left); 1-based inclusive coordinates."""
        L = _lib.lib()
        n = L.mgc_result_num_genes(self._h, model)
        if n < 0:
            raise IndexError(model)
        if n == 0:
            return np.zeros(0, GENE_DTYPE)
        p = L.mgc_result_genes(self._h, model)
        # copied as plain bytes (numpy copies a structured array field by field, 10x slower), then viewed as rows
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n * 32,)).copy().view(GENE_DTYPE)

    def contig_totals(self, model: int) -> np.ndarray:
        p = _lib.lib().mgc_result_contig_totals(self._h, model)
        return np.ctypeslib.as_array(p, shape=(self.n_contigs,)).copy()

    def model_cents(self) -> np.ndarray:
        out = np.zeros(self.n_models, dtype=np.int64)
        _check(_lib.lib().mgc_result_model_cents(self._h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def write(self, model: int, contig_ids, contig_len: Sequence[int], seq_file: str, mod_file: str,
              gcode: int, out_path: str, fmt: str = "gff") -> None:
        """``contig_ids``: the ids as strings, or a ready array of C strings (``data.Fasta.id_array()``: the reader's own
        storage — four writer threads encoding 20 000 ids each would hold the interpreter lock against the caller)."""
        lens = np.asarray(contig_len, dtype=np.int64)
        if isinstance(contig_ids, C.c_void_p):
            ids = C.cast(contig_ids, C.POINTER(C.c_char_p))
        else:
            ids = (C.c_char_p * len(contig_ids))(*[s.encode() for s in contig_ids])
        _check(_lib.lib().mgc_write_gff(self._h, model, ids, lens.ctypes.data_as(C.POINTER(C.c_int64)), len(lens),
                                        seq_file.encode(), mod_file.encode(), gcode, out_path.encode(), fmt.encode()))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "lib", None):   # None during interpreter shutdown
            _lib.lib().mgc_result_free(self._h)
            self._h = None


class DeviceGenes:
    """n gene rows of 32 bytes in device memory (see :meth:`Engine.device_genes`)."""

    def __init__(self, ptr: int, n: int):
        self.ptr, self.n = ptr, n
        self.__cuda_array_interface__ = {"shape": (n * 32,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}

    def __len__(self):
        return self.n


class Engine:
    """One per process per GPU.  ``models`` in label order (Mgcod: codes 11, 4, 15, 101)."""

    def __init__(self, models: Sequence[ModelFile | str], device: int = 0):
        self.models = [m if isinstance(m, ModelFile) else ModelFile(m) for m in models]
        arr = (C.c_void_p * len(self.models))(*[m._h for m in self.models])
        self._h = C.c_void_p()
        _check(_lib.lib().mgc_engine_create(device, arr, len(self.models), C.byref(self._h)))
        self.device = device
        self.n_contigs = 0
        self._keep = None

    @property
    def n_models(self) -> int:
        return len(self.models)

    @staticmethod
    def _checked(buf: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        """the letter buffer must be one C-contiguous run of bytes that covers the offsets (its address crosses the C-ABI)"""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        if not isinstance(buf, np.ndarray) or buf.dtype.itemsize != 1 or not buf.flags.c_contiguous:
            raise TypeError("sequence buffer must be a C-contiguous numpy array of bytes (uint8 / int8 / S1)")
        if len(offsets) < 2 or offsets[0] < 0 or (np.diff(offsets) < 0).any() or offsets[-1] > buf.size:
            raise ValueError("offsets must ascend, start at >= 0 and end inside the buffer")
        return offsets

    def upload(self, buf: np.ndarray, offsets: np.ndarray) -> None:
        """Host→device copy of a batch (``buf`` may be a pinned torch tensor's numpy view)."""
        offsets = self._checked(buf, offsets)
        self._keep = (buf, offsets)
        self.n_contigs = len(offsets) - 1
        _check(_lib.lib().mgc_upload_batch(self._h, C.c_void_p(buf.ctypes.data), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                           self.n_contigs))

    def prefetch(self, buf: np.ndarray, offsets: np.ndarray) -> None:
        """Starts copying the NEXT batch to the device on a second stream (returns at once); the following ``upload`` /
        ``decode_buffer`` of the same arrays adopts it.  ``buf`` should be page-locked and must not change until then."""
        offsets = self._checked(buf, offsets)
        self._keep_next = (buf, offsets)
        _check(_lib.lib().mgc_prefetch_batch(self._h, C.c_void_p(buf.ctypes.data), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                             len(offsets) - 1))

    def decode_stream(self, batches: Iterable[tuple[np.ndarray, np.ndarray]]) -> Iterator[Result]:
        """Decodes a sequence of (buf, offsets) batches, copying batch i+1 to the device while batch i is decoded —
        what bin/multiprocess_mgcod.py:140-145 gets from running several genomes' processes side by side."""
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        self.upload(*cur)
        while cur is not None:
            nxt = next(it, None)
            if nxt is not None:
                nxt = (nxt[0], np.ascontiguousarray(nxt[1], dtype=np.int64))
                self.prefetch(*nxt)
            yield self.decode_resident()
            if nxt is not None:
                self.upload(*nxt)
            cur = nxt

    MAX_BATCH_BP = 1_800_000_000     # one device batch (mgc_decode_batch accepts ~2e9 letters; 70 GB of scratch per Gbp x 4 models)

    def decode_large(self, buf: np.ndarray, offsets: np.ndarray, max_bp: Optional[int] = None) -> Result:
        """Decodes inputs of any size: contigs are cut into batches of at most ``max_bp`` letters (whole contigs), streamed
        through the device with ``decode_stream``, and the gene tables are merged on the host.  The returned result is
        host-only (its genes are not resident on the device: window labels / recoded stops need a single batch)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        max_bp = int(max_bp or self.MAX_BATCH_BP)
        cuts = [0]
        while cuts[-1] < n:
            a = cuts[-1]
            b = int(np.searchsorted(offsets, offsets[a] + max_bp, side="right")) - 1
            b = min(max(b, a + 1), n)
            if offsets[b] - offsets[a] > 0x7ffffff0 * 0.93:
                raise MgcodError(_lib.MGC_E_ARG, f"contig {a} alone exceeds the size of one device batch")
            cuts.append(b)
        if len(cuts) == 2:
            return self.decode_buffer(buf, offsets)
        genes = [[] for _ in range(self.n_models)]
        totals = [[] for _ in range(self.n_models)]
        batches = [(buf, offsets[cuts[i]:cuts[i + 1] + 1]) for i in range(len(cuts) - 1)]
        for i, res in enumerate(self.decode_stream(batches)):
            for m in range(self.n_models):
                g = res.genes(m)
                g["contig"] += cuts[i]
                genes[m].append(g)
                totals[m].append(res.contig_totals(m))
        min_len = [int(_lib.lib().mgc_result_min_gene_length(res._h, m)) for m in range(self.n_models)]
        return Result.from_genes([np.concatenate(g) for g in genes], [np.concatenate(t) for t in totals], min_len)

    def decode_resident(self) -> Result:
        r = C.c_void_p()
        _check(_lib.lib().mgc_decode_resident(self._h, C.byref(r)))
        return Result(r, self.n_models, self.n_contigs)

    def decode(self, seqs: Sequence[bytes | str]) -> Result:
        buf, offs = pack_sequences(seqs)
        return self.decode_buffer(buf, offs)

    def decode_buffer(self, buf: np.ndarray, offsets: np.ndarray) -> Result:
        offsets = self._checked(buf, offsets)
        self.n_contigs = len(offsets) - 1
        r = C.c_void_p()
        _check(_lib.lib().mgc_decode_batch(self._h, C.c_void_p(buf.ctypes.data), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                           self.n_contigs, C.byref(r)))
        return Result(r, self.n_models, self.n_contigs)

    def device_genes(self, model: int) -> "DeviceGenes":
        """The gene table of the last decode where it sits in device memory (GENE_DTYPE rows as bytes), exposed through
        ``__cuda_array_interface__`` so that ``torch.as_tensor(x, device='cuda')`` wraps it without a copy — the send buffer of
        the multi-GPU gather (:func:`mgcod_b200.shard.gather_genes_device`).  Valid until the next decode on this engine."""
        p, n = C.c_void_p(), C.c_int64()
        _check(_lib.lib().mgc_engine_device_genes(self._h, model, C.byref(p), C.byref(n)))
        return DeviceGenes(p.value or 0, int(n.value))

    def stats(self) -> dict:
        s = Stats()
        _check(_lib.lib().mgc_engine_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def candidates(self, model: int) -> np.ndarray:
        """Testing hook: pruned candidates (total > 0) of the last decode with their chain scores."""
        L = _lib.lib()
        n = L.mgc_debug_candidates(self._h, model, None, 0)
        if n < 0:
            raise MgcodError(int(n), _lib.last_error())
        out = np.zeros(n, CAND_DTYPE)
        if n:
            L.mgc_debug_candidates(self._h, model, out.ctypes.data_as(C.POINTER(Candidate)), n)
        return out

    def recoded_stops(self, genes: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """In-frame stop codons inside ``genes`` (GENE_DTYPE rows of the last decoded batch), on the device
        (reference bin/results.py:253-294).  → (counts[n, 3] for TAA/TAG/TGA, offsets[n*3+1], codon indices)."""
        L = _lib.lib()
        g = np.ascontiguousarray(genes, dtype=GENE_DTYPE)
        n = len(g)
        counts = np.zeros((max(n, 1), 3), dtype=np.int32)
        offs = np.zeros(n * 3 + 1, dtype=np.int64)
        args = (self._h, C.c_void_p(g.ctypes.data), n, counts.ctypes.data_as(C.POINTER(C.c_int32)), offs.ctypes.data_as(C.POINTER(C.c_int64)))
        total = L.mgc_recoded_stops(*args, None, 0)
        if total < 0:
            raise MgcodError(int(total), _lib.last_error())
        pos = np.zeros(max(int(total), 1), dtype=np.int32)
        if total:
            total = L.mgc_recoded_stops(*args, pos.ctypes.data_as(C.POINTER(C.c_int32)), int(total))
            if total < 0:
                raise MgcodError(int(total), _lib.last_error())
        return counts[:n], offs, pos[:int(total)]

    def window_labels(self, contig_len: Sequence[int], stride: Sequence[int], window: int, consecutive: int,
                      short_contigs: bool) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """→ (win_offsets[n+1], labels[nwin], cents[nwin, n_models]) for the last decoded batch
        (reference bin/pipeline.py:157-385)."""
        L = _lib.lib()
        lens = np.asarray(contig_len, dtype=np.int64)
        st = np.asarray(stride, dtype=np.int32)
        offs = np.zeros(len(lens) + 1, dtype=np.int64)
        _check(L.mgc_window_counts(lens.ctypes.data_as(C.POINTER(C.c_int64)), st.ctypes.data_as(C.POINTER(C.c_int32)), len(lens),
                                   window, consecutive, int(short_contigs), offs.ctypes.data_as(C.POINTER(C.c_int64))))
        nw = int(offs[-1])
        labels = np.zeros(max(nw, 1), dtype=np.int32)
        cents = np.zeros((max(nw, 1), self.n_models), dtype=np.int64)
        _check(L.mgc_window_labels(self._h, st.ctypes.data_as(C.POINTER(C.c_int32)), len(lens), window, consecutive, int(short_contigs),
                                   offs.ctypes.data_as(C.POINTER(C.c_int64)), labels.ctypes.data_as(C.POINTER(C.c_int32)),
                                   cents.ctypes.data_as(C.POINTER(C.c_int64))))
        return offs, labels[:nw], cents[:nw]

    def isoform_labels(self, contig_len: Sequence[int], stride: Sequence[int], window: int, consecutive: int,
                       short_contigs: bool) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Window sums, argmax labels and their smoothing for the last decoded batch in one device call (reference
        bin/pipeline.py:157-385, then :98-110 with :1064-1168): → (win_offsets[n+1], labels[nwin], cents[nwin, n_models],
        merged[nwin], status[n]).  Contigs whose status is 1 (the reference raises IndexError there) keep their input labels."""
        L = _lib.lib()
        lens = np.ascontiguousarray(contig_len, dtype=np.int64)
        st = np.ascontiguousarray(stride, dtype=np.int32)
        n = len(lens)
        offs = np.zeros(n + 1, dtype=np.int64)
        _check(L.mgc_window_counts(lens.ctypes.data_as(C.POINTER(C.c_int64)), st.ctypes.data_as(C.POINTER(C.c_int32)), n,
                                   window, consecutive, int(short_contigs), offs.ctypes.data_as(C.POINTER(C.c_int64))))
        nw = int(offs[-1])
        labels = np.empty(max(nw, 1), dtype=np.int32)
        merged = np.empty(max(nw, 1), dtype=np.int32)
        cents = np.empty((max(nw, 1), self.n_models), dtype=np.int64)
        status = np.zeros(max(n, 1), dtype=np.int32)
        i32p = C.POINTER(C.c_int32)
        _check(L.mgc_isoform_labels(self._h, st.ctypes.data_as(i32p), n, window, consecutive, int(short_contigs),
                                    offs.ctypes.data_as(C.POINTER(C.c_int64)), labels.ctypes.data_as(i32p),
                                    cents.ctypes.data_as(C.POINTER(C.c_int64)), merged.ctypes.data_as(i32p), status.ctypes.data_as(i32p)))
        labels, merged, cents, status = labels[:nw], merged[:nw], cents[:nw], status[:n]
        for c in np.flatnonzero(status):
            merged[offs[c]:offs[c + 1]] = labels[offs[c]:offs[c + 1]]
        return offs, labels, cents, merged, status

    def merge_labels_flat(self, flat: np.ndarray, offs: np.ndarray, consecutive: int) -> tuple[np.ndarray, np.ndarray]:
        """Label smoothing of all contigs in one device call on the flat layout ``window_labels`` returns.
        → (merged labels, status per contig: 0 ok, 1 = the reference raises IndexError there; those keep their input)."""
        from .labels import noise_labels
        flat = np.ascontiguousarray(flat, dtype=np.int32)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        n = len(offs) - 1
        mask = 0
        for l in noise_labels(flat) if flat.size else []:
            mask |= 1 << int(l)
        out = np.zeros(max(int(offs[-1]), 1), dtype=np.int32)
        status = np.zeros(max(n, 1), dtype=np.int32)
        _check(_lib.lib().mgc_merge_labels(self._h, flat.ctypes.data_as(C.POINTER(C.c_int32)), offs.ctypes.data_as(C.POINTER(C.c_int64)),
                                           n, consecutive, mask, out.ctypes.data_as(C.POINTER(C.c_int32)),
                                           status.ctypes.data_as(C.POINTER(C.c_int32))))
        out = out[:int(offs[-1])]
        status = status[:n]
        for c in np.flatnonzero(status):
            out[offs[c]:offs[c + 1]] = flat[offs[c]:offs[c + 1]]
        return out, status

    def merge_labels(self, labels_per_contig: Sequence[np.ndarray], consecutive: int, strict: bool = True) -> list[np.ndarray]:
        """Device-side label smoothing (reference bin/pipeline.py:1064-1168) of every contig's window labels; noise labels =
        all but the two most frequent labels over the whole input.  Raises IndexError where the reference does
        (``strict=False``: such contigs keep their input labels, for bulk runs)."""
        labs = [np.ascontiguousarray(l, dtype=np.int32) for l in labels_per_contig]
        offs = np.zeros(len(labs) + 1, dtype=np.int64)
        np.cumsum([len(l) for l in labs], out=offs[1:])
        flat = np.concatenate(labs) if labs else np.zeros(0, np.int32)
        out, status = self.merge_labels_flat(flat, offs, consecutive)
        if strict and (status != 0).any():
            raise IndexError(f"label smoothing walks off the array for contig index {int(np.flatnonzero(status)[0])} (as the reference does)")
        return [out[offs[i]:offs[i + 1]].astype(np.int64) for i in range(len(labs))]

    def close(self) -> None:
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "lib", None):   # None during interpreter shutdown
            _lib.lib().mgc_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def synth_metagenome(*args, **kw):
    """The synthetic workload of BASELINE.json config 5 — lives in the host-only ``workload`` package (its own small
    library), re-exported here for the tests and tools that have always imported it from the engine module."""
    from workload import synth_metagenome as _synth
    return _synth(*args, **kw)
