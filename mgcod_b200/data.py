"""``Data`` — drop-in for the reference's boundary class (bin/data.py:16-124), backed by the CUDA engine.

Same constructor, same three methods, same DataFrame schema; what changes is *how* ``run_mgm`` gets its
genes: instead of ``subprocess.check_call([gmhmmp2, '-M', mgm_<code>.mod, '-s', fasta, '-f', 'gff', '-o', ...])``
four times (bin/data.py:39-54) one :class:`mgcod_b200.engine.Engine` decodes every contig of the FASTA under
the four parameter files in a single batch on the GPU.  The four ``mgm_<code>_<species>.gff`` files are still
written in gmhmmp2's text format (they are documented output, README.md:65, and an on-disk cache: a model whose
file already exists is not rewritten, bin/data.py:39,43,47,51).
"""
from __future__ import annotations

import logging
import os

import numpy as np
import pandas as pd

from . import _lib
from .engine import MGCOD_CODES, Engine, ModelFile

GFF_COLUMNS = ["sequence", "source", "feature", "start", "end", "logodd", "strand", "phase", "attributes"]


def rows_by_contig(df) -> dict:
    """contig id -> positions of its rows in ``df`` (frame order), found in one pass: the reference filters the whole frame
    with ``df[df.contig == id]`` once per contig, which is quadratic for a metagenome"""
    return df.groupby("contig", sort=False).indices if len(df) else {}


def contig_rows(df, index: dict, contig):
    """the rows of one contig (what ``df[df.contig == contig]`` returns), through :func:`rows_by_contig`"""
    pos = index.get(contig)
    return df.iloc[:0] if pos is None else df.iloc[pos]


def cents_of(logodd: np.ndarray) -> np.ndarray:
    """the %.2f value of each log-odds score in 1/100 units, as printf rounds it (exact decimal value of the double, ties
    to even): rint(v * 100) except within 1e-6 of a rounding boundary, where the formatted text decides — the rule of the
    C writer (mgc_host.cpp TextOut::fixed2), so frames built from arrays equal frames parsed from the GFF files"""
    a = np.abs(logodd) * 100.0
    c = np.rint(a)
    near = np.flatnonzero(np.abs(a - np.floor(a) - 0.5) < 1e-6)
    for i in near:
        c[i] = round(abs(float("%.2f" % logodd[i])) * 100.0)
    return (np.sign(logodd) * c).astype(np.int64)


class GeneFrames:
    """label -> DataFrame of one model's genes in the reference's schema (bin/data.py:66-106 + the ``contig`` column of
    bin/pipeline.py:68-76), built from the decoded gene arrays instead of re-reading the GFF files the run has just
    written, and only when somebody asks: a metagenome run in --isoforms mode touches the rows of the few contigs with a
    genetic-code switch, not 5 million attribute strings.  Behaves like the reference's dict of four frames."""

    COLUMNS = GFF_COLUMNS + ["length", "contig"]

    def __init__(self, genes_per_model, ids, codes=MGCOD_CODES):
        self.genes = list(genes_per_model)            # GENE_DTYPE, (contig, left) order, contig = index into ids
        self.ids = np.asarray(list(ids), dtype=object)
        self.codes = tuple(codes)
        n = len(self.ids)
        # row range of every contig (the contig field is a strided view of the 32-byte rows: made contiguous for the search)
        self.first = [np.searchsorted(np.ascontiguousarray(g["contig"]), np.arange(n + 1)) for g in self.genes]
        self._index_of = None
        self._full = {}

    # -- dict protocol (the reference iterates .items() / .values() and indexes by label)
    def __len__(self):
        return len(self.genes)

    def __iter__(self):
        return iter(range(len(self.genes)))

    def keys(self):
        return range(len(self.genes))

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __contains__(self, k):
        return isinstance(k, (int, np.integer)) and 0 <= k < len(self.genes)

    def __getitem__(self, lab):
        lab = int(lab)
        if lab not in self._full:
            self._full[lab] = self.rows(lab, 0, len(self.genes[lab]))
        return self._full[lab]

    def contig_index(self, contig_id):
        if self._index_of is None:
            self._index_of = {c: i for i, c in enumerate(self.ids.tolist())}
        return self._index_of.get(contig_id)

    def contig_range(self, lab, contig_id):
        c = self.contig_index(contig_id)
        if c is None:
            return 0, 0
        f = self.first[int(lab)]
        return int(f[c]), int(f[c + 1])

    def contig_frame(self, lab, contig_id):
        """the rows of one contig — what ``df[df.contig == contig_id]`` gives in the reference"""
        a, b = self.contig_range(lab, contig_id)
        if int(lab) in self._full:
            return self._full[int(lab)].iloc[a:b]
        return self.rows(lab, a, b)

    def contig_columns(self, lab, contig_id) -> dict:
        """the rows of one contig as a dict of columns (same names, order and values as the frame's) — for consumers that
        turn the rows into Python objects anyway (the isoform comparison)"""
        a, b = self.contig_range(lab, contig_id)
        return self.columns(lab, a, b)

    def columns(self, lab, a, b) -> dict:
        g = self.genes[int(lab)][a:b]
        code = self.codes[int(lab)]
        n = len(g)
        seq = self.ids[g["contig"]].tolist() if n else []
        left, right = g["left"].astype(np.int64), g["right"].astype(np.int64)
        length = right - left + 1
        dom, pl, pr, gc = g["domain"].tolist(), g["partial_left"].tolist(), g["partial_right"].tolist(), g["gc"].tolist()
        ln = length.tolist()
        attrs = [f"gene_id {a + 1 + i}; gene_type {'archaea' if dom[i] else 'bacteria'}; "
                 + (f"partial {pl[i]}{pr[i]}; " if (pl[i] or pr[i]) else "") + f"gc {gc[i]}; length {ln[i]};" for i in range(n)]
        return {"sequence": seq, "source": [f"GeneMark.hmm2_{code}"] * n, "feature": ["CDS"] * n, "start": left, "end": right,
                "logodd": cents_of(g["logodd"]) / 100.0, "strand": np.where(g["strand"] > 0, "+", "-").tolist(),
                "phase": [0] * n, "attributes": attrs, "length": length, "contig": list(seq)}

    def rows(self, lab, a, b):
        g = self.genes[int(lab)][a:b]
        code = self.codes[int(lab)]
        n = len(g)
        seq = self.ids[g["contig"]] if n else np.zeros(0, dtype=object)
        left, right = g["left"].astype(np.int64), g["right"].astype(np.int64)
        length = right - left + 1
        dom, pl, pr, gc = g["domain"].tolist(), g["partial_left"].tolist(), g["partial_right"].tolist(), g["gc"].tolist()
        ln = length.tolist()
        attrs = np.empty(n, dtype=object)
        attrs[:] = [f"gene_id {a + 1 + i}; gene_type {'archaea' if dom[i] else 'bacteria'}; "
                    + (f"partial {pl[i]}{pr[i]}; " if (pl[i] or pr[i]) else "") + f"gc {gc[i]}; length {ln[i]};" for i in range(n)]
        df = pd.DataFrame({"sequence": pd.array(seq, dtype="str") if n else pd.array([], dtype="str"),
                           "source": pd.array(np.full(n, f"GeneMark.hmm2_{code}", dtype=object), dtype="str"),
                           "feature": pd.array(np.full(n, "CDS", dtype=object), dtype="str"),
                           "start": left, "end": right, "logodd": cents_of(g["logodd"]) / 100.0,
                           "strand": pd.array(np.where(g["strand"] > 0, "+", "-").astype(object), dtype="str"),
                           "phase": np.zeros(n, dtype=np.int64), "attributes": pd.array(attrs, dtype="str"), "length": length},
                          index=pd.RangeIndex(a, b))
        df["contig"] = list(df.sequence.values)      # as bin/pipeline.py:73-76 adds it
        return df


def frame_columns(frames) -> list:
    return list(frames.COLUMNS) if isinstance(frames, GeneFrames) else list(frames[0].columns.values)


class Record:
    """The slice of Biopython's SeqRecord that Mgcod touches: ``id``, ``description``, ``seq``, ``len()``.
    ``seq`` is a ``str``, a uint8 view into a :class:`Fasta` buffer, or — what :meth:`Fasta.records` makes for the 20 000
    contigs of a metagenome — just the record's number in its file: the view is cut and decoded to ``str`` on first use."""

    __slots__ = ("id", "description", "_seq", "_len", "_fasta", "_k")

    def __init__(self, rid: str, description: str, seq=None, fasta=None, k=0, length=None):
        self.id, self.description, self._seq = rid, description, seq
        self._fasta, self._k = fasta, k
        self._len = len(seq) if length is None else length

    @property
    def seq(self) -> str:
        if self._seq is None:
            f = self._fasta
            self._seq = f.buffer[f.offsets[self._k]:f.offsets[self._k + 1]] if f is not None and len(f.buffer) else ""
        if not isinstance(self._seq, str):
            self._seq = self._seq.tobytes().decode("ascii", errors="replace")
        return self._seq

    def detach(self):
        """the letters are about to be released: keep the sequence only if somebody has asked for it"""
        if not isinstance(self._seq, str):
            self._seq = ""
        self._fasta = None

    def __len__(self):
        return self._len


def record_ids(recs) -> list:
    """ids of a record sequence without touching the records where it is a :class:`Records`"""
    return recs.ids if isinstance(recs, Records) else [r.id for r in recs]


def record_lengths(recs) -> np.ndarray:
    return (np.asarray(recs.lengths, dtype=np.int64) if isinstance(recs, Records)
            else np.fromiter((len(r) for r in recs), dtype=np.int64, count=len(recs)))


class Records:
    """The records of a :class:`Fasta` as a read-only sequence (``len``, indexing, slicing to a list, iteration) that makes a
    :class:`Record` when one is asked for — a metagenome run touches the few contigs with a switch, not 20 000 objects.
    ``ids`` and ``lengths`` give the two columns the pipeline needs of every record without making any.
    ``order``: positions in the file, e.g. the records sorted by id (bin/data.py:122)."""

    def __init__(self, fasta, order=None):
        self._fasta = fasta
        self._order = None if order is None else np.asarray(order, dtype=np.int64)
        lens = np.diff(fasta.offsets)
        if self._order is None:
            self.ids, self._deflines, self.lengths = fasta.ids, fasta.deflines, lens
        else:
            o = self._order.tolist()
            self.ids, self._deflines, self.lengths = [fasta.ids[k] for k in o], [fasta.deflines[k] for k in o], lens[self._order]
        self._made: dict = {}

    def __len__(self):
        return len(self.ids)

    def _make(self, i):
        r = self._made.get(i)
        if r is None:
            k = i if self._order is None else int(self._order[i])
            r = self._made[i] = self._fasta._record(k)
        return r

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._make(j) for j in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._make(i)

    def __iter__(self):
        return (self._make(i) for i in range(len(self)))

    def detach(self):
        for r in self._made.values():
            r.detach()


class Fasta:
    """A FASTA file read by the native reader (``mgc_fasta_read``): one contiguous letter buffer (page-locked when a GPU is
    present) + offsets, ready for ``Engine.decode_buffer`` without a Python-side copy.  Replaces ``SeqIO.parse`` at
    bin/data.py:117 (minutes for a 1 Gbp metagenome)."""

    def __init__(self, path: str, page_locked: bool = True):
        import ctypes as C
        L = _lib.lib()
        self._h = C.c_void_p()
        rc = L.mgc_fasta_read(os.fspath(path).encode(), int(page_locked), C.byref(self._h))
        if rc != 0:
            raise OSError(_lib.last_error())
        n = L.mgc_fasta_num_records(self._h)

        def joined(which):       # all ids / deflines in one call, each followed by a newline
            ln = C.c_int64()
            p = L.mgc_fasta_ids_joined(self._h, which, C.byref(ln))
            return C.string_at(p, ln.value).decode(errors="replace").split("\n")[:n] if n else []
        self.ids = joined(0)
        self.deflines = joined(1)
        self.offsets = np.ctypeslib.as_array(L.mgc_fasta_offsets(self._h), shape=(n + 1,)).copy()
        total = int(self.offsets[-1])
        ptr = L.mgc_fasta_sequence(self._h)
        self.buffer = (np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(total, 1),))[:total]
                       if ptr else np.zeros(0, np.uint8))

    def id_array(self):
        """the ids as an array of C strings inside the reader (``Result.write`` takes it as it is); valid until close()"""
        import ctypes as C
        return C.c_void_p(_lib.lib().mgc_fasta_id_array(self._h))

    def _record(self, k: int) -> Record:
        """the one Record object of record k (shared by every :class:`Records` view of this file)"""
        made = self.__dict__.setdefault("_records_made", {})
        r = made.get(k)
        if r is None:
            r = made[k] = Record(self.ids[k], self.deflines[k], None, self, k, int(self.offsets[k + 1] - self.offsets[k]))
        return r

    def records(self, order=None) -> Records:
        return Records(self, order)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "lib", None):   # None during interpreter shutdown
            self.buffer = np.zeros(0, np.uint8)
            _lib.lib().mgc_fasta_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


_ENGINES: dict = {}


def shared_engine(model_dir: str, device: int = 0) -> Engine:
    """One engine per (process, GPU, model directory): the parameter files are parsed and uploaded once,
    not once per genome as in the reference (4 subprocesses per FASTA)."""
    key = (os.path.abspath(model_dir), device)
    if key not in _ENGINES:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(MGCOD_CODES)) as pool:      # the four parameter files are parsed side by side (no GIL inside)
            models = list(pool.map(lambda c: ModelFile(os.path.join(model_dir, f"mgm_{c}.mod")), MGCOD_CODES))
        _ENGINES[key] = Engine(models, device=device)
    return _ENGINES[key]


class Data:
    def __init__(self, path_to_genome, path_to_predictions, verbose, cwd, species, model_dir=None, device=0):
        """Same positional arguments as the reference.  ``model_dir`` defaults to ``<cwd>/dependencies``
        (where the reference keeps mgm_11/4/15/101.mod) or ``$MGCOD_MODEL_DIR``."""
        self.path_to_genome = path_to_genome
        self.path_to_predictions = path_to_predictions
        self.verbose = verbose
        self.cwd = cwd
        self.species = species
        self.model_dir = model_dir or os.environ.get("MGCOD_MODEL_DIR") or os.path.join(cwd, "dependencies")
        self.device = device
        self.records = None
        self.result = None          # engine.Result of the last run_mgm
        self.engine = None

    # -- bin/data.py:108-124
    def read_dna_sequence(self):
        logging.info(f"Reading {self.path_to_genome}")
        if self.verbose:
            print(f"Reading {self.path_to_genome}")
        self._fasta = Fasta(self.path_to_genome)          # native reader; the letters stay in its (page-locked) buffer
        recs = self._fasta.records()
        self._file_order = recs
        # the reference sorts a multi-FASTA by record id (bin/data.py:122); the engine keeps file order (gmhmmp2's output order)
        self._sorted_to_file = np.arange(len(recs), dtype=np.int64)
        if len(recs) > 1:
            self._sorted_to_file = np.argsort(self._fasta.ids)            # the reference's very call
            recs = self._fasta.records(self._sorted_to_file)
        self.records = recs
        return recs

    def _gff_path(self, code):
        return f"{self.path_to_predictions}mgm_{code}_{self.species}.gff"

    # -- bin/data.py:31-54
    def run_mgm(self, need_genes=False):
        """Decode and write the four ``mgm_<code>_<species>.gff`` files (those that do not exist yet, bin/data.py:39-51).  The
        files are formatted by the C writer on background threads (no GIL) while the caller goes on with the gene arrays;
        :meth:`finish_writes` waits for them.  ``need_genes``: decode even when all four files are cached (--isoforms works on
        the decoded arrays and on the sequence resident on the device)."""
        logging.info(f"Running MGM models on {self.species}")
        if self.verbose:
            print(f"Running MGM models on {self.species}")
        missing = [c for c in MGCOD_CODES if not os.path.isfile(self._gff_path(c))]
        if not missing and not need_genes:
            return
        if self.records is None:
            self.read_dna_sequence()
        recs = self._file_order                       # gmhmmp2 writes records in file order
        self.result = self._decode()
        self._decoded_order = recs
        if not missing:
            return
        ids = self._fasta.id_array() if hasattr(self, "_fasta") else [r.id for r in recs]
        lens = np.diff(self._fasta.offsets) if hasattr(self, "_fasta") else [len(r) for r in recs]
        # the four files are formatted side by side (the C writer runs without the GIL)
        from concurrent.futures import ThreadPoolExecutor
        self._writer = ThreadPoolExecutor(max_workers=4)
        self._write_jobs = [self._writer.submit(self.result.write, mi, ids, lens, str(self.path_to_genome),
                                                os.path.join(self.model_dir, f"mgm_{code}.mod"), code, self._gff_path(code), "gff")
                            for mi, code in enumerate(MGCOD_CODES) if code in missing]

    def finish_writes(self):
        """waits for the GFF files of :meth:`run_mgm`; raises what a writer raised"""
        jobs, self._write_jobs = getattr(self, "_write_jobs", []), []
        try:
            for j in jobs:
                j.result()
        finally:
            if getattr(self, "_writer", None) is not None:
                self._writer.shutdown(wait=True)
                self._writer = None

    def close(self):
        """releases the letter buffer (back to the reader's page-locked pool) and the decode result; the records' sequences
        are detached first (materialised as str only where somebody already asked for them)"""
        self.finish_writes()
        made = getattr(getattr(self, "_fasta", None), "_records_made", None) or {}
        for r in made.values():
            r.detach()
        self.result = None
        f = getattr(self, "_fasta", None)
        if f is not None:
            f.close()

    def _decode(self):
        """all contigs of the file under the four parameter files on the GPU → engine.Result"""
        import time
        t0 = time.perf_counter()
        self.engine = shared_engine(self.model_dir, self.device)
        self.timings = {"engine_setup": round(time.perf_counter() - t0, 4)}
        # inputs larger than one device batch are streamed in pieces and merged on the host (GFF files only: the window
        # labels of --isoforms and the recoded-stop scan need the whole input resident)
        result = self.engine.decode_large(self._fasta.buffer, self._fasta.offsets)
        self.engine._resident_owner = self if int(self._fasta.offsets[-1]) <= self.engine.MAX_BATCH_BP else None
        return result

    def ensure_resident(self):
        """→ (engine, record ids in batch order) with this genome's genes and packed sequence resident on the device — what
        the window labels of --isoforms and the recoded-stop scan of :class:`mgcod_b200.results.Results` work on.  Decodes
        again when run_mgm was served from the on-disk cache or another genome has been decoded since; an input that does
        not fit one device batch cannot be made resident."""
        if self.records is None:
            self.read_dna_sequence()
        recs = self._file_order
        if self.engine is None:
            self.engine = shared_engine(self.model_dir, self.device)
        if self.result is None or getattr(self.engine, "_resident_owner", None) is not self:
            if int(self._fasta.offsets[-1]) > self.engine.MAX_BATCH_BP:
                raise RuntimeError(f"{self.path_to_genome}: {int(self._fasta.offsets[-1])} letters exceed one device batch "
                                   f"({self.engine.MAX_BATCH_BP}); --isoforms and the recoded-stop scan need the whole input "
                                   f"resident — split the file (bin/multiprocess_mgcod.py usage) or run the default mode")
            self.result = self.engine.decode_buffer(self._fasta.buffer, self._fasta.offsets)
        self.engine._resident_owner = self
        return self.engine, list(record_ids(recs))

    def gene_frames(self):
        """label → DataFrame of that model's genes (the reference's ``label_to_gene_coords``).  From the decoded arrays when
        this run decoded (no parse of the files just written), else parsed from the cached GFF files as the reference does."""
        if self.result is not None:
            ids = self._fasta.ids if hasattr(self, "_fasta") else [r.id for r in self._file_order]
            return GeneFrames([self.result.genes(m) for m in range(len(MGCOD_CODES))], ids)
        frames = self.extract_gene_coords()
        for df in frames:
            df.loc[:, "contig"] = list(df.sequence.values)       # bin/pipeline.py:73-76
        return dict(enumerate(frames))

    # -- bin/data.py:56-106
    def extract_gene_coords(self):
        logging.info(f"Extracting gene coordinates predicted in {self.species}")
        if self.verbose:
            print(f"Extracting gene coordinates predicted in {self.species}")
        self.finish_writes()
        frames = []
        for code in MGCOD_CODES:
            df = pd.read_csv(self._gff_path(code), sep="\t", header=None, comment="#", names=GFF_COLUMNS)
            df["source"] = df["source"] + f"_{code}"
            df["length"] = df["end"] - df["start"] + 1
            df["sequence"] = df["sequence"].astype(str)
            frames.append(df)
        return tuple(frames)
