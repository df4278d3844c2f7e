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


class Record:
    """The slice of Biopython's SeqRecord that Mgcod touches: ``id``, ``description``, ``seq``, ``len()``.
    ``seq`` may be given as a uint8 view into a :class:`Fasta` buffer; it is decoded to ``str`` on first use."""

    def __init__(self, rid: str, description: str, seq):
        self.id, self.description, self._seq = rid, description, seq

    @property
    def seq(self) -> str:
        if not isinstance(self._seq, str):
            self._seq = self._seq.tobytes().decode("ascii", errors="replace")
        return self._seq

    def __len__(self):
        return len(self._seq)


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
        self.ids = [L.mgc_fasta_id(self._h, i).decode() for i in range(n)]
        self.deflines = [L.mgc_fasta_defline(self._h, i).decode() for i in range(n)]
        self.offsets = np.ctypeslib.as_array(L.mgc_fasta_offsets(self._h), shape=(n + 1,)).copy()
        total = int(self.offsets[-1])
        ptr = L.mgc_fasta_sequence(self._h)
        self.buffer = (np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(total, 1),))[:total]
                       if ptr else np.zeros(0, np.uint8))

    def records(self) -> list:
        return [Record(i, d, self.buffer[self.offsets[k]:self.offsets[k + 1]]) for k, (i, d) in enumerate(zip(self.ids, self.deflines))]

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
        _ENGINES[key] = Engine([ModelFile(os.path.join(model_dir, f"mgm_{c}.mod")) for c in MGCOD_CODES], device=device)
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
        if len(recs) > 1:
            recs = [recs[x] for x in np.argsort([r.id for r in recs])]
        self.records = recs
        return recs

    def _gff_path(self, code):
        return f"{self.path_to_predictions}mgm_{code}_{self.species}.gff"

    # -- bin/data.py:31-54
    def run_mgm(self):
        logging.info(f"Running MGM models on {self.species}")
        if self.verbose:
            print(f"Running MGM models on {self.species}")
        missing = [c for c in MGCOD_CODES if not os.path.isfile(self._gff_path(c))]
        if not missing:
            return
        if self.records is None:
            self.read_dna_sequence()
        recs = self._file_order                       # gmhmmp2 writes records in file order
        self.engine = shared_engine(self.model_dir, self.device)
        # inputs larger than one device batch are streamed in pieces and merged on the host (GFF files only: the window
        # labels of --isoforms and the recoded-stop scan need the whole input resident)
        self.result = self.engine.decode_large(self._fasta.buffer, self._fasta.offsets)
        self.engine._resident_owner = self if int(self._fasta.offsets[-1]) <= self.engine.MAX_BATCH_BP else None
        self._decoded_order = recs
        ids = [r.id for r in recs]
        lens = [len(r) for r in recs]
        # the four files are formatted side by side (the C writer runs without the GIL)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=4) as pool:
            jobs = [pool.submit(self.result.write, mi, ids, lens, str(self.path_to_genome),
                                os.path.join(self.model_dir, f"mgm_{code}.mod"), code, self._gff_path(code), "gff")
                    for mi, code in enumerate(MGCOD_CODES) if code in missing]
            for j in jobs:
                j.result()

    def ensure_resident(self):
        """→ (engine, record ids in batch order) with this genome's packed sequence resident on the device — what
        :class:`mgcod_b200.results.Results` scans for recoded stops.  Decodes again when run_mgm was served from the
        on-disk cache (or another genome has been decoded since)."""
        if self.records is None:
            self.read_dna_sequence()
        recs = self._file_order
        if self.engine is None:
            self.engine = shared_engine(self.model_dir, self.device)
        if self.result is None or getattr(self.engine, "_resident_owner", None) is not self:
            self.result = self.engine.decode_buffer(self._fasta.buffer, self._fasta.offsets)
        self.engine._resident_owner = self
        return self.engine, [r.id for r in recs]

    # -- bin/data.py:56-106
    def extract_gene_coords(self):
        logging.info(f"Extracting gene coordinates predicted in {self.species}")
        if self.verbose:
            print(f"Extracting gene coordinates predicted in {self.species}")
        frames = []
        for code in MGCOD_CODES:
            df = pd.read_csv(self._gff_path(code), sep="\t", header=None, comment="#", names=GFF_COLUMNS)
            df["source"] = df["source"] + f"_{code}"
            df["length"] = df["end"] - df["start"] + 1
            df["sequence"] = df["sequence"].astype(str)
            frames.append(df)
        return tuple(frames)
