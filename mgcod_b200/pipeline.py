"""``Pipeline`` — the reference's analysis class (bin/pipeline.py:16-137) on top of the CUDA engine.

Same constructor arguments and result attributes (``genome_label``, ``logodd_scores``,
``updated_genome_labels``, ``predicted_switch_regions``, ``label_to_gene_coords``, ``confident_1/2``,
``segment_labels``).  What runs where:

* gene finding under the four models ............ GPU (:class:`mgcod_b200.data.Data` → engine)
* per-file sums and argmax (classify_complete_sequence, pipeline.py:139-155) ... GPU (integer 1/100 units)
* per-window sums, spanning-gene fallback, argmax labels (identify_most_common_codes, :157-385) ... GPU
* label smoothing (find_best_sequence_of_labels, :1064-1168) ... GPU (`k_merge_labels`, thread per contig); a host
  restatement lives in :mod:`mgcod_b200.labels`
* switch coordinates (segment_genome_based_on_window_labels, :387-492) ... host, :mod:`mgcod_b200.labels`

* isoform comparison of the two gene sets of a contig with a switch (compare_sets_of_predicted_genes, :559-951)
  ... host, :mod:`mgcod_b200.isoforms` (indexed lookups instead of the reference's per-gene DataFrame masks)
"""
from __future__ import annotations

import logging
import warnings

import numpy as np

from . import isoforms as iso_compare
from . import labels as L
from .data import Data, GeneFrames, record_ids, record_lengths, rows_by_contig
from .engine import MGCOD_CODES
from .percontig import PerContig, Ragged


class Pipeline:
    def __init__(self, path_to_genome, path_to_predictions, species, isoforms, circular, consecutive_windows,
                 consecutive_gene_labels, window_size, stride, tolerance, short_contigs, verbosity, working_dir,
                 model_dir=None, device=0):
        self.verbose = verbosity
        self.cwd = working_dir
        self.path_to_genome = path_to_genome
        self.path_to_predictions = path_to_predictions
        self.species = species
        self.circular = circular
        self.tolerance = tolerance
        self.predicted_switch_regions = None
        self.window_size = window_size
        self.stride = stride
        self.isoforms = isoforms
        self.short_contigs = short_contigs
        self.consecutive_windows = consecutive_windows
        self.consecutive_gene_labels = consecutive_gene_labels
        self.genome_label = 11
        self.label_to_gcode = {i: c for i, c in enumerate(MGCOD_CODES)}
        self.gcode_to_label = {c: i for i, c in enumerate(MGCOD_CODES)}
        self.model_dir = model_dir
        self.device = device
        self.coords_w_isoforms = None

    # the reference's four frame attributes, built on first use (GeneFrames)
    gene_coords_11 = property(lambda self: self.label_to_gene_coords[0])
    gene_coords_4 = property(lambda self: self.label_to_gene_coords[1])
    gene_coords_15 = property(lambda self: self.label_to_gene_coords[2])
    gene_coords_101 = property(lambda self: self.label_to_gene_coords[3])

    def __call__(self):
        import time
        t = [time.perf_counter()]
        data = Data(self.path_to_genome, self.path_to_predictions, self.verbose, self.cwd, self.species,
                    model_dir=self.model_dir, device=self.device)
        self.data = data
        self.sequence = data.read_dna_sequence()
        self.nr_contigs = len(self.sequence)
        t.append(time.perf_counter())
        data.run_mgm(need_genes=self.isoforms)      # the four GFF files are written in the background meanwhile
        t.append(time.perf_counter())
        try:
            self._analyse()
            t.append(time.perf_counter())
        finally:
            data.finish_writes()
        t.append(time.perf_counter())
        # where the wall time of this run went (seconds): reading the FASTA, decode (incl. engine set-up on first use),
        # labelling + comparison + segmentation, waiting for the background GFF writers
        self.timings = dict(zip(("read", "decode", "analyse", "wait_for_gff_writers"), np.diff(t).round(4).tolist()))
        self.timings.update(getattr(data, "timings", {}))

    def close(self):
        """Hands the run's large buffers back (the page-locked letter buffer of the FASTA reader, the gene tables): called by
        the command line when the result files are written; a later genome of the same process reuses them."""
        data = getattr(self, "data", None)
        if data is not None:
            data.close()

    def _analyse(self):
        data = self.data
        # label -> frame of that model's genes (reference: four DataFrames parsed back from the GFF files, plus the `contig`
        # column of bin/pipeline.py:68-76); here built from the decoded arrays, lazily
        self.label_to_gene_coords = data.gene_frames()
        if not self.isoforms:
            self.genome_label = self.classify_complete_sequence()
            self.coords_w_isoforms = self.label_to_gene_coords[self.gcode_to_label[self.genome_label]]
            return
        n = self.consecutive_windows
        if self.circular and self.nr_contigs == 1:
            self.genome_label, self.logodd_scores = self.identify_most_common_codes()
            # circular single contig: wrap n labels around both ends (pipeline.py:91-97), smooth, drop the padding
            first, last = self.genome_label[0][:n].copy(), self.genome_label[0][-n:].copy()
            if first.shape[0] < n or last.shape[0] < n:
                self.consecutive_windows = n = min(first.shape[0], last.shape[0])
                warnings.warn(f"Setting consecutive windows to {n}")
            padded = [np.concatenate([last, self.genome_label[0].copy(), first])]
            updated = [self.data.engine.merge_labels(padded, n)[0][n:-n]]
        else:
            # window sums, labels and their smoothing in one device call (host restatement: labels.find_best_sequence_of_labels)
            self.genome_label, self.logodd_scores, updated = self.identify_most_common_codes(smooth=True)
        self.updated_genome_labels = updated
        flat = updated.concatenated() if isinstance(updated, Ragged) else (np.concatenate(updated) if updated else np.zeros(0, dtype=np.int64))
        # per contig: smallest / largest smoothed label (equal = one genetic code on that contig)
        if isinstance(updated, Ragged):
            lo, sizes = updated.lo, updated.hi - updated.lo
            src = updated.flat
        else:
            sizes = np.array([u.shape[0] for u in updated], dtype=np.int64)
            lo = np.concatenate([[0], np.cumsum(sizes)[:-1]]) if len(sizes) else np.zeros(0, dtype=np.int64)
            src = flat
        nonempty = sizes > 0
        self._label_min = np.full(len(sizes), -1, dtype=np.int64)
        self._label_max = np.full(len(sizes), -1, dtype=np.int64)
        if nonempty.any():
            # reduceat wants ascending starts: the contigs in the order of their slices, results scattered back
            order = np.argsort(lo[nonempty], kind="stable")
            idx = np.flatnonzero(nonempty)[order]
            starts = lo[idx]
            ends = starts + sizes[idx]
            if bool((ends[:-1] == starts[1:]).all()) and ends[-1] == src.shape[0] and starts[0] == 0:
                self._label_min[idx] = np.minimum.reduceat(src, starts)
                self._label_max[idx] = np.maximum.reduceat(src, starts)
            else:       # slices with holes between them (does not happen for the engine's layouts)
                self._label_min[idx] = [int(src[a:b].min()) for a, b in zip(starts, ends)]
                self._label_max[idx] = [int(src[a:b].max()) for a, b in zip(starts, ends)]
        present = self._label_min[nonempty]
        if present.size and int(present.min()) == int(self._label_max[nonempty].max()):
            code = self.label_to_gcode[int(present[0])]
            logging.info(f"{self.species} is genetic code {code}")
            if self.verbose:
                print(f"{self.species} is genetic code {code}")
            self.genome_label = code
            self.coords_w_isoforms = self.label_to_gene_coords[self.gcode_to_label[code]]
            self.predicted_switch_regions = {rid: None for rid in record_ids(self.sequence)}
            return
        self._ids, self._lens = list(record_ids(self.sequence)), record_lengths(self.sequence)   # no 20 000 record objects
        self.coords, self.gene_label_to_gcode, self.coords_w_isoforms = self.compare_sets_of_predicted_genes()
        self.confident_1, self.confident_2, self.segment_labels = self.segment_genome_based_on_window_labels()
        explicit = {}
        for k in np.flatnonzero(self._label_min != self._label_max).tolist():
            cid = self._ids[k]
            segs = sorted(self.confident_1[cid] + self.confident_2[cid])
            explicit[cid] = (np.unique(self.updated_genome_labels[k]), [(segs[i][1], segs[i + 1][0]) for i in range(len(segs) - 1)])
        upd = self.updated_genome_labels
        self.predicted_switch_regions = PerContig(self._ids, explicit, lambda k: (upd[k][:1].copy(), []))

    # ---- bin/pipeline.py:139-155 -------------------------------------------------------------------------
    def classify_complete_sequence(self):
        logging.info(f"Classifying {self.species}")
        if self.verbose:
            print(f"Classifying {self.species}")
        if self.data.result is not None:
            sums = self.data.result.model_cents()          # device-side integer sums of the %.2f values
        else:                                              # GFFs came from the on-disk cache
            sums = [int(round(df["logodd"].sum() * 100)) for df in self.label_to_gene_coords.values()]
        best = int(np.argmax(sums))
        logging.info(f"{self.species} is genetic code {self.label_to_gcode[best]}")
        if self.verbose:
            print(f"{self.species} is genetic code {self.label_to_gcode[best]}")
        return self.label_to_gcode[best]

    # ---- bin/pipeline.py:157-385 -------------------------------------------------------------------------
    def identify_most_common_codes(self, smooth=False):
        """→ (labels per contig, window scores per contig), contigs in the reference's order (sorted by id).  ``smooth=True``
        also returns the smoothed labels (the reference's next step, find_best_sequence_of_labels per contig with the noise
        labels of the whole input, bin/pipeline.py:98-110), computed in the same device call."""
        logging.info(f"Identifying most common genetic codes in {self.species}")
        if self.verbose:
            print(f"Identifying most common genetic codes in {self.species}")
        engine, _ = self.data.ensure_resident()            # decodes (again) when run_mgm was served from cached GFF files
        recs = self.data._file_order
        lens_file = record_lengths(recs)
        if self.stride is not None:
            stride_file = np.full(len(recs), int(self.stride), dtype=np.int32)
        else:
            stride_file = np.where(lens_file <= 100000, 2500, 5000).astype(np.int32)      # bin/pipeline.py:172-183
        if smooth:
            offs, labs, cents, merged, status = engine.isoform_labels(lens_file, stride_file, self.window_size,
                                                                      self.consecutive_windows, self.short_contigs)
            if (status != 0).any():
                raise IndexError(f"label smoothing walks off the array for contig {recs[int(np.flatnonzero(status)[0])].id} "
                                 f"(as the reference does)")
        else:
            offs, labs, cents = engine.window_labels(lens_file, stride_file, self.window_size, self.consecutive_windows,
                                                     self.short_contigs)
            merged = None
        labs = labs.astype(np.int64)
        scores = cents.astype(np.float64) / 100.0
        order = self.data._sorted_to_file                    # reference order: sorted by contig id
        lo, hi = offs[:-1][order], offs[1:][order]
        out_l = Ragged(labs, lo, hi)                         # list-like: one array per contig, made when asked for
        out_s = Ragged(scores, lo, hi)
        lens_sorted = lens_file[order]
        short = np.flatnonzero(lens_sorted <= 5000)
        not_classified, below = [], []
        for k in short.tolist():
            seq = self.sequence[k]
            if len(seq) < 5000:
                logging.warning(f"Contig {seq.id} in {self.species} is shorter than 5000bp. Classification results may be "
                                f"not as reliable")
            if not self.short_contigs:
                not_classified.append(seq.id)
            elif len(seq) < 5000:
                below.append(seq.id)
        if below and self.short_contigs:
            warnings.warn(f"Contig {below[-1]} in {self.species} is shorter than 5000bp. Classification results may be not as reliable")
        if not_classified:
            warnings.warn(f"Contig(s) {', '.join(not_classified)} in {self.species} are shorter than 5000bp. Genetic code has been "
                          f"set to 11. Use --short_contigs to predict genetic code of these contigs")
        if smooth:
            return out_l, out_s, Ragged(merged.astype(np.int64), lo, hi)
        return out_l, out_s

    # ---- bin/pipeline.py:387-492 -------------------------------------------------------------------------
    def segment_genome_based_on_window_labels(self):
        logging.info(f"Segmenting genome of {self.species}")
        if self.verbose:
            print(f"Segmenting genome of {self.species}")
        frames = self.label_to_gene_coords
        ids = getattr(self, "_ids", None) or list(record_ids(self.sequence))
        lens = getattr(self, "_lens", None)
        if lens is None:
            lens = record_lengths(self.sequence)
        lmin, lmax = getattr(self, "_label_min", None), getattr(self, "_label_max", None)
        if lmin is None:
            lmin = np.array([int(np.min(u)) if len(u) else -1 for u in self.updated_genome_labels], dtype=np.int64)
            lmax = np.array([int(np.max(u)) if len(u) else -1 for u in self.updated_genome_labels], dtype=np.int64)
        code_of = {int(k): v for k, v in self.label_to_gcode.items()}
        c1, c2, sl = {}, {}, {}
        switch_msgs = {}
        for k in np.flatnonzero(lmin != lmax).tolist():        # genes matter only where the label changes
            labels_k = self.updated_genome_labels[k]
            genes = {int(lab): _gene_columns(frames, int(lab), ids[k]) for lab in np.unique(labels_k).tolist()}
            a, b, codes = L.segment_contig(labels_k, int(lens[k]), L.resolve_stride(self.stride, int(lens[k])), self.window_size, genes,
                                           self.label_to_gcode)
            c1[ids[k]], c2[ids[k]], sl[ids[k]] = a, b, codes
            switch_msgs[k] = [f"Contig {ids[k]} in {self.species} has genetic code {code} from {x} to {y}"
                              for (x, y), code in zip(sorted(a + b), codes)]
        # one genetic code on the whole contig: one block, no gene coordinates needed — made when somebody asks
        lens_l, lmin_l = lens.tolist(), lmin.tolist()
        conf_1 = PerContig(ids, c1, lambda k: [(1, lens_l[k])])
        conf_2 = PerContig(ids, c2, lambda k: [])
        seg_labels = PerContig(ids, sl, lambda k: [code_of.get(lmin_l[k], code_of[0])])
        if logging.getLogger().isEnabledFor(logging.INFO) or self.verbose:
            sp = self.species
            messages = []
            for k, cid in enumerate(ids):
                if k in switch_msgs:
                    messages.extend(switch_msgs[k])
                else:
                    messages.append(f"Contig {cid} in {sp} has genetic code {code_of.get(lmin_l[k], code_of[0])} from 1 to {lens_l[k]}")
            logging.info("\n".join(messages))        # one record: a metagenome has tens of thousands of these lines
            if self.verbose:
                print("\n".join(messages))
        return conf_1, conf_2, seg_labels

    # ---- bin/pipeline.py:559-951 -------------------------------------------------------------------------
    def compare_sets_of_predicted_genes(self):
        return iso_compare.compare_sets_of_predicted_genes(self)

    find_best_sequence_of_labels = staticmethod(L.find_best_sequence_of_labels)


def _gene_columns(frames, lab, contig_id):
    """(left, right, strand as +1/-1) of one model's genes on one contig, in file order"""
    if isinstance(frames, GeneFrames):
        a, b = frames.contig_range(lab, contig_id)
        g = frames.genes[lab][a:b]
        return g["left"].astype(np.int64), g["right"].astype(np.int64), np.where(g["strand"] > 0, 1, -1)
    df = frames[lab]
    df = df[df.contig == contig_id]
    return df.start.values, df.end.values, np.where(df.strand.values == "+", 1, -1)
