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
from .data import Data, rows_by_contig
from .engine import MGCOD_CODES


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

    def __call__(self):
        data = Data(self.path_to_genome, self.path_to_predictions, self.verbose, self.cwd, self.species,
                    model_dir=self.model_dir, device=self.device)
        self.data = data
        self.sequence = data.read_dna_sequence()
        self.nr_contigs = len(self.sequence)
        data.run_mgm()
        (self.gene_coords_11, self.gene_coords_4, self.gene_coords_15, self.gene_coords_101) = data.extract_gene_coords()
        for df in (self.gene_coords_11, self.gene_coords_4, self.gene_coords_15, self.gene_coords_101):
            df.loc[:, "contig"] = list(df.sequence.values)
        self.label_to_gene_coords = {0: self.gene_coords_11, 1: self.gene_coords_4, 2: self.gene_coords_15,
                                     3: self.gene_coords_101}
        if not self.isoforms:
            self.genome_label = self.classify_complete_sequence()
            self.coords_w_isoforms = self.label_to_gene_coords[self.gcode_to_label[self.genome_label]]
            return
        self.genome_label, self.logodd_scores = self.identify_most_common_codes()
        n = self.consecutive_windows
        padded = self.circular and self.nr_contigs == 1
        if padded:
            # circular single contig: wrap n labels around both ends (pipeline.py:91-97)
            first, last = self.genome_label[0][:n].copy(), self.genome_label[0][-n:].copy()
            if first.shape[0] < n or last.shape[0] < n:
                self.consecutive_windows = n = min(first.shape[0], last.shape[0])
                warnings.warn(f"Setting consecutive windows to {n}")
            self.genome_label = [np.concatenate([last, self.genome_label[0].copy(), first])]
        # smoothing of every contig's labels in one device call (host restatement: labels.find_best_sequence_of_labels)
        updated = self.data.engine.merge_labels(self.genome_label, n)
        if padded:
            updated = [updated[0][n:-n]]
            self.genome_label = [self.genome_label[0][n:-n]]
        self.updated_genome_labels = updated
        if np.unique(np.concatenate(updated)).shape[0] == 1:
            code = self.label_to_gcode[int(updated[0][0])]
            logging.info(f"{self.species} is genetic code {code}")
            if self.verbose:
                print(f"{self.species} is genetic code {code}")
            self.genome_label = code
            self.coords_w_isoforms = self.label_to_gene_coords[self.gcode_to_label[code]]
            self.predicted_switch_regions = {seq.id: None for seq in self.sequence}
            return
        self.coords, self.gene_label_to_gcode, self.coords_w_isoforms = self.compare_sets_of_predicted_genes()
        self.confident_1, self.confident_2, self.segment_labels = self.segment_genome_based_on_window_labels()
        self.predicted_switch_regions = {}
        for k, seq in enumerate(self.sequence):
            segs = sorted(self.confident_1[seq.id] + self.confident_2[seq.id])
            self.predicted_switch_regions[seq.id] = (np.unique(self.updated_genome_labels[k]),
                                                     [(segs[i][1], segs[i + 1][0]) for i in range(len(segs) - 1)])

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
    def identify_most_common_codes(self):
        logging.info(f"Identifying most common genetic codes in {self.species}")
        if self.verbose:
            print(f"Identifying most common genetic codes in {self.species}")
        if self.data.result is None:
            raise RuntimeError("window labels need the genes of this run on the device: remove the cached "
                               "mgm_*.gff files (or run without -p reuse) so that run_mgm decodes the input")
        order = {r.id: i for i, r in enumerate(self.data._decoded_order)}   # engine contig index (file order)
        lens_file = [len(r) for r in self.data._decoded_order]
        stride_file = [L.resolve_stride(self.stride, n) for n in lens_file]
        offs, labs, cents = self.data.engine.window_labels(lens_file, stride_file, self.window_size, self.consecutive_windows,
                                                           self.short_contigs)
        out_l, out_s = [], []
        not_classified, short = [], []
        for seq in self.sequence:                          # reference order: sorted by contig id
            c = order[seq.id]
            a, b = int(offs[c]), int(offs[c + 1])
            out_l.append(labs[a:b].astype(np.int64))
            out_s.append(cents[a:b].astype(np.float64) / 100.0)
            if len(seq) < 5000:
                logging.warning(f"Contig {seq.id} in {self.species} is shorter than 5000bp. Classification results may be "
                                f"not as reliable")
            if len(seq) <= 5000 and not self.short_contigs:
                not_classified.append(seq.id)
            elif len(seq) < 5000:
                short.append(seq.id)
        if short and self.short_contigs:
            warnings.warn(f"Contig {short[-1]} in {self.species} is shorter than 5000bp. Classification results may be not as reliable")
        if not_classified:
            warnings.warn(f"Contig(s) {', '.join(not_classified)} in {self.species} are shorter than 5000bp. Genetic code has been "
                          f"set to 11. Use --short_contigs to predict genetic code of these contigs")
        return out_l, out_s

    # ---- bin/pipeline.py:387-492 -------------------------------------------------------------------------
    def segment_genome_based_on_window_labels(self):
        logging.info(f"Segmenting genome of {self.species}")
        if self.verbose:
            print(f"Segmenting genome of {self.species}")
        conf_1, conf_2, seg_labels = {}, {}, {}
        index = {lab: rows_by_contig(df) for lab, df in self.label_to_gene_coords.items()}
        cols = {lab: (df.start.values, df.end.values, np.where(df.strand.values == "+", 1, -1))
                for lab, df in self.label_to_gene_coords.items()}
        none = np.zeros(0, dtype=np.int64)
        for k, seq in enumerate(self.sequence):
            genes = {}
            if np.unique(self.updated_genome_labels[k]).shape[0] > 1:      # genes matter only where the label changes
                for lab in self.label_to_gene_coords:
                    pos = index[lab].get(seq.id, none)
                    genes[lab] = tuple(c[pos] for c in cols[lab])
            c1, c2, codes = L.segment_contig(self.updated_genome_labels[k], len(seq), L.resolve_stride(self.stride, len(seq)),
                                             self.window_size, genes, self.label_to_gcode)
            conf_1[seq.id], conf_2[seq.id], seg_labels[seq.id] = c1, c2, codes
            for (a, b), code in zip(sorted(c1 + c2), codes):
                logging.info(f"Contig {seq.id} in {self.species} has genetic code {code} from {a} to {b}")
                if self.verbose:
                    print(f"Contig {seq.id} in {self.species} has genetic code {code} from {a} to {b}")
        return conf_1, conf_2, seg_labels

    # ---- bin/pipeline.py:559-951 -------------------------------------------------------------------------
    def compare_sets_of_predicted_genes(self):
        return iso_compare.compare_sets_of_predicted_genes(self)

    find_best_sequence_of_labels = staticmethod(L.find_best_sequence_of_labels)
