"""``Results`` — the reference's output class (bin/results.py:19-385) on top of the CUDA engine.

Same constructor arguments, same files: ``<out><species>.gff`` (final annotation with a ``number of recodings``
attribute per gene), ``<out><species>.tab`` (classification summary), ``<out><species>_stats.tab`` (codon indices
of in-frame recoded stops), optional ``proteins_nt_/proteins_aa_<species>.fasta``.

What runs where: the per-gene in-frame stop scan (``find_recoded_stops`` through ``make_seqrecord``,
bin/results.py:253-294) runs on the GPU over the packed sequence left resident by the decode
(``mgc_recoded_stops``); slicing/formatting the text files stays on the host.

The reference attaches recodings to rows by *position*: ``stop_recodings`` is built contig by contig in sorted-id
order (bin/results.py:362-372) but indexed with the row number of ``coords_w_isoforms``, which is in FASTA order
(bin/results.py:58-72, 175); for a multi-FASTA whose records are not sorted the counts land on other genes.  The
mirror reproduces that, so its files are byte-identical to the reference's (tests/golden/results_*).

Runs with a genetic-code switch (``--isoforms``) take the other branches of the same reference methods: parents of
isoform pairs (``logodd == '.'``) carry no recodings, the ``.tab`` file gets one column per segment, and the GFF
summary lines average each isoform pair (bin/results.py:61-75, 101-157, 213-251).  Golden files: tests/golden/iso_*.
"""
from __future__ import annotations

import logging
import os
import warnings

import numpy as np

from .data import contig_rows, frame_columns, record_ids, rows_by_contig
from .engine import GENE_DTYPE, MGCOD_CODES

VERSION_BANNER = "GeneMark.hmm-2 version 1.17"   # line 2 of the binary's usage text (bin/results.py:185)
STOPS = ("TAA", "TAG", "TGA")                      # order of mgc_recoded_stops' counts
_COMP = str.maketrans("ACGTacgtNn", "TGCAtgcaNn")

# standard table (bin/results.py:307-325); per genetic code one stop becomes a sense codon (:328-345)
_BASES = "TCAG"
_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
_TABLE = {a + b + c: _AAS[16 * i + 4 * j + k] for i, a in enumerate(_BASES) for j, b in enumerate(_BASES)
          for k, c in enumerate(_BASES) if _AAS[16 * i + 4 * j + k] != "*"}
_RECODED = {"4": ("TGA", "W"), "15": ("TAG", "Q"), "101": ("TAA", "Y")}


class Results:
    def __init__(self, pipeline, species, path_to_predictions, path_to_genome, path_to_output, date, delete, amino_acids,
                 nucleotides, verbosity, working_dir):
        self.verbose = verbosity
        self.cwd = working_dir
        self.results = pipeline
        self.species = species
        self.path_to_predictions = path_to_predictions
        self.path_to_final_annotations = path_to_output + species + ".gff"
        self.path_to_nt_seqs = "{}proteins_nt_{}.fasta".format(path_to_output, species)
        self.path_to_aa_seqs = "{}proteins_aa_{}.fasta".format(path_to_output, species)
        self.path_to_genome = path_to_genome
        self.path_to_output = path_to_output + species + ".tab"
        self.date = date
        self.delete = delete
        self.amino_acids = amino_acids
        self.nucleotides = nucleotides
        if self.results.coords_w_isoforms is None:
            raise RuntimeError("Results needs a Pipeline that has been run")
        self.stop_recodings = self.extract_sequences()
        logging.info(f"Writing results for {self.species}")
        if self.verbose:
            print(f"Writing results for {self.species}")
        self.save_results()
        # number of recodings into the GFF attributes, row i <- stop_recodings[i] (bin/results.py:57-75)
        coords = self.results.coords_w_isoforms
        attributes = coords.attributes.to_numpy(dtype=object).copy()     # not the arrow-backed array: item assignment on it copies
        coords["gene_id"] = [float(a.split(";")[0].split(" ")[1]) for a in attributes]
        logodd = coords.logodd.values
        for i in range(len(attributes)):
            if not attributes[i].endswith(";"):
                attributes[i] += ";"
            if isinstance(logodd[i], str) and logodd[i] == ".":      # parent of an isoform pair: no sequence of its own
                self.stop_recodings.insert(i, {"TAA": [], "TGA": [], "TAG": []})
                continue
            r = self.stop_recodings[i]
            attributes[i] = (attributes[i] + " number of recodings: "
                             + f"total {len(r['TAA']) + len(r['TGA']) + len(r['TAG'])}, TAA {len(r['TAA'])}, "
                             + f"TGA {len(r['TGA'])}, TAG {len(r['TAG'])};")
        coords["attributes"] = attributes

    def __call__(self):
        self.write_stats()
        self.write_final_results()
        if self.delete:
            for c in MGCOD_CODES:
                os.remove(f"{self.path_to_predictions}mgm_{c}_{self.species}.gff")

    # ---- bin/results.py:358-385 + 269-294 --------------------------------------------------------------------
    def extract_sequences(self):
        P = self.results
        coords = P.coords_w_isoforms
        engine, file_ids = P.data.ensure_resident()
        index_of = {rid: i for i, rid in enumerate(file_ids)}
        # rows in the order the reference visits them: contig by contig in sorted-id order, frame order inside a contig,
        # parents of isoform pairs (logodd '.') left out (bin/results.py:362-372)
        index = rows_by_contig(coords)
        seq_ids = record_ids(P.sequence)
        present = [(n, index[rid]) for n, rid in enumerate(seq_ids) if rid in index] if len(index) else []
        rows = np.concatenate([pos for _, pos in present]) if present else np.zeros(0, dtype=np.int64)
        contig_file = (np.concatenate([np.full(len(pos), index_of[seq_ids[n]], dtype=np.int32) for n, pos in present])
                       if present else np.zeros(0, dtype=np.int32))
        logodd = coords.logodd.values
        if logodd.dtype == object and len(rows):
            keep = np.fromiter((not (isinstance(v, str) and v == ".") for v in logodd[rows]), dtype=bool, count=len(rows))
            rows, contig_file = rows[keep], contig_file[keep]
        genes = np.zeros(len(rows), dtype=GENE_DTYPE)
        genes["contig"] = contig_file
        genes["left"] = coords.start.values[rows].astype(np.int64)
        genes["right"] = coords.end.values[rows].astype(np.int64)
        genes["strand"] = np.where(coords.strand.values[rows] == "-", -1, 1)
        counts, offs, pos = engine.recoded_stops(genes)      # GPU: in-frame stops of every gene
        # dict order of the reference: TAG, TGA, TAA (bin/results.py:285-288); stop order of the engine: TAA, TAG, TGA
        pos_l, offs_l = pos.tolist(), offs.tolist()
        stop_recodings = [{"TAG": pos_l[offs_l[3 * g + 1]:offs_l[3 * g + 2]], "TGA": pos_l[offs_l[3 * g + 2]:offs_l[3 * g + 3]],
                           "TAA": pos_l[offs_l[3 * g]:offs_l[3 * g + 1]]} for g in range(len(genes))]
        if self.nucleotides or self.amino_acids:
            per_contig = []          # (position in P.sequence, rows of that contig without parents)
            for n, pos_n in present:
                cur = coords.iloc[pos_n]
                per_contig.append((n, cur[cur.logodd != "."].reset_index(drop=True)))
            nt = self._nucleotide_records(per_contig, stop_recodings)
            if self.nucleotides:
                _write_fasta(nt, self.path_to_nt_seqs)
            if self.amino_acids:
                # as in the reference, the loop bound is the gene count of the LAST contig (bin/results.py:378)
                last = per_contig[-1][1] if per_contig and per_contig[-1][0] == P.nr_contigs - 1 else []
                aa = [self.translate_seq(last.iloc[i], nt[i], stop_recodings[i]) for i in range(len(last))]
                _write_fasta(aa, self.path_to_aa_seqs)
        return stop_recodings

    def _nucleotide_records(self, per_contig, stop_recodings):
        """[(id, description, sequence)]: lower case, recoded in-frame stop codons upper case (bin/results.py:269-294)"""
        out = []
        g = 0
        for n, cur in per_contig:
            rec = self.results.sequence[n]
            for i in range(len(cur)):
                p = cur.iloc[i]
                s = rec.seq[p.start - 1:p.end]
                if p.strand == "-":
                    s = s[::-1].translate(_COMP)
                s = list(s.lower())
                r = stop_recodings[g]
                for name in STOPS:
                    for k in r[name]:
                        s[3 * k:3 * k + 3] = [c.upper() for c in s[3 * k:3 * k + 3]]
                desc = (f"{p.start}..{p.end},{p.source}, {p.attributes} number recodings: TAA {len(r['TAA'])},"
                        f"TGA {len(r['TGA'])},TAG {len(r['TAG'])}")
                out.append((f"{i + 1} {rec.id}", desc, "".join(s)))
                g += 1
        return out

    @staticmethod
    def find_recoded_stops(seq, stop):
        """Host restatement (bin/results.py:253-267) for callers that have a string, not a resident batch: codon indices
        of in-frame ``stop`` codons except the last codon, and the sequence with those codons upper-cased."""
        seq = str(seq)
        hits = [p for p in range(0, len(seq) - 3, 3) if seq[p:p + 3] == stop]
        for p in hits:
            seq = seq[:p] + seq[p:p + 3].upper() + seq[p + 3:]
        return [p // 3 for p in hits], seq

    @staticmethod
    def translate_seq(prediction, nucleotide_seq, start_recodings):
        """bin/results.py:296-356: translation under the gene's genetic code (stop codons dropped), lower case, amino
        acids of recoded stops upper case."""
        rid, desc, s = nucleotide_seq
        table = dict(_TABLE)
        for key, (codon, aa) in _RECODED.items():     # same substring tests as the reference: '4' in source, ...
            if key in prediction.source:
                table[codon] = aa
                break
        s = s.upper()
        aa = []
        for p in range(0, len(s) - len(s) % 3, 3):
            codon = s[p:p + 3]
            if codon in table:
                aa.append(table[codon])
            elif codon in STOPS:
                continue
            else:
                aa.append("X")
        aa = list("".join(aa).lower())
        for positions in start_recodings.values():
            for k in positions:
                aa[k] = aa[k].upper()
        return rid, desc, "".join(aa)

    # ---- bin/results.py:86-157 --------------------------------------------------------------------------------
    def save_results(self):
        P = self.results
        if isinstance(P.genome_label, (int, np.integer)):
            first = P.sequence[0].id
            with open(self.path_to_output, "w") as fh:
                fh.write("\t\tSequence\n")
                fh.write(f"{first}\tMajor genetic code\t{float(P.genome_label)}\n")
                fh.write(f"{first}\tTotal gene count\t{float(P.coords_w_isoforms.shape[0])}\n")
            return
        # a switch somewhere: one column per segment.  The reference fills a DataFrame cell by cell; _Table keeps what
        # decides the bytes of its to_csv: row and column order of first use, and whether a column started numeric
        # (floats, "22.0") or with text (values kept as they are, "22").
        level_1 = getattr(P, "_ids", None) or list(record_ids(P.sequence))
        coords_all = P.coords_w_isoforms
        index = rows_by_contig(coords_all)
        lmin, lmax = getattr(P, "_label_min", None), getattr(P, "_label_max", None)
        if lmin is not None:
            single = lmin == lmax
            code_of = {int(k): v for k, v in P.label_to_gcode.items()}
            single_code = [code_of.get(x, code_of[0]) for x in lmin.tolist()]
            lens = P._lens.tolist()
        else:       # a Pipeline-like object with plain dicts (tests): the same rules, contig by contig
            single = np.array([len(P.gene_label_to_gcode[c]) == 1 for c in level_1], dtype=bool)
            single_code = [P.gene_label_to_gcode[c][1] for c in level_1]
            lens = None
        first_g = P.gene_label_to_gcode[level_1[0]]
        table = _Table(level_1, ("Coordinates in bp", "Major genetic code", f"Nr genes with {first_g[1]}", "Nr identical genes",
                                 "Nr isoforms", "Total gene count"))
        prev = 0
        for i in np.flatnonzero(~single).tolist() + [len(level_1)]:
            if i > prev:
                # a run of contigs with one genetic code: each sets a text cell first ("Coordinates in bp") and two numbers in
                # the column "Sequence"; for the table's state that is: the column exists and holds text from here on
                table.touch_text("Sequence")
            prev = i + 1
            if i == len(level_1):
                break
            contig = level_1[i]
            g = P.gene_label_to_gcode[contig]
            segments = sorted(P.confident_1[contig] + P.confident_2[contig])
            coords = contig_rows(coords_all, index, contig)
            start = coords.start.values.astype(np.int64)
            is_parent = np.array([isinstance(v, str) and v == "." for v in coords.logodd.values], dtype=bool)
            source = coords.source.values
            both = "GeneMark.hmm2_{}_{}".format(*sorted([g[1], g[2]]))
            keys = (f"Nr genes with {g[1]}", f"Nr genes with {g[2]}", "Nr identical genes", "Nr isoforms", "Total gene count")
            for n, segment in enumerate(segments, start=1):
                col = f"Segment {n}"
                table.set((contig, "Coordinates in bp"), col, str(segment))
                if segment in P.confident_1[contig]:
                    table.set((contig, "Major genetic code"), col, g[1])
                elif segment in P.confident_2[contig]:
                    table.set((contig, "Major genetic code"), col, g[2])
                here = (start >= segment[0]) & (start <= segment[1])
                nr_isoforms = int((here & is_parent).sum())
                nr_1 = int((here & (source == f"GeneMark.hmm2_{g[1]}")).sum()) - nr_isoforms
                nr_2 = int((here & (source == f"GeneMark.hmm2_{g[2]}")).sum()) - nr_isoforms
                nr_identical = int((here & (source == both)).sum()) - nr_isoforms
                for key, v in zip(keys, (nr_1, nr_2, nr_identical, nr_isoforms, nr_identical + nr_1 + nr_2 + nr_isoforms)):
                    table.set((contig, key), col, v)
            for key in keys:
                table.set((contig, key), "Sequence", table.row_sum((contig, key)))
        # the contigs with one code, as the reference fills them (bin/results.py:112-118): coordinates of their one segment,
        # the code, and the number of rows they have in the isoform table (none: only contigs with a switch go there)
        if lens is not None:
            coord_text = [f"(1, {n})" for n in lens]
        else:
            coord_text = [str(sorted(P.confident_1[c] + P.confident_2[c])[0]) if s_ else "" for c, s_ in zip(level_1, single.tolist())]
        counts = [len(index.get(c, ())) for c in level_1] if len(index) else None
        table.write(self.path_to_output, single.tolist(), coord_text, single_code, counts)

    # ---- bin/results.py:159-181 --------------------------------------------------------------------------------
    def write_stats(self):
        coords = self.results.coords_w_isoforms
        attributes = coords["attributes"].tolist()      # row n of the frame (its index is 0..n-1, bin/results.py:175 uses .loc[n])
        if list(coords.index[:3]) != list(range(min(3, len(coords)))):
            attributes = [coords.loc[n, "attributes"] for n in range(len(self.stop_recodings))]
        lines = []
        with open(self.path_to_output.replace(".tab", "_stats.tab"), "w+") as stat:
            for n, recodings in enumerate(self.stop_recodings):
                positions, stops = [], []
                for name, where in recodings.items():
                    positions.extend(where)
                    stops.extend([name] * len(where))
                if len(set(stops)) != 1:
                    order = np.argsort(positions)
                    positions = [positions[i] for i in order]
                    stops = [stops[i] for i in order]
                line = f"{attributes[n].split(';')[0]} recodings {len(positions)} codon_index_values"
                for p, s in zip(positions, stops):
                    line += f" {p} {s}"
                lines.append(line.replace(" ", "\t"))
            stat.write("\n".join(lines) + ("\n" if lines else ""))

    # ---- bin/results.py:181-251 --------------------------------------------------------------------------------
    def write_final_results(self):
        P = self.results
        if isinstance(P.genome_label, (int, np.integer)):
            tables = [str(P.genome_label)]
        else:   # every label a window ever got, not only the two that survive the smoothing
            labels = (P.genome_label.concatenated() if hasattr(P.genome_label, "concatenated")
                      else np.concatenate([np.asarray(y) for y in P.genome_label]))
            tables = sorted({str(P.label_to_gcode[int(x)]) for x in np.unique(labels).tolist()})
        parameters = ", ".join(f"{self.cwd}/dependencies/mgm_{x}.mod" for x in tables)
        coords = P.coords_w_isoforms
        if coords.shape[0] == 0:
            warnings.warn("No genes were predicted, automatically assign genetic code 11")
        columns = [c for c in frame_columns(P.label_to_gene_coords) if c in coords.columns]
        with open(self.path_to_final_annotations, "w+") as out:
            out.write("##gff-version 2\n")
            out.write("#" + VERSION_BANNER + "\n")
            out.write(f"# File with sequence: {self.path_to_genome}\n")
            out.write(f"# File with MetaGeneMark parameters: {parameters}\n")
            out.write(f"# translation table: {', '.join(tables)}\n")
            out.write(f"# output date start: {self.date.strftime('%a %b %d %Y %H:%M:%S')}\n")
            index = rows_by_contig(coords)
            for n, rid in enumerate(record_ids(P.sequence)):
                if rid not in index:          # no genes on this contig: no block in the file
                    continue
                seq = P.sequence[n]
                size = len(seq)
                cur = contig_rows(coords, index, seq.id).copy()
                cur = cur.loc[:, columns]
                cur.reset_index(drop=True, inplace=True)
                accession = cur.iloc[0, 0]
                out.write("\n")
                out.write(f"##sequence-region {accession} 1 {size}\n")
                # an isoform pair counts once, with the mean of its two genes (bin/results.py:221-245)
                logodd = cur.logodd.values
                parents = [i for i in range(len(logodd)) if isinstance(logodd[i], str) and logodd[i] == "."]
                if parents:
                    members = sorted([i + 1 for i in parents] + [i + 2 for i in parents])
                    plain = [i for i in range(len(logodd)) if i not in set(members) and i not in set(parents)]
                    length = cur.length.values
                    mean_length = [(length[members[i]] + length[members[i + 1]]) / 2 for i in range(0, len(members), 2)]
                    mean_logodd = [(logodd[members[i]] + logodd[members[i + 1]]) / 2 for i in range(0, len(members), 2)]
                    lengths = [length[i] for i in plain] + mean_length
                    average_gene_length = int(np.round(sum(lengths[1:], lengths[0]) / len(lengths)))
                    # object column in the reference: added up one by one, left to right
                    total = 0
                    for i in plain:
                        total = total + logodd[i]
                    total_logodd_score = np.round(total + sum(mean_logodd), decimals=1)
                    n_genes = cur.shape[0] - 2 * len(parents)
                else:
                    average_gene_length = int(np.round(cur.length.values.mean()))
                    total_logodd_score = np.round(cur.logodd.sum(), decimals=1)
                    n_genes = cur.shape[0]
                average_gene_density = np.round(n_genes / (size / 1000), decimals=2)
                cur.drop(["length", "contig"], axis=1, inplace=True)
                cur.to_csv(out, header=False, sep="\t", index=False)
                out.write(f"# {accession}\ttotal_logodd\t{total_logodd_score}\taverage_length\t"
                          f"{average_gene_length}\taverage_density\t{average_gene_density}\n")


class _Table:
    """The cells of the reference's classification DataFrame in the order it creates them (bin/results.py:101-157): every contig
    gets the six standard rows first (implicit here), rows with other keys are appended at first use, columns appear at first
    use, and a column is numeric (floats, "22.0") until a text value is stored in it (values kept as they are from then on)."""

    def __init__(self, contigs, std_keys):
        self.contigs, self.std_keys = contigs, tuple(std_keys)
        self._std = set(self.std_keys)
        self.extra, self.cols, self.kind, self.cells = [], [], {}, {}

    def touch_text(self, col):
        if col not in self.kind:
            self.cols.append(col)
        self.kind[col] = "object"

    def set(self, row, col, value):
        if row not in self.cells:
            self.cells[row] = {}
            if row[1] not in self._std:
                self.extra.append(row)
        text = isinstance(value, str)
        if col not in self.kind:
            self.cols.append(col)
            self.kind[col] = "object" if text else "float"
        elif text:
            self.kind[col] = "object"
        self.cells[row][col] = float(value) if self.kind[col] == "float" else value

    def row_sum(self, row):
        vals = self.cells.get(row, {})
        vals = [vals[c] for c in self.cols if c in vals]
        return sum(vals) if vals else 0

    def write(self, path, single, coord_text, single_code, counts):
        cols = self.cols
        dashes = "\t".join(["-"] * len(cols))
        k0, k1, k2, k3, k4, k5 = self.std_keys
        if "Sequence" in cols:
            si = cols.index("Sequence")
            left, right = "-\t" * si, "\t-" * (len(cols) - si - 1)
        out = ["\t\t" + "\t".join(cols)]
        cells = self.cells
        for n, c in enumerate(self.contigs):
            if single[n]:
                cnt = counts[n] if counts is not None else 0
                out.append(f"{c}\t{k0}\t{left}{coord_text[n]}{right}\n{c}\t{k1}\t{left}{single_code[n]}{right}\n"
                           f"{c}\t{k2}\t{dashes}\n{c}\t{k3}\t{dashes}\n{c}\t{k4}\t{dashes}\n{c}\t{k5}\t{left}{cnt}{right}")
                continue
            for key in self.std_keys:
                row = cells.get((c, key))
                if row:
                    out.append(f"{c}\t{key}\t" + "\t".join([str(row[x]) if x in row else "-" for x in cols]))
                else:
                    out.append(f"{c}\t{key}\t{dashes}")
        for row_key in self.extra:
            row = cells[row_key]
            out.append(f"{row_key[0]}\t{row_key[1]}\t" + "\t".join([str(row[x]) if x in row else "-" for x in cols]))
        with open(path, "w") as fh:
            fh.write("\n".join(out) + "\n")


def _write_fasta(records, path):
    """Biopython's FASTA writer: ``>id description`` and 60 letters per line."""
    with open(path, "w") as fh:
        for rid, desc, s in records:
            fh.write(f">{rid} {desc}\n")
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")
