"""Isoform comparison of the two gene sets of a contig with a genetic-code switch.

Mirror of ``Pipeline.compare_sets_of_predicted_genes`` and its helpers (bin/pipeline.py:559-951; ``get_child_1`` /
``format_child_2`` / ``format_parent`` :494-557; ``insert_unseen_genes`` / ``update_gene_ids`` :1170-1238).  Same
results — the two tables ``coords`` and ``coords_w_isoforms`` and ``gene_labels_to_gcode`` — but not the same
machinery: the reference evaluates five boolean masks over the whole second gene set for every gene of the first
(O(genes²) DataFrame work, minutes for a large genome); here each set is indexed once (sorted starts, ends by value)
and every case is a binary search plus a filter over a handful of rows, and rows are plain dicts until the end.

The cases, in the order the reference tests them for a gene (s, e, strand) of set 1 (bin/pipeline.py:614-905):

  A  set 2 has exactly one gene with the same s, e, strand ............ identical, label 3
  B  exactly one with the same e and strand ............................ isoforms (parent = union), label 3
  C  exactly one with the same s and strand, ends within ``tolerance`` .. isoforms, label 3
  D  exactly one set-2 gene strictly over-spanning it .................. the higher score wins (sum over the set-1
                                                                          genes inside), isoforms if both ends are
                                                                          closer than ``tolerance``
  E  set-2 genes strictly inside it .................................... likewise the other way round
  F  otherwise ......................................................... label 1

Quirks of the reference that change its output are kept and marked ``# ref:`` (strand-blind masks on set 1, the
"first row's attributes for all rows" rewrite, ``previous_end`` only updated in some branches, gene ids renumbered by
string replacement).  Pinned against the reference's own outputs: tests/golden/iso_* (make_golden_isoforms.py).
"""
from __future__ import annotations

import logging

import numpy as np
import pandas as pd

from . import labels as L
from .data import GeneFrames, contig_rows, frame_columns, record_ids, rows_by_contig
from .percontig import PerContig


def _rest(attributes: str) -> str:
    return ";".join(attributes.split(";")[1:])


def gc_percent(seq: str) -> float:
    """Biopython's ``SeqUtils.GC``: G, C and S in either case over the length."""
    if not seq:
        return 0.0
    return sum(seq.count(x) for x in "GCgcSs") * 100.0 / len(seq)


class _Idx(list):
    """row indices of an exact-start lookup: a list that answers ``.shape[0]`` like the index arrays of the range lookups"""

    @property
    def shape(self):
        return (len(self),)


class _GeneSet:
    """One model's genes on one contig, in file order, indexed for the six lookups."""

    def __init__(self, frame):
        if isinstance(frame, dict):      # columns as lists / arrays (GeneFrames.columns: no DataFrame on the way)
            keys = list(frame)
            cols = [frame[k].tolist() if isinstance(frame[k], np.ndarray) else list(frame[k]) for k in keys]
            self.rows = [dict(zip(keys, vals)) for vals in zip(*cols)]
            self.start = np.asarray(frame["start"], dtype=np.int64)
            self.end = np.asarray(frame["end"], dtype=np.int64)
            self.strand = np.asarray(frame["strand"], dtype=object)
            self.logodd = np.asarray(frame["logodd"], dtype=np.float64)
        else:
            self.rows = frame.to_dict("records")
            self.start = frame.start.values.astype(np.int64)
            self.end = frame.end.values.astype(np.int64)
            self.strand = frame.strand.values.astype(object)
            self.logodd = frame.logodd.values.astype(np.float64)
        self.order = np.argsort(self.start, kind="stable")
        self.sorted_start = self.start[self.order]
        self.max_len = int((self.end - self.start).max()) if len(self.rows) else 0
        self.by_end: dict[int, list[int]] = {}
        self.end_l = self.end.tolist()
        for i, e in enumerate(self.end_l):
            self.by_end.setdefault(e, []).append(i)
        self.by_start: dict[int, list[int]] = {}      # the three exact-start lookups are dictionary hits, no array call
        for i, b in enumerate(self.start.tolist()):
            self.by_start.setdefault(b, []).append(i)
        self.strand_l = self.strand.tolist()
        self.seen = np.zeros(len(self.rows), dtype=bool)

    def __len__(self):
        return len(self.rows)

    def _window(self, lo: int, hi: int) -> np.ndarray:
        """rows with lo <= start < hi, in file order"""
        a = np.searchsorted(self.sorted_start, lo, "left")
        b = np.searchsorted(self.sorted_start, hi, "left")
        return np.sort(self.order[a:b])

    def identical(self, s, e, st):
        return _Idx(i for i in self.by_start.get(int(s), ()) if self.end_l[i] == e and self.strand_l[i] == st)

    def same_end(self, e, st):
        i = np.asarray(self.by_end.get(int(e), []), dtype=np.int64)
        return i[self.strand[i] == st]

    def same_start(self, s, st):
        return _Idx(i for i in self.by_start.get(int(s), ()) if self.strand_l[i] == st)

    def overspanning(self, s, e, st):
        i = self._window(s - self.max_len - 1, s)          # a gene reaching beyond e >= s starts no earlier than that
        return i[(self.end[i] > e) & (self.strand[i] == st)]

    def inside(self, s, e, st=None):
        """start >= s and end < e (and the strand, if given)"""
        i = self._window(s, e)
        m = self.end[i] < e
        if st is not None:
            m &= self.strand[i] == st
        return i[m]

    def equal(self, s, e):
        return _Idx(i for i in self.by_start.get(int(s), ()) if self.end_l[i] == e)      # ref: no strand test on set 1

    def copies(self, idx) -> list[dict]:
        return [dict(self.rows[int(i)]) for i in idx]


def update_gene_ids(rows: list[dict]) -> list[dict]:
    """bin/pipeline.py:1211-1238: every gene id of ``rows`` (already copies) + 1, children follow their parent."""
    reference = rows[0]["attributes"].split(";")[0].split(" ")[1]
    current = int(reference) + 1

    def bump(r):
        r["attributes"] = r["attributes"].replace(f"gene_id {reference}", f"gene_id {current}").replace(
            f"Parent={reference}", f"Parent={current}")

    bump(rows[0])
    for r in rows[1:]:
        head = r["attributes"].split(";")[0]
        if not head.startswith(f"gene_id {reference}"):     # ref: a prefix test on the text
            current += 1
            reference = head.split(" ")[1]
        bump(r)
    return rows


def _insert_unseen_stepwise(rows: list[dict], unseen: list[dict]) -> list[dict]:
    """bin/pipeline.py:1170-1209 step by step: every insertion renumbers the rows behind it (O(unseen x rows) string work)."""
    for g in unseen:
        gene = dict(g)
        ind = -1
        for k in range(len(rows) - 1, -1, -1):
            if rows[k]["start"] <= gene["start"]:
                ind = k + 1
                break
        if ind < 0:
            gene["attributes"] = "gene_id 0;" + _rest(gene["attributes"])
            rows = [gene] + update_gene_ids([dict(r) for r in rows])
            continue
        while ind <= len(rows) - 1 and "Parent" in rows[ind]["attributes"]:
            ind += 1
        if ind >= len(rows):
            last = int(rows[-1]["attributes"].split(";")[0].split(" ")[1].split(".")[0])
            gene["attributes"] = f"gene_id {last + 1};" + _rest(gene["attributes"])
            rows = rows + [gene]
        else:
            here = int(rows[ind]["attributes"].split(";")[0].split(" ")[1].split(".")[0])
            gene["attributes"] = f"gene_id {here};" + _rest(gene["attributes"])
            rows = rows[:ind] + [gene] + update_gene_ids([dict(r) for r in rows[ind:]])
    return rows


def _insert_unseen_renumbered(rows: list[dict], unseen: list[dict]):
    """The same result in one pass, for a table in the shape the comparison normally produces: every id is a plain
    integer used by one row (or by a parent and its ``Parent=`` children), ids go up by one from row group to row group,
    and no gene lands in front of the first row.  Then "take the id of the displaced row, everything behind moves up by
    one" keeps the numbering consecutive in list order, so it is enough to find the positions (they depend on the starts
    and the Parent marks only) and number the groups once at the end.  Returns None when the table is not of that shape
    (a parent that kept its model's id, several rows sharing an id): the caller then goes step by step."""
    if not rows:
        return None
    token = [r["attributes"].split(";")[0].split(" ")[1] for r in rows]
    is_child = ["Parent" in r["attributes"] for r in rows]
    group = []                       # id of the group each row belongs to
    for k, (t, ch) in enumerate(zip(token, is_child)):
        if ch:
            if k == 0 or not t.startswith(f"{group[-1]}.") or f"Parent={group[-1]};" not in rows[k]["attributes"]:
                return None
            group.append(group[-1])
        else:
            if not t.isdigit() or (k > 0 and int(t) != group[-1] + 1) or rows[k]["attributes"].count(f"gene_id {t}") != 1:
                return None
            group.append(int(t))
    starts = np.array([r["start"] for r in rows], dtype=np.int64)
    child = np.array(is_child, dtype=bool)
    origin = np.arange(len(rows), dtype=np.int64)            # >= 0: original row; -1 - j: unseen gene j
    for j, g in enumerate(unseen):
        at = np.flatnonzero(starts <= g["start"])
        if at.shape[0] == 0:
            return None
        ind = int(at[-1]) + 1
        while ind < starts.shape[0] and child[ind]:
            ind += 1
        starts = np.insert(starts, ind, g["start"])
        child = np.insert(child, ind, False)
        origin = np.insert(origin, ind, -1 - j)
    out = []
    current = group[0] - 1
    for k, o in enumerate(origin.tolist()):
        if o < 0:
            current += 1
            gene = dict(unseen[-1 - o])
            gene["attributes"] = f"gene_id {current};" + _rest(gene["attributes"])
            out.append(gene)
            continue
        if not is_child[o]:
            current += 1
        r = dict(rows[o])
        if current != group[o]:
            r["attributes"] = r["attributes"].replace(f"gene_id {group[o]}", f"gene_id {current}").replace(
                f"Parent={group[o]}", f"Parent={current}")
        out.append(r)
    return out


def insert_unseen_genes(rows: list[dict], unseen: list[dict]) -> list[dict]:
    """bin/pipeline.py:1170-1209: set-2 genes that matched nothing go behind the last row starting at or before them
    (behind that row's children), take the id of the row they displace, and every later id moves up by one."""
    out = _insert_unseen_renumbered(rows, unseen)
    return out if out is not None else _insert_unseen_stepwise(rows, unseen)


def compare_contig(S1: _GeneSet, S2: _GeneSet, code_1, code_2, tolerance: int, seq: str):
    """One contig: returns (current_gene_set rows with gene_label/source set, rows with isoforms)."""
    g3 = f"{code_1}_{code_2}"
    source = {1: f"GeneMark.hmm2_{code_1}", 2: f"GeneMark.hmm2_{code_2}", 3: f"GeneMark.hmm2_{g3}"}
    both = "_".join(sorted(map(str, [code_1, code_2])))
    cur: list[dict] = []
    iso: list[dict] = []
    gene_labels: list[int] = []
    sources: list[str] = []

    def child_1(s, e, gid, attributes):
        rows = S1.copies(S1.equal(s, e))
        rows[0]["attributes"] = f"gene_id {gid}.{code_1}; Parent={gid};" + _rest(attributes)
        rows[0]["gene_label"] = 1
        return rows

    def child_2(idx, gid):
        rows = S2.copies(idx)
        rows[0]["attributes"] = f"gene_id {gid}.{code_2}; Parent={gid};" + _rest(rows[0]["attributes"])
        rows[0]["gene_label"] = 2
        return rows

    def format_parent(rows, gid, attributes, length=False):
        p = rows[0]
        p["source"] = p["source"].split("_")[0] + "_" + both
        if not length:
            p["attributes"] = f"gene_id {gid};" + _rest(attributes)
        else:
            n = p["end"] - p["start"] + 1
            p["length"] = n
            gc = int(round(gc_percent(seq[p["start"] - 1:p["end"]]), 0))
            p["attributes"] = f"gene_id {gid}; gene_type bacteria; gc {gc}; length {n};"
        p["gene_label"] = 3
        return rows

    def retagged(rows, gid, label, all_rows=False):
        """plain (non-isoform) rows for the isoform table: id rewritten, label set"""
        if rows:
            head = f"gene_id {gid};" + _rest(rows[0]["attributes"])
            for r in rows:                                   # ref: every row gets the first row's attributes
                r["attributes"] = head
                r["gene_label"] = label
        return rows

    previous_end = 0
    gene_id = 1
    for i in range(len(S1)):
        s, e, st = int(S1.start[i]), int(S1.end[i]), S1.strand[i]
        logodd, attributes = float(S1.logodd[i]), S1.rows[i]["attributes"]
        if e <= previous_end:
            continue
        own = S1.equal(s, e)
        hit = S2.identical(s, e, st)
        same_e = S2.same_end(e, st) if hit.shape[0] != 1 else hit
        same_s = S2.same_start(s, st)
        if hit.shape[0] == 1:                                                        # ---- A
            gene_labels.append(3)
            sources.append(source[3])
            S2.seen[hit] = True
            logodd_2 = float(S2.logodd[hit[0]])
            cur.extend(S1.copies(own) if logodd >= logodd_2 else S2.copies(hit))
            c1, c2 = child_1(s, e, gene_id, attributes), child_2(hit, gene_id)
            parent = format_parent(S1.copies(own), gene_id, attributes)
            parent[0]["logodd"] = max(c1[0]["logodd"], c2[0]["logodd"])
            iso.extend(parent)
        elif same_e.shape[0] == 1:                                                   # ---- B
            gene_labels.append(3)
            sources.append(source[3])
            S2.seen[same_e] = True
            logodd_2 = float(S2.logodd[same_e[0]])
            cur.extend(S1.copies(own) if logodd >= logodd_2 else S2.copies(same_e))
            c1, c2 = child_1(s, e, gene_id, attributes), child_2(same_e, gene_id)
            parent = S1.copies(own)
            parent[0]["start"] = min(c1[0]["start"], c2[0]["start"])
            parent = format_parent(parent, gene_id, attributes, length=True)
            parent[0]["logodd"] = "."
            iso.extend(parent + c1 + c2)
        elif same_s.shape[0] == 1 and abs(int(S2.end[same_s[0]]) - e) <= tolerance:   # ---- C
            gene_labels.append(3)
            sources.append(source[3])
            S2.seen[same_s] = True
            logodd_2 = float(S2.logodd[same_s[0]])
            if logodd >= logodd_2:
                cur.extend(S1.copies(own))
            else:
                cur.extend(S2.copies(same_s))
                previous_end = int(S2.end[same_s[0]])
            c1, c2 = child_1(s, e, gene_id, attributes), child_2(same_s, gene_id)
            parent = S1.copies(own)
            parent[0]["end"] = max(c1[0]["end"], c2[0]["end"])
            parent = format_parent(parent, gene_id, attributes, length=True)
            parent[0]["logodd"] = "."
            iso.extend(parent + c1 + c2)
        elif (over := S2.overspanning(s, e, st)).shape[0] == 1:                     # ---- D
            o = int(over[0])
            o_start, o_end, o_logodd = int(S2.start[o]), int(S2.end[o]), float(S2.logodd[o])
            S2.seen[over] = True
            within = S1.inside(o_start, o_end)                                       # ref: either strand
            logodd_1 = float(np.sum(S1.logodd[within]))
            isoforms = s - o_start < tolerance and o_end - e < tolerance
            if isoforms:
                parent = S2.copies(over)
                parent = format_parent(parent, gene_id, parent[0]["attributes"])
                parent[0]["logodd"] = "."
                iso.extend(parent + child_1(s, e, gene_id, attributes) + child_2(over, gene_id))
            if o_logodd > logodd_1:
                gene_labels.append(2)
                sources.append(source[2])
                previous_end = o_end
                cur.extend(S2.copies(over))
                if not isoforms:
                    iso.extend(retagged(S2.copies(over), gene_id, 2))
            elif o_logodd < logodd_1:
                gene_labels.extend([1] * within.shape[0])
                sources.extend([source[1]] * within.shape[0])
                previous_end = int(S1.end[within].max()) if within.shape[0] else float("nan")
                cur.extend(S1.copies(within))
                if not isoforms:
                    iso.extend(retagged(S1.copies(within), gene_id, 1))
            else:
                gene_labels.append(3)
                sources.append(source[3])
                cur.extend(S2.copies(over))
                if not isoforms:
                    iso.extend(retagged(S2.copies(over), gene_id, 3))
        elif (inner := S2.inside(s, e, st)).shape[0] >= 1:                           # ---- E
            logodd_2 = float(np.sum(S2.logodd[inner]))
            S2.seen[inner] = True
            isoforms = (inner.shape[0] == 1 and int(S2.start[inner[0]]) - s < tolerance
                        and e - int(S2.end[inner[0]]) < tolerance)
            if isoforms:
                parent = S1.copies(own)
                parent[0]["source"] = parent[0]["source"].split("_")[0] + "_" + both
                parent[0]["logodd"] = "."                    # ref: the parent keeps the id the model gave the gene
                iso.extend(parent + child_1(s, e, gene_id, attributes) + child_2(inner, gene_id))
            if logodd > logodd_2:
                gene_labels.append(1)
                sources.append(source[1])
                cur.extend(S1.copies(own))
                if not isoforms:
                    iso.extend(retagged(S1.copies(own), gene_id, 1))
            elif logodd < logodd_2:
                gene_labels.extend([2] * inner.shape[0])
                sources.extend([source[2]] * inner.shape[0])
                cur.extend(S2.copies(inner))
                if not isoforms:
                    iso.extend(retagged(S2.copies(inner), gene_id, 2))
            else:
                gene_labels.append(3)
                sources.append(source[3])
                cur.extend(S1.copies(own))
                if not isoforms:
                    iso.extend(retagged(S1.copies(own), gene_id, 1))
        else:                                                                        # ---- F
            gene_labels.append(1)
            sources.append(source[1])
            cur.extend(S1.copies(own))
            iso.extend(retagged(S1.copies(own), gene_id, 1))
        gene_id += 1

    if len(gene_labels) != len(cur):
        # the reference assigns the label list to the rows of the contig and fails the same way
        raise ValueError(f"Must have equal len keys and value when setting with an iterable ({len(cur)} rows, "
                         f"{len(gene_labels)} labels)")
    for r, lab, src in zip(cur, gene_labels, sources):
        r["gene_label"] = lab
        r["source"] = src
    unseen = [dict(S2.rows[int(i)], gene_label=2) for i in np.flatnonzero(~S2.seen)]
    if unseen:
        cur.extend(dict(r) for r in unseen)
        iso = insert_unseen_genes(iso, unseen)
    return cur, iso


def compare_sets_of_predicted_genes(P):
    """``P`` is a :class:`mgcod_b200.pipeline.Pipeline` after label smoothing.  Returns what the reference returns:
    (coords, gene_labels_to_gcode, coords_w_isoforms).

    Only contigs whose smoothed labels name two genetic codes are compared; the others contribute nothing to the isoform
    table and a one-entry dict to ``gene_labels_to_gcode``.  ``coords`` (the plain table, which nothing downstream reads)
    is returned as a :class:`LazyCoords` that builds the frame when somebody looks at it."""
    logging.info(f"Comparing sets of predicted genes in {P.species}")
    if P.verbose:
        print(f"Comparing sets of predicted genes in {P.species}")
    frames = P.label_to_gene_coords
    columns = frame_columns(frames)
    lazy = isinstance(frames, GeneFrames)
    index = None if lazy else {lab: rows_by_contig(df) for lab, df in frames.items()}

    def contig_frame(lab, contig_id):
        return frames.contig_frame(lab, contig_id) if lazy else contig_rows(frames[lab], index[lab], contig_id)

    def contig_columns(lab, contig_id):      # the same rows as plain columns when the frames come from arrays
        return frames.contig_columns(lab, contig_id) if lazy else contig_frame(lab, contig_id)

    lmin, lmax = getattr(P, "_label_min", None), getattr(P, "_label_max", None)
    if lmin is None:
        lmin = np.array([int(np.min(u)) if len(u) else -1 for u in P.updated_genome_labels], dtype=np.int64)
        lmax = np.array([int(np.max(u)) if len(u) else -1 for u in P.updated_genome_labels], dtype=np.int64)
    compared: dict[int, pd.DataFrame] = {}
    explicit: dict = {}
    with_isoforms: list[dict] = []
    code_of = {int(k): v for k, v in P.label_to_gcode.items()}
    messages = []
    seq_ids = getattr(P, "_ids", None) or list(record_ids(P.sequence))
    for n in np.flatnonzero(lmin != lmax).tolist():          # contigs whose smoothed labels name two genetic codes
        rec = P.sequence[n]
        labs = np.unique(P.updated_genome_labels[n])
        label_1, label_2 = int(labs[0]), int(labs[1])
        code_1, code_2 = P.label_to_gcode[label_1], P.label_to_gcode[label_2]
        messages.append(f"Contig {rec.id} in {P.species} carries genetic code {code_1} and {code_2}")
        explicit[rec.id] = {1: str(code_1), 2: str(code_2), 3: f"{code_1}_{code_2}"}
        S1 = _GeneSet(contig_columns(label_1, rec.id))
        S2 = _GeneSet(contig_columns(label_2, rec.id))
        cur, iso = compare_contig(S1, S2, code_1, code_2, P.tolerance, rec.seq)
        # ref: DataFrame.sort_values('start') on an object column = numpy's quicksort over Python ints; rows with equal
        # starts (an isoform pair's two genes) come out in that sort's order, and the smoothing below sees them so
        starts = np.empty(len(cur), dtype=object)
        starts[:] = [r["start"] for r in cur]
        cur = [cur[int(k)] for k in np.argsort(starts, kind="quicksort")]
        updated = L.find_best_sequence_of_labels(np.array([r["gene_label"] for r in cur], dtype=np.int64), None,
                                                 P.consecutive_gene_labels, allow_multiple_labels=True)
        frame = pd.DataFrame(cur, columns=columns + ["gene_label"])
        frame["updated_gene_label"] = updated
        compared[n] = frame
        with_isoforms.extend(iso)
        messages.append(f"Identified {sum(1 for r in iso if r['logodd'] == '.')} possible isoforms in {rec.id} in {P.species}")
    # Contigs with one genetic code: nothing goes to the isoform table (the reference's append to it is commented out,
    # bin/pipeline.py:577), so the final GFF of a mixed input lists only the contigs with a switch.  One deliberate difference
    # in the plain table, which nothing downstream reads: the reference appends the model's genes of EVERY contig there
    # (no contig filter, :574-576) — contigs x genes rows for a metagenome; this keeps the contig's own genes.
    lmin_l = lmin.tolist()
    gene_labels_to_gcode = PerContig(seq_ids, explicit, lambda n: {1: code_of.get(lmin_l[n], code_of[0])})
    if messages:
        logging.info("\n".join(messages))
        if P.verbose:
            print("\n".join(messages))

    ids = list(record_ids(P.sequence))      # the closure below must not hold P: P.coords -> closure -> P would be a cycle, and the
    #                                           FASTA buffer and gene tables of a finished run would wait for the cyclic collector

    def build_coords():       # one block per contig, in contig order
        blocks = dict(compared)
        for n in np.flatnonzero((lmin == lmax) & (lmin >= 0)).tolist():
            lab = lmin_l[n]
            f = contig_frame(lab, ids[n])
            if len(f):
                f = f.copy()
                f["updated_gene_label"] = 1
                blocks[n] = f
        if not blocks:
            return pd.DataFrame(columns=columns)
        return pd.concat([blocks[n] for n in sorted(blocks)], sort=True)

    coords_w_isoforms = pd.DataFrame(with_isoforms, columns=columns + ["gene_label"])
    return LazyCoords(build_coords), gene_labels_to_gcode, coords_w_isoforms


class LazyCoords:
    """The plain comparison table (``Pipeline.coords``): every gene of every contig with its label.  Nothing in the
    reference or here reads it after it is made, and for a metagenome it is the whole gene table, so it is only put
    together when somebody touches it (any attribute / item access goes to the frame)."""

    def __init__(self, build):
        self._build, self._frame = build, None

    def frame(self):
        if self._frame is None:
            self._frame = self._build()
        return self._frame

    def __getattr__(self, name):
        return getattr(self.frame(), name)

    def __getitem__(self, key):
        return self.frame()[key]

    def __len__(self):
        return len(self.frame())
