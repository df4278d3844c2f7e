/*
 * mgcod_b200.h — C-ABI of the B200 engine that replaces Mgcod's gene-finder boundary.
 *
 * What it replaces in the reference (gatech-genemark/Mgcod):
 *   - bin/data.py:39-54   Data.run_mgm(): four fork/exec of dependencies/gmhmmp2
 *                         (-M mgm_<code>.mod -s <fasta> -f gff -o <pred>/mgm_<code>_<species>.gff)
 *   - bin/data.py:56-106  Data.extract_gene_coords(): re-reading the four GFF files
 *   - bin/pipeline.py:139-155, 157-385   per-file / per-window log-odds sums and argmax labels
 *   - bin/multiprocess_mgcod.py:140-145  process-level fan-out over genomes (here: one batch of contigs)
 *
 * Conventions: plain pointers and sizes; every function returns 0 on success and a negative
 * MGC_E_* code on failure, with a message available from mgc_last_error() (thread-local).
 * No function falls back to a CPU path: engine functions fail with MGC_E_CUDA when no
 * sm_100 device is usable.
 */
#ifndef MGCOD_B200_H
#define MGCOD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGC_OK 0
#define MGC_E_ARG (-1)     /* bad argument */
#define MGC_E_IO (-2)      /* file cannot be read / written */
#define MGC_E_PARSE (-3)   /* malformed .mod or FASTA (gmhmmp2 prints "error, ..." and exits non-zero) */
#define MGC_E_CUDA (-4)    /* no usable device / CUDA failure */
#define MGC_E_ALPHABET (-5)/* a letter outside the accepted alphabet (ACGTU + IUPAC RYSWKMBDHVN, either case); gmhmmp2:
                                "error, unexpected letter found in FASTA sequence" */
#define MGC_E_NOMEM (-6)

#define MGC_MAX_MODELS 8

typedef struct mgc_model mgc_model;     /* parsed parameter file (host) */
typedef struct mgc_engine mgc_engine;   /* one per process per GPU: device tables + scratch + stream */
typedef struct mgc_result mgc_result;   /* genes of one decoded batch, host side */

/* One predicted gene == one CDS line of gmhmmp2's GFF (format string in SURVEY.md §8b). */
typedef struct {
    int32_t contig;        /* index into the batch */
    int32_t left, right;   /* 1-based inclusive, as GFF columns 4,5 */
    int8_t strand;         /* +1 / -1 */
    int8_t domain;         /* 0 bacteria, 1 archaea  (attribute gene_type) */
    int8_t partial_left;   /* attribute "partial XY": X */
    int8_t partial_right;  /*                         Y */
    int32_t gc;            /* attribute gc */
    double logodd;         /* column 6 before %.2f formatting */
} mgc_gene;

/* Device-side timings of the last decode, milliseconds (CUDA events on the engine's stream). */
typedef struct {
    float h2d, pack, orf_find, orf_prep, orf_score /* walk */, orf_final, order, chain, aggregate, d2h, total;
    int64_t n_orfs, n_starts, n_walks, n_candidates, n_genes; /* summed over models */
    int32_t kernel_launches;
} mgc_stats;

const char *mgc_version(void);
const char *mgc_last_error(void);
int mgc_device_count(void);

/* ---- parameter files (replaces gmhmmp2's ParameterParser; host only, no GPU needed) ---- */
int mgc_model_load(const char *mod_path, mgc_model **out);
void mgc_model_free(mgc_model *m);
/* info[0]=gcode [1]=cod_order [2]=non_order [3]=gene_min_length [4]=#bacterial sections [5]=#archaeal */
int mgc_model_info(const mgc_model *m, int32_t info[6]);

/* ---- engine ---- */
/* models[i] is decoded as "label i" (Mgcod order: 11, 4, 15, 101 — bin/pipeline.py:53-54). */
int mgc_engine_create(int device, const mgc_model *const *models, int n_models, mgc_engine **out);
void mgc_engine_destroy(mgc_engine *e);

/*
 * Decode a batch of contigs under every model of the engine.
 *   seq            concatenated ASCII sequences (host memory, pinned or pageable)
 *   offsets[n+1]   contig i occupies seq[offsets[i] .. offsets[i+1])
 * Equivalent to running gmhmmp2 once per model on a multi-FASTA of these contigs.
 */
int mgc_decode_batch(mgc_engine *e, const char *seq, const int64_t *offsets, int32_t n_contigs,
                     mgc_result **out);

/* Two-step form used by the benchmark: copy + pack once, decode many times with inputs resident in HBM. */
int mgc_upload_batch(mgc_engine *e, const char *seq, const int64_t *offsets, int32_t n_contigs);
int mgc_decode_resident(mgc_engine *e, mgc_result **out);
/* Streams of batches: starts the host->device copy of the NEXT batch on a second stream and returns at once; the copy
 * runs while mgc_decode_resident works on the current batch.  The next mgc_upload_batch / mgc_decode_batch with the same
 * seq pointer and offsets adopts the copied letters instead of copying again.  seq must stay valid and unchanged until
 * then, and must be page-locked for the copy to overlap (pageable memory works but is staged synchronously).
 * Replaces the process-level overlap of bin/multiprocess_mgcod.py:140-145 (one genome loading while another decodes). */
int mgc_prefetch_batch(mgc_engine *e, const char *seq, const int64_t *offsets, int32_t n_contigs);

int mgc_engine_stats(const mgc_engine *e, mgc_stats *out);
/* The gene table of the last decode as it sits in DEVICE memory (mgc_gene rows ordered by (contig, left); contig index local
 * to the batch): for a multi-GPU gather over NCCL straight from the engine's memory — the final gather of
 * bin/multiprocess_mgcod.py:140-145's fan-out.  Valid until the next decode / upload on this engine. */
int mgc_engine_device_genes(mgc_engine *e, int model, const void **dev_genes, int64_t *count);

/* ---- results ----
 * A result is returned as soon as the decode has finished on the device; its gene tables and totals travel to page-locked host
 * memory on a copy-out stream while the caller may already start the next decode.  mgc_result_genes / _contig_totals /
 * mgc_write_gff wait for that copy; gene counts and mgc_result_model_cents are available at once. */
int64_t mgc_result_num_genes(const mgc_result *r, int model);
/* genes of one model ordered by (contig, left); pointer valid until mgc_result_free */
const mgc_gene *mgc_result_genes(const mgc_result *r, int model);
/* trailer value "total_logodd" per contig (chain score incl. overlap penalties) */
const double *mgc_result_contig_totals(const mgc_result *r, int model);
/* sum over all genes of round(logodd,2) in 1/100 units — Pipeline.classify_complete_sequence */
int mgc_result_model_cents(const mgc_result *r, int64_t *cents /*[n_models]*/);
/* $GENE_MIN_LENGTH of the model the genes came from (90 for results made by mgc_result_create until set): gmhmmp2
 * skips records shorter than it, and mgc_write_gff then omits their summary line */
int32_t mgc_result_min_gene_length(const mgc_result *r, int model);
int mgc_result_set_min_gene_length(mgc_result *r, int model, int32_t min_len);
void mgc_result_free(mgc_result *r);
/* Host-side constructor (no GPU): wraps caller-provided genes, e.g. parsed back from cached GFF files
 * (bin/data.py:39 skips gmhmmp2 when mgm_<code>_<species>.gff exists).  genes[m] holds counts[m] genes ordered
 * by (contig, left); totals[m] holds n_contigs chain totals. */
int mgc_result_create(int n_models, int32_t n_contigs, const mgc_gene *const *genes, const int64_t *counts,
                      const double *const *totals, mgc_result **out);

/*
 * In-frame stop codons inside genes (bin/results.py:253-294, Results.find_recoded_stops via make_seqrecord): for each
 * gene (contig, left, right, strand as in mgc_gene, coordinates of the last uploaded batch) the codon indices
 * k = 0 .. ncodons-2 whose codon, read on the gene's strand, is TAA / TAG / TGA.  Runs on the packed sequence resident
 * on the device after a decode.
 *   counts[n_genes*3]     per gene: number of TAA, TAG, TGA codons found (in that order)
 *   pos_off[n_genes*3+1]  exclusive prefix of counts
 *   positions[pos_cap]    codon indices, ascending per (gene, stop); filled only if pos_cap >= the total
 * Returns the total number of positions (call once with positions = NULL to size the array), negative on error.
 */
int64_t mgc_recoded_stops(mgc_engine *e, const mgc_gene *genes, int64_t n_genes, int32_t *counts, int64_t *pos_off,
                          int32_t *positions, int64_t pos_cap);

/*
 * Sliding-window labels (bin/pipeline.py:157-385) computed on the device from the genes of the
 * last decode.  For contig c: windows start at 0, stride, ... < len; window [start, start+w).
 *   stride[c]      per-contig stride (the host resolves the 2500/5000 default, pipeline.py:172-183)
 *   win_offsets    [n_contigs+1] prefix of window counts, filled by the caller via mgc_window_counts
 *   labels         [total windows] argmax label (0..n_models-1)
 *   cents          [total windows * n_models] window sums in 1/100 units
 */
int mgc_window_counts(const int64_t *contig_len, const int32_t *stride, int32_t n_contigs, int32_t window,
                      int32_t consecutive, int short_contigs, int64_t *win_offsets);
int mgc_window_labels(mgc_engine *e, const int32_t *stride, int32_t n_contigs, int32_t window, int32_t consecutive,
                      int short_contigs, const int64_t *win_offsets, int32_t *labels, int64_t *cents);
/* n_contigs = entries of stride[] (win_offsets has one more); MGC_E_ARG when it differs from the batch resident on the device
 * (e.g. an input that was streamed in several batches: only the last one is resident). */

/*
 * --isoforms labelling in ONE call on the device (bin/pipeline.py:157-385 followed by :98-110 and :1064-1168): window sums and
 * argmax labels, the label histogram of the whole input (its two most frequent labels stay, the others are noise), and the
 * smoothing of every contig's labels — the labels do not visit the host in between.
 *   labels, merged   [total windows]      raw argmax labels / smoothed labels
 *   cents            [total windows * n_models] window sums in 1/100 units; may be NULL (only plots read them)
 *   status           [n_contigs]          as mgc_merge_labels (1 = the reference raises IndexError on that contig)
 * Not for the circular single-contig case (-r), whose labels are padded on the host between the two steps.
 */
int mgc_isoform_labels(mgc_engine *e, const int32_t *stride, int32_t n_contigs, int32_t window, int32_t consecutive,
                       int short_contigs, const int64_t *win_offsets, int32_t *labels, int64_t *cents, int32_t *merged,
                       int32_t *status);

/*
 * Label smoothing on the device (Pipeline.find_best_sequence_of_labels, bin/pipeline.py:1064-1168) for every contig of
 * a label array laid out by win_offsets.  Contigs with a single distinct label are copied.  noise_labels: bit l set =
 * label l is treated as noise (the host passes all but the two most frequent labels of the whole input,
 * pipeline.py:1084-1086).  status[c]: 0 ok, 1 = the reference raises IndexError on this input.
 */
int mgc_merge_labels(mgc_engine *e, const int32_t *labels, const int64_t *win_offsets, int32_t n_contigs,
                     int32_t consecutive, uint32_t noise_labels, int32_t *merged, int32_t *status);

/* ---- testing hook: the pruned candidate list (total > 0) of the last decode, in chain order ----
 * Mirrors what `gmhmmp2 --debug 1` prints per start point; used by the parity tests only. */
typedef struct {
    int32_t contig, left, right;   /* 0-based inclusive, forward coordinates */
    int32_t strand, gc, domain, partial_start, partial_stop;
    double total, cum;
} mgc_candidate;
int64_t mgc_debug_candidates(mgc_engine *e, int model, mgc_candidate *out, int64_t cap);

/* ---- text writers (gmhmmp2 Output::PrintGenes_GFF / _LST; bin/data.py reads the GFF back) ---- */
int mgc_write_gff(const mgc_result *r, int model, const char *const *contig_ids, const int64_t *contig_len,
                  int32_t n_contigs, const char *seq_file_name, const char *mod_file_name, int gcode,
                  const char *out_path, const char *format /* "gff" | "lst" */);

/* ---- FASTA ingest (bin/data.py:108-124, Biopython SeqIO.parse) — host only ----
 * Reads a (multi-)FASTA file into ONE contiguous letter buffer + offsets, the layout mgc_decode_batch takes.  Record id =
 * first whitespace-delimited token of the defline (Biopython's rec.id); records keep file order (gmhmmp2's output order).
 * page_locked != 0 asks for page-locked memory so that the H2D copy can overlap (falls back to ordinary pages). */
typedef struct mgc_fasta mgc_fasta;
int mgc_fasta_read(const char *path, int page_locked, mgc_fasta **out);
int32_t mgc_fasta_num_records(const mgc_fasta *f);
const char *mgc_fasta_id(const mgc_fasta *f, int32_t i);
const char *mgc_fasta_defline(const mgc_fasta *f, int32_t i);
/* all ids (deflines != 0: all deflines) in one buffer, each followed by '\n'; *len = its size.  Valid until mgc_fasta_free. */
const char *mgc_fasta_ids_joined(const mgc_fasta *f, int deflines, int64_t *len);
/* the ids as an array of n C strings — what mgc_write_gff takes as contig_ids — without a string per record on the caller's
 * side.  Valid until mgc_fasta_free. */
const char *const *mgc_fasta_id_array(const mgc_fasta *f);
const char *mgc_fasta_sequence(const mgc_fasta *f);      /* concatenated letters, no separators */
const int64_t *mgc_fasta_offsets(const mgc_fasta *f);    /* n + 1 entries */
void mgc_fasta_free(mgc_fasta *f);

#ifdef __cplusplus
}
#endif
#endif
