/* gmhmm_oracle.h — C interface of the CPU oracle (TEST INFRASTRUCTURE; see gmhmm_oracle.c). */
#ifndef GMHMM_ORACLE_H
#define GMHMM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;

/* one scored start/stop pair that survived pruning (total > 0); coordinates 0-based inclusive */
typedef struct {
    long left, right;
    int strand;            /* +1 / -1 */
    int len;               /* nt incl. both end codons */
    int gc;                /* round-half-up 100*GC/len */
    int domain;            /* 0 bacteria, 1 archaea */
    int partial_start, partial_stop;
    double total;          /* log-odds of the candidate (GFF score) */
    double cum;            /* best chain score ending with this candidate (filled by orc_chain) */
    long prev;             /* index of the predecessor candidate or -1 */
} orc_cand;

/* one predicted gene; coordinates 1-based inclusive as in GFF */
typedef struct {
    long left, right;
    int strand;
    int domain;
    int gc;
    int partial_left, partial_right;
    double logodd;
} orc_gene;

int orc_model_load(const char *path, orc_model **out);
void orc_model_free(orc_model *m);
int orc_model_gcode(const orc_model *m);
const char *orc_last_error(void);

/* candidates sorted by (right, strand, left); caller frees with orc_free */
long orc_candidates(const orc_model *m, const char *seq, long len, orc_cand **out);
/* fills cum/prev, writes the chain (indices into c) and returns its length */
long orc_chain(orc_cand *c, long n, long *chain, double *total_out);
/* sequence -> genes; returns gene count, caller frees *genes with orc_free */
long orc_decode(const orc_model *m, const char *seq, long len, orc_gene **genes, double *total_logodd);
void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
