// mgc_kernels.hpp — launch wrappers implemented in mgc_kernels.cu
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mgcod_b200.h"

namespace mgc {

struct DevModel;
struct OrfRec;
struct StartRec;
struct ItemRec;
struct Cand;
struct Planes;
struct ChainBufs;

int launch_exclusive_scan(const uint32_t *in, uint32_t *out, uint32_t *tile_sums, int64_t n, uint32_t *grand_total,
                          cudaStream_t st);
void launch_bucket_totals(const uint32_t *const *bend_m, int64_t nb, unsigned long long *counters_m, int n_models, cudaStream_t st);
int launch_exclusive_scan_multi(const uint32_t *const *in_m, uint32_t *const *out_m, int n_arrays, uint32_t *tile_sums, int64_t n,
                                uint32_t *grand_totals, int total_stride, cudaStream_t st);
void launch_groups(const int64_t *cwoff, int32_t n_contigs, int64_t n_groups, int32_t *grp2contig, cudaStream_t st);
void launch_pack(const char *ascii, const int64_t *aoff, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig,
                 int64_t W, uint32_t *lo, uint32_t *hi, uint32_t *nm, uint32_t *gccount, uint32_t *ncount, int *dflags,
                 cudaStream_t st);
void launch_recoded_stops(Planes pl, const int64_t *cwoff, const int32_t *clen, const GeneRef *genes, int64_t n_genes, int32_t *counts,
                          const int64_t *pos_off, int32_t *pos_out, bool fill, int n_sm, cudaStream_t st);
void launch_codon_planes(Planes pl, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig, uint32_t *codon, cudaStream_t st);
void launch_find_gaps(const uint32_t *nm, const int64_t *cwoff, const int32_t *clen, int32_t n_contigs, int min_run, int32_t *gaps,
                      unsigned int *gap_count, unsigned int cap, cudaStream_t st);
int launch_laststop(Planes pl, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig, const DevModel *models,
                    int n_models, uint32_t *const *ls_m, uint32_t *const *ls_tile_m, cudaStream_t st);
void launch_orf_find(Planes pl, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig, const DevModel *models,
                     int n_models, const uint32_t *const *ls_m, const uint32_t *const *ls_tile_m, OrfRec *const *orfs_m,
                     unsigned long long *counters_m, uint32_t cap_orf, cudaStream_t st);
int launch_orf_prep(Planes pl, const uint32_t *gcpre, const int64_t *cwoff, const int32_t *clen, const DevModel *models, int n_models,
                    OrfRec *const *orfs_m, StartRec *const *starts_m, ItemRec *const *items_m, uint32_t *const *hist_m,
                    unsigned long long *counters_m, const ChainBufs *bufs_m, uint32_t cap_orf, uint32_t cap_starts, uint32_t cap_items,
                    int n_sm, cudaStream_t st);
int launch_item_sort(const ItemRec *const *items_m, ItemRec *const *sorted_m, uint32_t *const *hist_m, uint32_t *const *item_off_m,
                     uint32_t *const *unit_off_m, const DevModel *models, int n_models, uint32_t *unit_totals,
                     const unsigned long long *counters_m, uint32_t cap_items, int n_tiles, int n_sm, bool pair_walk, cudaStream_t st);
int launch_walk(Planes pl, const int64_t *cwoff, const DevModel *models, int n_models, int max_cod_order, int max_non_order,
                const OrfRec *const *orfs_m,
                const StartRec *const *starts_m, const ItemRec *const *items_m, const uint32_t *const *item_off_m,
                const uint32_t *const *unit_off_m, const uint32_t *unit_totals, void *const *sums_m, unsigned int *next_unit, int n_sm,
                int pair_smem /* bytes of the pair walk's table, 0 = single-position walk */,
                bool pair_orders_uniform /* every model has exactly the orders max_cod_order / max_non_order */, cudaStream_t st);
void launch_final(Planes pl, const int64_t *cwoff, const int32_t *clen, const DevModel *models, int n_models, const OrfRec *const *orfs_m,
                  const StartRec *const *starts_m, void *const *sums_m, const unsigned long long *counters_m, Cand *const *surv_m,
                  uint32_t *const *bcount_m, uint32_t *const *cover_m, unsigned long long *surv_counters, uint32_t cap_starts, uint32_t cap_surv, int n_sm,
                  bool pair_walk, cudaStream_t st);
void launch_scatter(const Cand *const *surv_m, Cand *const *sorted_m, uint32_t *const *boff_m,
                    const unsigned long long *surv_counters, const int64_t *cwoff, uint32_t cap_surv, int n_models, int n_sm,
                    cudaStream_t st);
// regions cleared by one k_clear launch (whole 32-bit words, 16-byte aligned bases)
struct ClearList {
    static constexpr int kMax = 64;
    void *ptr[kMax];
    size_t words[kMax];
    uint32_t value[kMax];
    int n = 0;
    void add(void *p, size_t bytes, uint32_t v = 0) {
        if (n < kMax && bytes) { ptr[n] = p; words[n] = (bytes + 3) / 4; value[n] = v; n++; }
    }
};
void launch_clear(const ClearList &list, int n_sm, cudaStream_t st);
void launch_piece_stats(const ChainBufs *bufs_m, const int *work_counters, int32_t n_slots, unsigned int *hist, cudaStream_t st);
int launch_chain(const ChainBufs *bufs_m, const Cand *const *scattered_m, Cand *const *ordered_m, const int64_t *cwoff, const int32_t *clen,
                 const int32_t *slotbase, const int32_t *slot_contig, int32_t n_slots, int64_t n_buckets, int *work_counters, int n_models,
                 int n_sm, int chain_mode /* 0 long pieces by warps + the others by threads, 1 warps only, 2 threads only */, cudaStream_t st);
int launch_gather(const ChainBufs *bufs_m, mgc_gene *const *genes_m, const int32_t *slot_contig, int32_t n_slots, const int32_t *slotbase,
                  int32_t n_contigs, long long *model_cents, int n_models, int n_sm, cudaStream_t st);
void launch_windows(const mgc_gene *const *genes_m, const uint32_t *const *gene_off_m, int n_models, const int32_t *clen,
                    const int32_t *stride, const int64_t *win_off, int32_t n_contigs, int32_t window, int32_t consecutive,
                    int short_contigs, int32_t *labels, long long *cents, unsigned long long *label_hist /* [MGC_MAX_MODELS], may be null */,
                    cudaStream_t st);
// noise_mask_dev != null: the mask is read from device memory (written by launch_noise_mask) instead of noise_mask
void launch_merge_labels(const int32_t *labels, const int64_t *win_off, int32_t n_contigs, int32_t consecutive, uint32_t noise_mask,
                         const uint32_t *noise_mask_dev, int32_t *out, int32_t *run, int32_t *status, cudaStream_t st);
void launch_noise_mask(const unsigned long long *label_hist, int n_models, uint32_t *mask_out, cudaStream_t st);

}  // namespace mgc
