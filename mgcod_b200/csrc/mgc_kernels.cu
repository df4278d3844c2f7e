// mgc_kernels.cu — the CUDA path (sm_100a) of the gene finder.
//
// Pipeline for one batch of contigs and M parameter files (all launches on one stream, no host
// synchronisation until the final counters are read):
//
//   k_groups         word-group -> contig map                                  (once per upload)
//   k_clear          every per-decode counter / histogram in one launch (no cudaMemset: it runs on a copy engine)
//   k_pack           ASCII -> 2-bit planes + ambiguity plane for both strands, GC / N counts per word; scan x2
//   k_codon_planes   start / ATG / TAA / TAG / TGA bit per letter, shared by the models
//   k_laststop_*     per model: last in-frame stop before every word (max-scan)
//   k_orf_find       per model: stop-delimited ORFs long enough to hold a gene and with a start codon
//   k_orf_expand     per model: start records (position, GC bin, section pair) of every ORF, load-balanced per block
//   k_start_link     chains the starts of an ORF that share a section pair; one walk item per chain
//   k_item_plan / k_item_scatter   counting sort of the items by (position tile, section pair, length class)
//   k_walk_pair      thread per item: 3-periodic Markov log-likelihood sums from the stop down, two positions per lookup, the
//                    pair's tables (brought in by cp.async.bulk) in the SM's whole shared memory; k_walk: single positions (order 5)
//   k_final          per start: order ramp, duration / start / stop terms, domain choice, pruning (total > 0), bucket histogram
//   scan + k_scatter survivors into position buckets (128 nt) -> right-end order per contig
//   k_bucket_order   candidates into chain order inside their buckets (rank among the bucket's members)
//   k_chain_plan     cut points: pieces that no candidate spans
//   k_chain_flat     max-plus chain with overlap penalties, a THREAD per piece, the lanes of a warp as state machines that take
//                    turns section by section; k_chain (a warp per piece, round 1) for small batches / their long pieces
//   k_chain_trace    traceback, thread per piece
//   scan + k_gather + k_contig_finish   genes ordered by (contig, left), per-model sums in 1/100 units, per-contig totals
//   on request: k_windows, k_merge_labels (window labels), k_recoded_stops (in-frame stops per gene), k_find_gaps
//
// Arithmetic follows SURVEY.md §A.4-A.6 as pinned by oracle/ (never linked here).
#include <cstdint>
#include <cstdlib>
#include <cstdio>

#include "mgc_device.cuh"
#include "mgc_kernels.hpp"

namespace mgc {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}
__device__ __forceinline__ long long warp_min_ll(long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

// 2*(G+C) + (ambiguous letters) over strand-0 positions [a, b] of the contig slot starting at word `w0`; `pre` is the
// exclusive per-word prefix of the same quantity.  (An ambiguous letter counts as half a G/C in the reference's GC %.)
__device__ __forceinline__ int gc2_count_fwd(const Planes &pl, const uint32_t *pre, int64_t w0, int a, int b) {
    const int64_t wa = w0 + (a >> 5), wb = w0 + ((b + 1) >> 5);
    const uint32_t ma = (1u << (a & 31)) - 1u, mb = (1u << ((b + 1) & 31)) - 1u;
    int pa = (int)pre[wa] + 2 * __popc((pl.lo[wa] ^ pl.hi[wa]) & ma);
    int pb = (int)pre[wb] + 2 * __popc((pl.lo[wb] ^ pl.hi[wb]) & mb);
    if (pl.dflags[1]) {
        pa += __popc(pl.nm[wa] & ma);
        pb += __popc(pl.nm[wb] & mb);
    }
    return pb - pa;
}

// number of ambiguous letters in strand-0 positions [a, b]
__device__ __forceinline__ int amb_count_fwd(const Planes &pl, int64_t w0, int a, int b) {
    const int64_t wa = w0 + (a >> 5), wb = w0 + ((b + 1) >> 5);
    const uint32_t ma = (1u << (a & 31)) - 1u, mb = (1u << ((b + 1) & 31)) - 1u;
    return (int)pl.npre[wb] + __popc(pl.nm[wb] & mb) - (int)pl.npre[wa] - __popc(pl.nm[wa] & ma);
}

// ------------------------------------------------------------------------------------------------
// generic exclusive scan of uint32 (three launches; n up to 2^31)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[32];
    int lane = lane_id(), wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        wsum[lane] = s;
    }
    __syncthreads();
    uint32_t base = wid ? wsum[wid - 1] : 0;
    *total = wsum[(blockDim.x >> 5) - 1];
    uint32_t r = base + x - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t *in, uint32_t *out, uint32_t *tile_sums, int64_t n) {
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t *tile_sums, int64_t nt, uint32_t *grand_total) {
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < nt; b += 1024) {
        int64_t i = b + threadIdx.x;
        uint32_t v = i < nt ? tile_sums[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (i < nt) tile_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t *out, const uint32_t *tile_sums, int64_t n) {
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] += add;
}

// the same for several arrays of equal length in one launch each (blockIdx.y = array): in_m / out_m are device tables of
// pointers, tile_sums holds nt entries per array, grand_totals (optional) one per array with the given stride in uint32
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles_multi(const uint32_t *const *in_m, uint32_t *const *out_m, uint32_t *tile_sums,
                                                                   int64_t n, int64_t nt) {
    const uint32_t *in = in_m[blockIdx.y];
    uint32_t *out = out_m[blockIdx.y];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) tile_sums[(int64_t)blockIdx.y * nt + blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums_multi(uint32_t *tile_sums, int64_t nt, uint32_t *grand_totals, int total_stride) {
    __shared__ uint32_t carry;
    uint32_t *ts = tile_sums + (int64_t)blockIdx.x * nt;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < nt; b += 1024) {
        int64_t i = b + threadIdx.x;
        uint32_t v = i < nt ? ts[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (i < nt) ts[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && grand_totals) grand_totals[(int64_t)blockIdx.x * total_stride] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add_multi(uint32_t *const *out_m, const uint32_t *tile_sums, int64_t n, int64_t nt) {
    uint32_t *out = out_m[blockIdx.y];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t add = tile_sums[(int64_t)blockIdx.y * nt + blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] += add;
}

int launch_exclusive_scan_multi(const uint32_t *const *in_m, uint32_t *const *out_m, int n_arrays, uint32_t *tile_sums, int64_t n,
                                uint32_t *grand_totals, int total_stride, cudaStream_t st) {
    if (n <= 0 || n_arrays <= 0) return 0;
    const int64_t nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles_multi<<<dim3((unsigned)nt, n_arrays), SCAN_THREADS, 0, st>>>(in_m, out_m, tile_sums, n, nt);
    k_scan_sums_multi<<<n_arrays, 1024, 0, st>>>(tile_sums, nt, grand_totals, total_stride);
    k_scan_add_multi<<<dim3((unsigned)nt, n_arrays), SCAN_THREADS, 0, st>>>(out_m, tile_sums, n, nt);
    return 3;
}

// after the in-place exclusive scan over nb + 2 entries (the last two were zero) entry nb holds the number of candidates
__global__ void k_bucket_totals(const uint32_t *const *bend_m, int64_t nb, unsigned long long *counters_m) {
    const int mi = threadIdx.x;
    counters_m[(size_t)mi * 8 + 4] = bend_m[mi][nb];
}

void launch_bucket_totals(const uint32_t *const *bend_m, int64_t nb, unsigned long long *counters_m, int n_models, cudaStream_t st) {
    k_bucket_totals<<<1, n_models, 0, st>>>(bend_m, nb, counters_m);
}

int launch_exclusive_scan(const uint32_t *in, uint32_t *out, uint32_t *tile_sums, int64_t n, uint32_t *grand_total,
                          cudaStream_t st) {
    if (n <= 0) return 0;
    int64_t nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<<<(unsigned)nt, SCAN_THREADS, 0, st>>>(in, out, tile_sums, n);
    k_scan_sums<<<1, 1024, 0, st>>>(tile_sums, nt, grand_total);
    k_scan_add<<<(unsigned)nt, SCAN_THREADS, 0, st>>>(out, tile_sums, n);
    return 3;
}

// ------------------------------------------------------------------------------------------------
// upload-time kernels
// ------------------------------------------------------------------------------------------------
// grp2contig[g] = contig owning word group g (4 words); cwoff ascending, multiples of 4
__global__ void k_groups(const int64_t *cwoff, int32_t n_contigs, int64_t n_groups, int32_t *grp2contig) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int64_t w = g * 4;
    int lo = 0, hi = n_contigs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (cwoff[mid] <= w) lo = mid; else hi = mid - 1;
    }
    grp2contig[g] = lo;
}

// one warp packs 32 words (1024 letters) of one strand: lane = bit position, ballot = plane word.
// Letters outside ACGT set the ambiguity plane (U included: the reference scores it like N); their code is 0.  Letters
// outside the IUPAC nucleotide alphabet (which the reference rejects: "error, unexpected letter found in FASTA sequence")
// raise flag 0.
//
// Thread per (strand, plane word).  The 32 letters of a word are handled four at a time inside 32-bit registers: the
// code bits are bits 1 and 2 of the ASCII letter (A 0x41, C 0x43, G 0x47, T 0x54: hi = b2, lo = b1 ^ b2), the ACGT test
// is a bit formula on the upper-cased letter, and a multiply gathers one bit per byte into a nibble.  A reverse-strand
// word is the bit-reversed complement of the forward window that mirrors it.  The ASCII buffer has PACK_SLACK bytes of
// slack on both sides so that the aligned 36-byte read of a window may start before / end after the contig.
__device__ __forceinline__ uint32_t nibble_of(uint32_t v) {  // v has one bit per byte (bit 0 of each): -> bit j = byte j
    return (v * 0x01020408u) >> 24;
}

__device__ __forceinline__ bool iupac_letter(unsigned char up) {
    return up == 'N' || up == 'U' || up == 'R' || up == 'Y' || up == 'S' || up == 'W' || up == 'K' || up == 'M' || up == 'B' ||
           up == 'D' || up == 'H' || up == 'V';
}

__global__ void __launch_bounds__(256) k_pack(const char *__restrict__ ascii, const int64_t *__restrict__ aoff,
                                              const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                              const int32_t *__restrict__ grp2contig, int64_t W, uint32_t *lo,
                                              uint32_t *hi, uint32_t *nm, uint32_t *gccount, uint32_t *ncount, int *dflags) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int strand = blockIdx.y;
    if (w >= W) return;
    const int c = grp2contig[w >> 2];
    const int L = clen[c];
    const int64_t wl = w - cwoff[c];
    // forward window [p, p + 32) of the contig that this word shows (mirrored for the reverse strand)
    const int64_t p = strand ? (int64_t)L - 32 * (wl + 1) : 32 * wl;
    uint32_t f_lo = 0, f_hi = 0, f_nm = 0, valid = 0;
    if (p < L && p > -32) {
        const int from = p < 0 ? (int)-p : 0, to = p + 31 >= L ? (int)(L - 1 - p) : 31;
        valid = bit_range(from, to);
        const char *a = ascii + aoff[c] + p;
        const uintptr_t ua = reinterpret_cast<uintptr_t>(a);
        const uint32_t *q = reinterpret_cast<const uint32_t *>(ua & ~(uintptr_t)3);
        const int sh8 = (int)(ua & 3) * 8;
        uint32_t wv[9];
#pragma unroll
        for (int k = 0; k < 9; k++) wv[k] = __ldg(q + k);
        uint32_t notacgt = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t x = __funnelshift_r(wv[k], wv[k + 1], sh8);   // letters 4k .. 4k+3, first letter in the low byte
            const uint32_t b0 = x, b1 = x >> 1, b2 = x >> 2, b3 = x >> 3, b4 = x >> 4, b6 = x >> 6, b7 = x >> 7;   // bit 5 = case
            const uint32_t acg = ~b4 & b0 & ~(b2 & ~b1), t = b4 & b2 & ~b1 & ~b0;
            const uint32_t ok = ~b7 & b6 & ~b3 & (acg | t) & 0x01010101u;
            f_hi |= nibble_of(b2 & 0x01010101u) << (4 * k);
            f_lo |= nibble_of((b1 ^ b2) & 0x01010101u) << (4 * k);
            notacgt |= nibble_of(ok ^ 0x01010101u) << (4 * k);
        }
        f_nm = notacgt & valid;
        if (f_nm) {   // rare: tell IUPAC codes from letters the reference rejects
            bool bad = false;
            uint32_t m = f_nm;
            while (m) {
                const int bpos = __ffs(m) - 1;
                m &= m - 1;
                if (!iupac_letter((unsigned char)a[bpos] & 0xDF)) bad = true;
            }
            if (!strand) {
                atomicOr(&dflags[1], 1);
                if (bad) atomicOr(&dflags[0], 1);
            }
        }
    }
    const uint32_t keep = valid & ~f_nm;
    uint32_t o_lo, o_hi, o_nm;
    if (strand) {
        o_lo = __brev(~f_lo & keep); o_hi = __brev(~f_hi & keep); o_nm = __brev(f_nm);
    } else {
        o_lo = f_lo & keep; o_hi = f_hi & keep; o_nm = f_nm;
    }
    lo[(int64_t)strand * W + w] = o_lo;
    hi[(int64_t)strand * W + w] = o_hi;
    nm[(int64_t)strand * W + w] = o_nm;
    if (!strand) {
        gccount[w] = 2u * __popc(o_lo ^ o_hi) + __popc(o_nm);   // padding letters are A (code 0): they count nothing
        ncount[w] = __popc(o_nm);
    }
}

// ------------------------------------------------------------------------------------------------
// last in-frame stop table:  ls[w*3 + f] = 1 + position of the last stop codon at or before the end of
// word w whose position is congruent to f mod 3 (positions count letters of the whole two-strand slot
// space; 0 = none).  A plain max-scan over words; consumers compare with their contig's first position.
// ------------------------------------------------------------------------------------------------
struct FindCtx {
    const uint32_t *cp;       // codon planes of the strand region at the contig slot (plane p at cp + p * cstride)
    int64_t cstride;
    int L;                    // contig length
    int stop_sel, atg_only;
};

// Codon planes (k_codon_planes): bit b of word w = "a codon of that kind starts at letter 32w + b".  Codon starts beyond
// L - 3 and codons holding an ambiguous letter are already masked out.
enum { CP_START3 = 0, CP_ATG = 1, CP_TAA = 2, CP_TAG = 3, CP_TGA = 4, CP_PLANES = 5 };

__device__ __forceinline__ FindCtx make_ctx(const Planes &pl, int strand, int64_t w0, int L, int stop_sel, int atg_only) {
    FindCtx cx;
    cx.cp = pl.codon + (int64_t)strand * pl.W + w0;
    cx.cstride = 2 * pl.W;
    cx.L = L; cx.stop_sel = stop_sel; cx.atg_only = atg_only;
    return cx;
}

__device__ __forceinline__ uint32_t starts_at(const FindCtx &cx, int w) { return cx.cp[(cx.atg_only ? CP_ATG : CP_START3) * cx.cstride + w]; }

__device__ __forceinline__ uint32_t stops_at(const FindCtx &cx, int w) {   // the stop codons of the model's genetic code
    uint32_t m = 0;
    if (cx.stop_sel & 1) m |= cx.cp[CP_TAA * cx.cstride + w];
    if (cx.stop_sel & 2) m |= cx.cp[CP_TAG * cx.cstride + w];
    if (cx.stop_sel & 4) m |= cx.cp[CP_TGA * cx.cstride + w];
    return m;
}

__device__ __forceinline__ CodonMasks masks_at(const FindCtx &cx, int w) {
    CodonMasks m;
    m.start = starts_at(cx, w);
    m.taa = cx.cp[CP_TAA * cx.cstride + w]; m.tag = cx.cp[CP_TAG * cx.cstride + w]; m.tga = cx.cp[CP_TGA * cx.cstride + w];
    return m;
}

// thread per (strand, plane word): the five codon planes from the letter planes, once for all models
__global__ void __launch_bounds__(256) k_codon_planes(Planes pl, const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                                      const int32_t *__restrict__ grp2contig, uint32_t *codon) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int strand = blockIdx.y;
    if (w >= pl.W) return;
    const int c = grp2contig[w >> 2];
    const int L = clen[c];
    const int wl = (int)(w - cwoff[c]);
    const int64_t at = (int64_t)strand * pl.W + w;
    const uint32_t l0 = pl.lo[at], h0 = pl.hi[at], l1 = pl.lo[at + 1], h1 = pl.hi[at + 1];
    CodonMasks m = codon_masks(l0, h0, l1, h1, 0);
    const uint32_t atg = codon_masks(l0, h0, l1, h1, 1).start;
    // codon start positions must satisfy i <= L-3
    const int last = L - 3 - 32 * wl;
    uint32_t valid = last < 0 ? 0u : (last >= 31 ? 0xffffffffu : ((1u << (last + 1)) - 1u));
    if (pl.dflags[1]) {  // a codon with an ambiguous letter is neither a start nor a stop
        const uint32_t n0 = pl.nm[at], n1 = pl.nm[at + 1];
        valid &= ~(n0 | __funnelshift_r(n0, n1, 1) | __funnelshift_r(n0, n1, 2));
    }
    const int64_t cs = 2 * pl.W;
    codon[CP_START3 * cs + at] = m.start & valid;
    codon[CP_ATG * cs + at] = atg & valid;
    codon[CP_TAA * cs + at] = m.taa & valid;
    codon[CP_TAG * cs + at] = m.tag & valid;
    codon[CP_TGA * cs + at] = m.tga & valid;
}

struct U3 {
    uint32_t v[3];
};
__device__ __forceinline__ U3 max3(U3 a, U3 b) {
    U3 r;
    r.v[0] = max(a.v[0], b.v[0]); r.v[1] = max(a.v[1], b.v[1]); r.v[2] = max(a.v[2], b.v[2]);
    return r;
}

constexpr int LS_THREADS = 256;

// grid.x = tiles of 256 words over the 2W words of both strands, grid.y = model
__global__ void __launch_bounds__(LS_THREADS) k_laststop_tiles(Planes pl, const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                                               const int32_t *__restrict__ grp2contig, const DevModel *__restrict__ models,
                                                               uint32_t *const *ls_m, uint32_t *const *ls_tile_m) {
    __shared__ U3 wsum[LS_THREADS / 32];
    const int mi = blockIdx.y;
    const int64_t w = (int64_t)blockIdx.x * LS_THREADS + threadIdx.x;  // word in [0, 2W)
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    U3 x = {{0, 0, 0}};
    if (w < 2 * pl.W) {
        const int strand = w >= pl.W;
        const int64_t ws = w - (strand ? pl.W : 0);
        const int c = grp2contig[ws >> 2];
        const int wl = (int)(ws - cwoff[c]);
        const int L = clen[c];
        if (L >= 3 && wl * 32 <= L - 3) {
            const FindCtx cx = make_ctx(pl, strand, cwoff[c], L, models[mi].stop_sel, models[mi].atg_only);
            const uint32_t sm = stops_at(cx, wl);
            const uint32_t P0 = (uint32_t)(w * 32);
            const int r0 = (int)(P0 % 3u);
#pragma unroll
            for (int f = 0; f < 3; f++) {
                const uint32_t m = sm & frame_pattern((f - r0 + 3) % 3);
                x.v[f] = m ? P0 + (31u - (uint32_t)__clz(m)) + 1u : 0u;
            }
        }
    }
    // inclusive max-scan within the warp, then across warps
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        U3 y;
        y.v[0] = __shfl_up_sync(0xffffffffu, x.v[0], o); y.v[1] = __shfl_up_sync(0xffffffffu, x.v[1], o);
        y.v[2] = __shfl_up_sync(0xffffffffu, x.v[2], o);
        if (lane >= o) x = max3(x, y);
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    U3 carry = {{0, 0, 0}};
    for (int k = 0; k < wid; k++) carry = max3(carry, wsum[k]);
    x = max3(x, carry);
    uint32_t *ls = ls_m[mi];
    if (w < 2 * pl.W) { ls[w * 3 + 0] = x.v[0]; ls[w * 3 + 1] = x.v[1]; ls[w * 3 + 2] = x.v[2]; }
    if (threadIdx.x == LS_THREADS - 1) {
        uint32_t *lt = ls_tile_m[mi];
        lt[(int64_t)blockIdx.x * 3 + 0] = x.v[0]; lt[(int64_t)blockIdx.x * 3 + 1] = x.v[1]; lt[(int64_t)blockIdx.x * 3 + 2] = x.v[2];
    }
}

// one block per model: running maximum over the tile results (exclusive: tile t receives the max of tiles < t)
__global__ void __launch_bounds__(1024) k_laststop_sums(uint32_t *const *ls_tile_m, int64_t n_tiles) {
    __shared__ U3 wsum[32];
    __shared__ U3 carry_s;
    uint32_t *lt = ls_tile_m[blockIdx.x];
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = U3{{0, 0, 0}};
    __syncthreads();
    for (int64_t b = 0; b < n_tiles; b += 1024) {
        const int64_t i = b + threadIdx.x;
        U3 own = {{0, 0, 0}};
        if (i < n_tiles) { own.v[0] = lt[i * 3]; own.v[1] = lt[i * 3 + 1]; own.v[2] = lt[i * 3 + 2]; }
        U3 x = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            U3 y;
            y.v[0] = __shfl_up_sync(0xffffffffu, x.v[0], o); y.v[1] = __shfl_up_sync(0xffffffffu, x.v[1], o);
            y.v[2] = __shfl_up_sync(0xffffffffu, x.v[2], o);
            if (lane >= o) x = max3(x, y);
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        U3 c = carry_s;
        for (int k = 0; k < wid; k++) c = max3(c, wsum[k]);
        // exclusive value = max over everything before me = (inclusive of previous lane) max carry
        U3 prev;
        prev.v[0] = __shfl_up_sync(0xffffffffu, x.v[0], 1); prev.v[1] = __shfl_up_sync(0xffffffffu, x.v[1], 1);
        prev.v[2] = __shfl_up_sync(0xffffffffu, x.v[2], 1);
        if (lane == 0) prev = U3{{0, 0, 0}};
        const U3 excl = max3(prev, c);
        if (i < n_tiles) { lt[i * 3] = excl.v[0]; lt[i * 3 + 1] = excl.v[1]; lt[i * 3 + 2] = excl.v[2]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = max3(x, c);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// ORF discovery: thread per word; the previous in-frame stop comes from the table in O(1)
// ------------------------------------------------------------------------------------------------
// Builds the record of the ORF ending at stop position t if it can hold a start codon; false otherwise.
__device__ __forceinline__ bool make_orf(const FindCtx &cx, const DevModel &dm, int contig, int strand, int t, bool virt,
                                         int p_prev /* local, < 0 = none in this contig */, bool right_is_record_end, OrfRec &rec) {
    const int f = t % 3;
    const int s_hi = min(t - 3, t + 3 - dm.min_len);
    const bool open_left = p_prev < 0;
    const int s_lo = open_left ? f : p_prev + 3;
    if (s_lo > s_hi) return false;
    if (!open_left) {
        // cheap existence test for a start codon: look at up to four words of the admissible range
        bool any = false;
        const int w_end = min(s_hi >> 5, (s_lo >> 5) + 3);
        for (int w = s_lo >> 5; w <= w_end && !any; w++) {
            const int r = (f + w) % 3;   // (f - 32 w) mod 3: 32 = 2 = -1 (mod 3); w >= 0
            any = (starts_at(cx, w) & frame_pattern(r) & bit_range(s_lo - 32 * w, s_hi - 32 * w)) != 0;
        }
        if (!any && (s_hi >> 5) <= w_end) return false;
    }
    rec.contig = contig; rec.t = t; rec.s_lo = s_lo; rec.base = 0; rec.n = 0;
    // the strand's left edge is the forward right edge for the reverse strand
    const bool from_first = open_left && !(strand == 1 && right_is_record_end);
    rec.sf = (uint32_t)strand | (open_left ? 2u : 0u) | (virt ? 0x100u : 0u) | (from_first ? 0x200u : 0u);
    return true;
}

// grid.x over the 2W words of both strands, grid.y = model.
// Phase A, lane = word: stop mask, previous in-frame stops, and the cheap length test (distance to the previous in-frame
// stop >= GENE_MIN_LENGTH; with a minimum length above one word only the first stop of each frame in the word can pass).
// Survivors go to a per-warp queue.  Phase B, lane = queued stop: the start-codon existence test and the record — dense
// lanes instead of a divergent per-word loop.
constexpr int FIND_Q = 6;   // queued stops per word: 3 frames + 3 virtual stops at the right edge
__global__ void __launch_bounds__(128) k_orf_find(Planes pl, const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                                  const int32_t *__restrict__ grp2contig, const DevModel *__restrict__ models,
                                                  const uint32_t *const *ls_m, const uint32_t *const *ls_tile_m, OrfRec *const *orfs_m,
                                                  unsigned long long *counters_m, uint32_t cap_orf) {
    __shared__ int q_t[4][32 * FIND_Q], q_prev[4][32 * FIND_Q];   // contig-local stop position (bit 31 = virtual), previous stop
    __shared__ unsigned char q_src[4][32 * FIND_Q];
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int mi = blockIdx.y;
    const DevModel &dm = models[mi];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    OrfRec *orfs = orfs_m[mi];
    unsigned long long *counters = counters_m + (size_t)mi * 8;
    // ---- phase A ----
    int my_t[FIND_Q], my_prev[FIND_Q];
    int nq = 0;
    int c = 0, L = 0, strand = 0;
    bool right_end = false;
    if (w < 2 * pl.W) {
        strand = w >= pl.W;
        const int64_t ws = w - (strand ? pl.W : 0);
        c = grp2contig[ws >> 2];
        L = clen[c];
        const int wl = (int)(ws - cwoff[c]);
        if (L >= 6 && wl * 32 <= L - 3) {  // the word holds codon starts (virtual stops sit at t <= L-3 too)
            const FindCtx cx = make_ctx(pl, strand, cwoff[c], L, dm.stop_sel, dm.atg_only);
            const uint32_t *ls = ls_m[mi];
            right_end = (pl.cflags[c] & 1) != 0;
            const uint32_t P0 = (uint32_t)(w * 32);                      // slot-space position of bit 0
            const uint32_t Pc = (uint32_t)(((int64_t)strand * pl.W + cwoff[c]) * 32);  // ... of the contig's first letter
            const uint32_t all = stops_at(cx, wl);
            // previous stop per frame coming from the words before this one
            uint32_t carry[3] = {0, 0, 0};
            if (w > 0) {  // within-tile scan value of the previous word, topped up with the carry into its tile
                const uint32_t *lt = ls_tile_m[mi] + ((w - 1) / LS_THREADS) * 3;
                carry[0] = max(ls[(w - 1) * 3], lt[0]); carry[1] = max(ls[(w - 1) * 3 + 1], lt[1]); carry[2] = max(ls[(w - 1) * 3 + 2], lt[2]);
            }
            auto prev_local = [&](int b) {  // previous in-frame stop of the codon at bit b, contig-local, -1 if none
                const uint32_t P = P0 + (uint32_t)b;
                const int f = (int)(P % 3u);
                const uint32_t below = b ? ((1u << b) - 1u) : 0u;
                const uint32_t m = all & frame_pattern((f - (int)(P0 % 3u) + 3) % 3) & below;
                const uint32_t pp1 = m ? P0 + (31u - (uint32_t)__clz(m)) + 1u : carry[f];
                return pp1 > Pc ? (int)(pp1 - 1u - Pc) : -1;
            };
            auto consider = [&](int b, bool virt) {
                const int t = wl * 32 + b, pp = prev_local(b);
                const int s_lo = pp < 0 ? t % 3 : pp + 3, s_hi = min(t - 3, t + 3 - dm.min_len);
                if (s_lo > s_hi) return;
                if (nq < FIND_Q) { my_t[nq] = t | (virt ? (int)0x80000000 : 0); my_prev[nq] = pp; nq++; }
                else {  // only with a minimum length below one word: finish this one here
                    OrfRec r;
                    if (make_orf(cx, dm, c, strand, t, virt, pp, right_end, r)) {
                        const unsigned long long oid = atomicAdd(&counters[0], 1ull);
                        if (oid < cap_orf) orfs[oid] = r;
                    }
                }
            };
            if (dm.min_len > 34) {
                // a stop with an earlier in-frame stop in the same word is closer than GENE_MIN_LENGTH to it
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const uint32_t m = all & frame_pattern(r);
                    if (m) consider(__ffs(m) - 1, false);
                }
            } else {
                uint32_t sm = all;
                while (sm) {
                    const int b = __ffs(sm) - 1;
                    sm &= sm - 1;
                    consider(b, false);
                }
            }
            // right edge: in each frame the last full codon closes an open ORF unless it is a real stop
            if (wl * 32 + 34 >= L - 3)
                for (int f = 0; f < 3; f++) {
                    const int tv = L - 3 - ((L - 3 - f) % 3 + 3) % 3;  // largest t <= L-3 with t % 3 == f
                    if (tv < 0 || (tv >> 5) != wl) continue;
                    if ((all >> (tv & 31)) & 1u) continue;
                    consider(tv & 31, true);
                }
        }
    }
    // ---- queue ----
    uint32_t pre = (uint32_t)nq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += y;
    }
    const int K = (int)__shfl_sync(0xffffffffu, pre, 31);
    if (K == 0) return;
    pre -= (uint32_t)nq;
    for (int q = 0; q < nq; q++) {
        q_t[wib][pre + q] = my_t[q];
        q_prev[wib][pre + q] = my_prev[q];
        q_src[wib][pre + q] = (unsigned char)lane;
    }
    __syncwarp();
    // ---- phase B ----
    for (int q0 = 0; q0 < K; q0 += 32) {
        const int q = q0 + lane;
        OrfRec r;
        bool ok = false;
        // context of the source word, fetched by all lanes (a warp may span two contigs or the strand boundary)
        const int src = q < K ? q_src[wib][q] : 0;
        const int cs = __shfl_sync(0xffffffffu, c, src);
        const int Ls = __shfl_sync(0xffffffffu, L, src);
        const int ss = __shfl_sync(0xffffffffu, strand, src);
        const int re = __shfl_sync(0xffffffffu, (int)right_end, src);
        if (q < K) {
            const int tq = q_t[wib][q], pp = q_prev[wib][q];
            const FindCtx cx = make_ctx(pl, ss, cwoff[cs], Ls, dm.stop_sel, dm.atg_only);
            ok = make_orf(cx, dm, cs, ss, tq & 0x7fffffff, tq < 0, pp, re != 0, r);
        }
        const uint32_t okm = __ballot_sync(0xffffffffu, ok);
        if (okm) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&counters[0], (unsigned long long)__popc(okm));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned long long at = base + __popc(okm & ((1u << lane) - 1u));
            if (ok && at < cap_orf) orfs[at] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ORF preparation: thread per ORF — start list with GC bin / section pair, walk items
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void window_bits(const uint32_t *lo, const uint32_t *hi, int first, int n, uint32_t &hb,
                                            uint32_t &lb) {
    int w = first >> 5, sh = first & 31;
    uint32_t mask = (1u << n) - 1u;
    lb = __funnelshift_r(lo[w], lo[w + 1], sh) & mask;
    hb = __funnelshift_r(hi[w], hi[w + 1], sh) & mask;
}

// warp-wide exclusive prefix + one atomic reservation for the whole warp; returns this lane's first index
__device__ __forceinline__ unsigned long long warp_reserve(unsigned long long *counter, uint32_t n) {
    const int lane = lane_id();
    uint32_t pre = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += y;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, pre, 31);
    unsigned long long base = 0;
    if (lane == 31 && tot) base = atomicAdd(counter, (unsigned long long)tot);
    return __shfl_sync(0xffffffffu, base, 31) + pre - n;
}

// One atomic per distinct key in the warp instead of one per lane: lanes holding the same key elect a leader that adds
// the group size; returns this lane's slot.  `active` lanes only (all 32 lanes must call).
__device__ __forceinline__ uint32_t warp_key_reserve(uint32_t *counters, uint32_t key, bool active) {
    const int lane = lane_id();
    const uint32_t peers = __match_any_sync(0xffffffffu, active ? key : 0xffffffffu);
    uint32_t base = 0;
    const int leader = __ffs(peers) - 1;
    if (active && lane == leader) base = atomicAdd(&counters[key], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

// prep 1+2, load-balanced: a block takes EXPAND_ORFS ORFs, expands them into their 32-letter words (exclusive scan of
// the word counts, binary search per word), counts the start codons of every word with dense lanes, reserves the tile's
// slice of the start array with one atomic, and writes the start records (position, GC bin, section pair) word by word.
// The flat prefix of the per-word counts in expansion order is at the same time the offset of the word's first start.
constexpr int EXPAND_ORFS = 256;
constexpr int EXPAND_CACHE = 7168;   // expanded words cached between the passes (a tile has ~3300 on gene-dense input)
__global__ void __launch_bounds__(EXPAND_ORFS) k_orf_expand(Planes pl, const uint32_t *__restrict__ gcpre, const int64_t *__restrict__ cwoff,
                                                            const int32_t *__restrict__ clen, const DevModel *__restrict__ models,
                                                            OrfRec *const *orfs_m, StartRec *const *starts_m, unsigned long long *counters_m,
                                                            const ChainBufs *bufs_m, uint32_t cap_orf, uint32_t cap_starts) {
    __shared__ uint32_t o_woff[EXPAND_ORFS + 1];     // first expanded word of each ORF
    __shared__ int o_wfirst[EXPAND_ORFS], o_t[EXPAND_ORFS], o_slo[EXPAND_ORFS], o_shi[EXPAND_ORFS], o_contig[EXPAND_ORFS];
    __shared__ uint32_t o_sf[EXPAND_ORFS], o_ok[EXPAND_ORFS];
    __shared__ unsigned long long s_tile_base;
    __shared__ uint32_t c_mask[EXPAND_CACHE];        // pass 1 -> pass 2: start mask and ORF of each expanded word
    __shared__ unsigned char c_orf[EXPAND_CACHE];
    const int mi = blockIdx.y;
    const DevModel &dm = models[mi];
    unsigned long long *counters = counters_m + (size_t)mi * 8;
    unsigned long long n_orfs = counters[0];
    if (n_orfs > cap_orf) n_orfs = cap_orf;
    OrfRec *orfs = orfs_m[mi];
    StartRec *starts = starts_m[mi];
    int32_t *gptr = bufs_m[mi].gptr;
    const int tid = threadIdx.x;
    for (int64_t tile0 = (int64_t)blockIdx.x * EXPAND_ORFS; tile0 < (int64_t)n_orfs; tile0 += (int64_t)gridDim.x * EXPAND_ORFS) {
        // ---- the tile's ORFs ----
        const int64_t oi = tile0 + tid;
        const bool has = oi < (int64_t)n_orfs;
        uint32_t nwords = 0;
        if (has) {
            const OrfRec rec = orfs[oi];
            gptr[oi] = GPTR_UNSET;
            const int s_hi = min(rec.t - 3, rec.t + 3 - dm.min_len);
            o_wfirst[tid] = rec.s_lo >> 5;
            o_t[tid] = rec.t; o_slo[tid] = rec.s_lo; o_shi[tid] = s_hi; o_contig[tid] = rec.contig; o_sf[tid] = rec.sf;
            nwords = (uint32_t)((s_hi >> 5) - (rec.s_lo >> 5) + 1);
        }
        uint32_t tw;
        const uint32_t wo = block_exclusive_scan(nwords, &tw);
        o_woff[tid] = wo;
        if (tid == 0) o_woff[EXPAND_ORFS] = tw;
        // each ORF stamps its number on its words (cheaper than a binary search per word; words beyond the cache search)
        for (uint32_t q = 0; q < nwords && wo + q < EXPAND_CACHE; q++) c_orf[wo + q] = (unsigned char)tid;
        __syncthreads();
        // the word an expansion index stands for, and its start-codon mask (plus the virtual start in an ORF's first word)
        auto word_of = [&](uint32_t k, int &i, int &w, uint32_t &mask, bool &virt_here) {
            if (k < EXPAND_CACHE) i = c_orf[k];
            else {
                int lo = 0, hi = EXPAND_ORFS;   // last ORF with o_woff <= k
                while (lo + 1 < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (o_woff[mid] <= k) lo = mid; else hi = mid;
                }
                i = lo;
            }
            const int j = (int)(k - o_woff[i]);
            w = o_wfirst[i] + j;
            const int c = o_contig[i], strand = (int)(o_sf[i] & 1u);
            const FindCtx cx = make_ctx(pl, strand, cwoff[c], clen[c], dm.stop_sel, dm.atg_only);
            const int f = o_t[i] % 3;
            const int r = (f + w) % 3;   // (f - 32 w) mod 3: 32 = 2 = -1 (mod 3); w >= 0
            mask = starts_at(cx, w) & frame_pattern(r) & bit_range(o_slo[i] - 32 * w, o_shi[i] - 32 * w);
            virt_here = j == 0 && (o_sf[i] & 2u) != 0;
        };
        // ---- pass 1: starts of the tile; (ORF, mask) of the first EXPAND_CACHE words are kept for pass 2 ----
        uint32_t mine = 0;
        for (uint32_t k = tid; k < tw; k += EXPAND_ORFS) {
            int i, w;
            uint32_t mask;
            bool vh;
            word_of(k, i, w, mask, vh);
            if (k < EXPAND_CACHE) c_mask[k] = mask;
            mine += (uint32_t)__popc(mask) + (vh ? 1u : 0u);
        }
        uint32_t tile_total;
        block_exclusive_scan(mine, &tile_total);
        if (tid == 0) s_tile_base = tile_total ? atomicAdd(&counters[1], (unsigned long long)tile_total) : 0ull;
        __syncthreads();
        const unsigned long long tile_base = s_tile_base;
        const bool tile_ok = tile_base + tile_total <= cap_starts;
        // ---- pass 2: flat prefix in expansion order = offset of each word's first start; write the records ----
        uint32_t carry = 0;
        for (uint32_t k0 = 0; k0 < tw; k0 += EXPAND_ORFS) {
            const uint32_t k = k0 + tid;
            int i = 0, w = 0;
            uint32_t mask = 0;
            bool vh = false;
            if (k < tw) {
                if (k < EXPAND_CACHE) {
                    i = c_orf[k];
                    mask = c_mask[k];
                    w = o_wfirst[i] + (int)(k - o_woff[i]);
                    vh = k == o_woff[i] && (o_sf[i] & 2u) != 0;
                } else word_of(k, i, w, mask, vh);
            }
            const uint32_t cnt = (uint32_t)__popc(mask) + (vh ? 1u : 0u);
            uint32_t chunk_total;
            const uint32_t pre = carry + block_exclusive_scan(cnt, &chunk_total);
            carry += chunk_total;
            if (k < tw) {
                if (k == o_woff[i]) o_ok[i] = pre;   // offset of the ORF's first start inside the tile
                if (tile_ok && cnt) {
                    const int c = o_contig[i], t = o_t[i], strand = (int)(o_sf[i] & 1u);
                    const int L = clen[c];
                    const int64_t w0 = cwoff[c];
                    unsigned long long kk = tile_base + pre;
                    auto put = [&](int sp, bool virt) {
                        const int len = t + 3 - sp;
                        const int a = strand ? (L - 1 - (t + 2)) : sp, b = strand ? (L - 1 - sp) : (t + 2);
                        const int gc2 = gc2_count_fwd(pl, gcpre, w0, a, b);
                        // the binary's own arithmetic, operation for operation (no fused multiply-add): the quotient is
                        // rounded before it is scaled, so exact .5 ties mostly go down (DESIGN.md §2 rule 2)
                        const int gc = (int)__dadd_rn(__dmul_rn(__ddiv_rn((double)gc2, (double)(2 * len)), 100.0), 0.5);
                        StartRec sr;
                        sr.sraw = (uint32_t)sp | (virt ? 0x80000000u : 0u);
                        sr.orf = (uint32_t)(tile0 + i);
                        sr.pg = (uint32_t)dm.pair[gc] | ((uint32_t)gc << 16);
                        sr.link = -1;
                        starts[kk++] = sr;
                    };
                    if (vh) put(t % 3, true);   // the virtual start at the contig edge comes first
                    while (mask) {
                        const int bpos = __ffs(mask) - 1;
                        mask &= mask - 1;
                        put(w * 32 + bpos, false);
                    }
                }
            }
        }
        __syncthreads();
        // ---- the ORFs' slices ----
        if (has) {
            const uint32_t first = o_ok[tid];
            const uint32_t next = tid + 1 < EXPAND_ORFS && tile0 + tid + 1 < (int64_t)n_orfs ? o_ok[tid + 1] : tile_total;
            orfs[oi].base = (uint32_t)(tile_base + first);
            orfs[oi].n = tile_ok ? next - first : 0u;
        }
        __syncthreads();
    }
}

// prep 3, thread per START: links every start to the nearest earlier start of its ORF with the same section pair, and
// the start that has no such later neighbour emits the walk item of its chain.  The starts of an ORF are contiguous, so
// within a warp's 32 consecutive starts the neighbours are found with one __match_any_sync on (ORF, pair) and the head of
// a chain by hopping over lanes with shuffles; only chains that cross the warp's edge look at memory, at most LINK_WIN
// starts away.  "Nearest same-pair neighbour" is a symmetric relation, so both ends of a link agree without talking;
// a pair that reappears further away simply starts a new chain (one more item, same sums).
constexpr int LINK_WIN = 24;
__global__ void __launch_bounds__(256) k_start_link(Planes pl, const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                                    const DevModel *__restrict__ models, const OrfRec *const *orfs_m,
                                                    StartRec *const *starts_m, ItemRec *const *items_m, uint32_t *const *hist_m,
                                                    unsigned long long *counters_m, uint32_t cap_starts, uint32_t cap_items) {
    const int mi = blockIdx.y;
    const DevModel &dm = models[mi];
    unsigned long long *counters = counters_m + (size_t)mi * 8;
    const unsigned long long n_starts = counters[1];
    if (n_starts > cap_starts) return;   // the start array overflowed and is only partly written: the host grows it and decodes again
    const OrfRec *orfs = orfs_m[mi];
    StartRec *starts = starts_m[mi];
    ItemRec *items = items_m[mi];
    uint32_t *hist = hist_m[mi];
    const int kc = dm.kc;
    const int lane = lane_id();
    const bool amb = pl.dflags[1] != 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // item slots are reserved per warp in chunks (one hot counter, see k_final); unfilled slots get key 0xffffffff
    const int chunk = n_starts >= (1ull << 22) ? SURV_CHUNK : 0;
    unsigned long long chunk_base = 0;
    int chunk_left = 0;
    for (int64_t k0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31LL; k0 < (int64_t)n_starts; k0 += stride) {
        const int64_t k = k0 + lane;
        StartRec me;
        OrfRec rec;
        me.sraw = 0; me.orf = 0xffffffffu; me.pg = 0; me.link = -1;
        rec.base = 0; rec.n = 0; rec.t = 0; rec.sf = 0; rec.contig = 0;
        bool valid = false;
        int j = 0, my_full = 0;
        if (k < (int64_t)n_starts) {
            me = starts[k];
            rec = orfs[me.orf];
            j = (int)(k - rec.base);
        }
        const int t = rec.t, n = (int)rec.n;
        const bool from_first = (rec.sf & 0x200u) != 0;
        auto full_of = [&](uint32_t sraw) { return (int)(sraw & 0x7fffffffu) + (((sraw >> 31) && from_first) ? 0 : 3) + kc; };
        const uint32_t pair = me.pg & 0xffffu;
        const int npair = dm.npair;
        if (k < (int64_t)n_starts) {
            my_full = full_of(me.sraw);
            valid = j < n && my_full <= t - 1;   // starts without a full-order region are not part of any walk
        }
        const StartRec *st = starts + rec.base;
        auto prev_in_memory = [&](int from) {   // nearest earlier start of the same pair with a full-order region, -1 if none near
            for (int d = 1; d <= LINK_WIN && from - d >= 0; d++) {
                const StartRec q = st[from - d];
                if ((q.pg & 0xffffu) == pair && full_of(q.sraw) <= t - 1) return from - d;
            }
            return -1;
        };
        // neighbours inside the warp
        const unsigned long long key = valid ? (((unsigned long long)me.orf << 16) | pair) : (0xffffffff00000000ull | (unsigned)lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        const uint32_t below = peers & ((1u << lane) - 1u), above = peers & ~((2u << lane) - 1u);
        int prev = -1;              // index inside the ORF
        int prev_lane = -1;         // its lane if it sits in this warp
        bool has_next = false;
        if (valid) {
            // the same rule everywhere: the nearest same-pair neighbour counts only if it is at most LINK_WIN starts away
            if (below) {
                const int pl_ = 31 - __clz(below);
                if (lane - pl_ <= LINK_WIN) { prev_lane = pl_; prev = j - (lane - pl_); }
            } else if (j > lane) prev = prev_in_memory(j);         // my ORF began before this warp's first start
            if (above) has_next = (__ffs(above) - 1) - lane <= LINK_WIN;
            else if (j + (31 - lane) < n - 1)                      // my ORF continues after this warp's last start
                for (int d = 32 - lane; d <= LINK_WIN && j + d < n && !has_next; d++) {
                    const StartRec q = st[j + d];
                    has_next = (q.pg & 0xffffu) == pair && full_of(q.sraw) <= t - 1;
                }
            if (prev >= 0) starts[k].link = prev;
        }
        // head of my chain: hop over lanes while the chain stays in the warp, then over memory
        int lo_pos = my_full, hop = prev_lane, tail_prev = prev;
        while (__any_sync(0xffffffffu, hop >= 0)) {
            const int src = hop >= 0 ? hop : lane;
            const int f_src = __shfl_sync(0xffffffffu, my_full, src);
            const int p_src = __shfl_sync(0xffffffffu, prev_lane, src);
            const int pm_src = __shfl_sync(0xffffffffu, prev, src);
            if (hop >= 0) { lo_pos = f_src; hop = p_src; tail_prev = pm_src; }
        }
        bool emit = false;
        ItemRec it;
        it.key = 0; it.t = 0; it.base = 0; it.woff = 0; it.sf = 0;
        if (valid && !has_next) {   // the chain ends here (closest to the stop): finish the way to its head and emit its item
            int cur = tail_prev;    // first start of the chain that is not in this warp (or -1)
            while (cur >= 0) {
                lo_pos = full_of(st[cur].sraw);
                cur = prev_in_memory(cur);
            }
            const int ext = t - lo_pos;
            it.orf = me.orf;
            if (amb) {  // does the walk [lo_pos - kc, t - 1] touch an ambiguous letter?  (bit 31 of orf selects the careful loop)
                const int strand = rec.sf & 1, Lc = clen[rec.contig];
                const int64_t w0c = cwoff[rec.contig];
                const int a = lo_pos - kc, b = t - 1;
                const int na = strand ? amb_count_fwd(pl, w0c, Lc - 1 - b, Lc - 1 - a) : amb_count_fwd(pl, w0c, a, b);
                if (na) it.orf |= 0x80000000u;
            }
            it.lo_pos = lo_pos;
            // position tile first (see WALK_TILE_WORDS), then section pair, then length class
            const long long tile_raw = (long long)(cwoff[rec.contig] + (t >> 5)) / pl.tile_words;
            const int tile = tile_raw < pl.n_tiles - 1 ? (int)tile_raw : pl.n_tiles - 1;
            it.key = (uint32_t)((tile * npair + (int)pair) * WALK_CLASSES + (WALK_CLASSES - 1 - min(ext >> 5, WALK_CLASSES - 1)));
            it.last = j;
            it.t = t;
            it.base = rec.base;
            it.woff = (uint32_t)((int64_t)(rec.sf & 1u) * pl.W + cwoff[rec.contig]);
            it.sf = rec.sf;
            emit = true;
        }
        const uint32_t em = __ballot_sync(0xffffffffu, emit);
        if (em) {
            const int cnt = __popc(em);
            unsigned long long base = 0;
            if (chunk) {
                if (cnt > chunk_left) {   // retire what is left of the old chunk, take a new one
                    if (lane < chunk_left && chunk_base + lane < cap_items) items[chunk_base + lane].key = 0xffffffffu;
                    if (lane == 0) chunk_base = atomicAdd(&counters[2], (unsigned long long)chunk);
                    chunk_base = __shfl_sync(0xffffffffu, chunk_base, 0);
                    chunk_left = chunk;
                }
                base = chunk_base;
                chunk_base += cnt;
                chunk_left -= cnt;
            } else {
                if (lane == 0) base = atomicAdd(&counters[2], (unsigned long long)cnt);
                base = __shfl_sync(0xffffffffu, base, 0);
            }
            const unsigned long long at = base + __popc(em & ((1u << lane) - 1u));
            const bool ok = emit && at < cap_items;
            if (ok) items[at] = it;
            warp_key_reserve(hist, it.key, ok);
        }
    }
    for (int i = lane; i < chunk_left; i += 32)   // unused tail of the warp's last chunk
        if (chunk_base + i < cap_items) items[chunk_base + i].key = 0xffffffffu;
}

// one block per model: item offsets per key (exclusive scan of the histogram, which is turned into the scatter
// cursor); work units = chunks of chunk_items (WALK_CHUNK or WALK2_CHUNK) items of one section pair, with their prefix over pairs
__global__ void __launch_bounds__(1024) k_item_plan(uint32_t *const *hist_m, uint32_t *const *item_off_m, uint32_t *const *unit_off_m,
                                                    const DevModel *__restrict__ models, uint32_t *unit_totals, unsigned long long *counters_m, int n_tiles,
                                                    uint32_t chunk_items) {
    __shared__ uint32_t carry_i;
    const int mi = blockIdx.x;
    const int npair = models[mi].npair * n_tiles, nkeys = npair * WALK_CLASSES;   // (tile, pair) combinations
    uint32_t *hist = hist_m[mi], *ioff = item_off_m[mi], *uoff = unit_off_m[mi];
    if (threadIdx.x == 0) carry_i = 0;
    __syncthreads();
    for (int b = 0; b < nkeys; b += 1024) {
        const int i = b + threadIdx.x;
        const uint32_t cnt = i < nkeys ? hist[i] : 0u;
        uint32_t ti;
        const uint32_t ei = block_exclusive_scan(cnt, &ti);
        if (i < nkeys) {
            ioff[i] = carry_i + ei;
            hist[i] = carry_i + ei;  // scatter cursor
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_i += ti;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ioff[nkeys] = carry_i;
        counters_m[(size_t)mi * 8 + 5] = carry_i;   // walk items (the slot counter 2 includes unfilled slots)
        uint32_t acc = 0;
        for (int p = 0; p < npair; p++) {
            uoff[p] = acc;
            const uint32_t cnt = ioff[(p + 1) * WALK_CLASSES] - ioff[p * WALK_CLASSES];
            acc += (cnt + chunk_items - 1) / chunk_items;
        }
        uoff[npair] = acc;
        unit_totals[mi] = acc;
    }
}

__global__ void __launch_bounds__(256) k_item_scatter(const ItemRec *const *items_m, ItemRec *const *sorted_m, uint32_t *const *cursor_m,
                                                      const unsigned long long *counters_m, uint32_t cap_items) {
    const int mi = blockIdx.y;
    unsigned long long n = counters_m[(size_t)mi * 8 + 2];
    if (n > cap_items) n = cap_items;
    const ItemRec *items = items_m[mi];
    ItemRec *sorted = sorted_m[mi];
    uint32_t *cursor = cursor_m[mi];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31LL; k0 < (int64_t)n; k0 += stride) {
        const int64_t k = k0 + lane_id();
        const bool has = k < (int64_t)n;
        ItemRec it;
        it.key = 0;
        if (has) it = items[k];
        const bool real = has && it.key != 0xffffffffu;   // slots reserved but not filled by k_start_link are skipped
        const uint32_t at = warp_key_reserve(cursor, it.key, real);
        if (real) sorted[at] = it;
    }
}

// ------------------------------------------------------------------------------------------------
// the walk: thread per item, the section pair's table in shared memory
// ------------------------------------------------------------------------------------------------
// A work unit is a chunk of up to WALK_CHUNK items of one (model, section pair), ordered long to short.  Persistent
// CTAs take units in order, so a CTA reloads its ~100 KB table only when the pair changes; inside a unit the CTA's warps
// take 32 items at a time from a shared counter and never wait for each other.  Each thread walks its item from the
// letter before the stop codon down to lo_pos, keeping the (k+1)-letter window as two shift registers, adds one Quad
// per position, and stores the running sums whenever it passes the first full-order position of a start codon of its
// pair (the starts of a pair are chained, the next one is prefetched).  No shuffles.
struct __align__(16) Acc4 {
    double cB, oB, cA, oA;
};

template <bool SMEM_TABLE>
__global__ void __launch_bounds__(WALK_UNIT, 2) k_walk(Planes pl, const int64_t *__restrict__ cwoff, const DevModel *__restrict__ models,
                                                       int n_models, const OrfRec *const *orfs_m, const StartRec *const *starts_m,
                                                       const ItemRec *const *items_m, const uint32_t *const *item_off_m,
                                                       const uint32_t *const *unit_off_m, const uint32_t *__restrict__ unit_totals,
                                                       Acc4 *const *sums_m, unsigned int *next_unit) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *tbl = reinterpret_cast<double2 *>(smem_raw);
    __shared__ int s_model, s_pair;
    __shared__ uint32_t s_first, s_end;
    __shared__ unsigned int s_next;
    const int lane = lane_id();
    int loaded_model = -1, loaded_pair = -1;
    while (true) {
        if (threadIdx.x == 0) {
            unsigned int u = atomicAdd(next_unit, 1u);
            int m = 0;
            while (m < n_models && u >= unit_totals[m]) { u -= unit_totals[m]; m++; }
            s_model = m;
            if (m < n_models) {
                const uint32_t *uoff = unit_off_m[m];
                int lo = 0, hi = models[m].npair * pl.n_tiles;  // last (tile, pair) with uoff[.] <= u (uoff has one more entry)
                while (lo + 1 < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (uoff[mid] <= u) lo = mid; else hi = mid;
                }
                s_pair = lo;
                const uint32_t *ioff = item_off_m[m];
                const uint32_t first = ioff[lo * WALK_CLASSES] + (u - uoff[lo]) * WALK_CHUNK;
                s_first = first;
                s_end = min(first + (uint32_t)WALK_CHUNK, ioff[(lo + 1) * WALK_CLASSES]);
                s_next = first;
            }
        }
        __syncthreads();
        const int mi = s_model;
        if (mi >= n_models) break;
        const DevModel &dm = models[mi];
        const int pair = s_pair % dm.npair;   // s_pair counts (tile, pair) combinations, tile-major
        const int kc = dm.kc, kn = dm.kn, nfull = dm.nfull;
        const double2 *gq = dm.wtab + (size_t)pair * dm.wtab_n16;
        if (SMEM_TABLE && (mi != loaded_model || pair != loaded_pair)) {
            const int4 *src = reinterpret_cast<const int4 *>(gq);
            int4 *dst = reinterpret_cast<int4 *>(tbl);
            const int n16 = dm.wtab_n16;
            for (int i = threadIdx.x; i < n16; i += WALK_UNIT) dst[i] = __ldg(src + i);
            loaded_model = mi;
            loaded_pair = pair;
            __syncthreads();
        }
        // coding part: {ln P_cod^B, ln P_cod^A} per (phase, k-mer), the two replicas of an entry side by side so that even and
        // odd lanes use disjoint bank groups (even / odd 16-byte slots); non-coding part: {ln P_non^B, ln P_non^A} per (kn+1)-mer, eight replicas so
        // that the eight lanes of a quarter-warp never collide
        const double2 *tc = SMEM_TABLE ? tbl : gq;
        const double2 *tn = tc + (size_t)3 * nfull * 2;
        const double2 *tc0 = tc + (lane & 1), *tn8 = tn + (lane & 7);   // my replicas
        const int pstep = 2 * nfull;                                     // slots per codon position
        // constant terms of windows with an ambiguous letter, per domain of this pair
        const double2 amb_c = make_double2(dm.pair_amb[pair * 4 + 0], dm.pair_amb[pair * 4 + 1]);
        const double2 amb_n = make_double2(dm.pair_amb[pair * 4 + 2], dm.pair_amb[pair * 4 + 3]);
        const int nsh = kc - kn;
        const uint32_t wmask = (1u << (kc + 1)) - 1u;
        const uint32_t end = s_end;
        const ItemRec *items = items_m[mi];
        while (true) {
            uint32_t w_first = 0;
            if (lane == 0) w_first = atomicAdd(&s_next, 32u);
            w_first = __shfl_sync(0xffffffffu, w_first, 0);
            if (w_first >= end) break;
            if (w_first + lane < end) {
                const ItemRec it = items[w_first + lane];
                const bool careful = (it.orf >> 31) != 0;   // the walk touches an ambiguous letter
                const OrfRec rec = orfs_m[mi][it.orf & 0x7fffffffu];
                const int t = rec.t, strand = rec.sf & 1;
                const bool from_first = (rec.sf & 0x200u) != 0;
                const uint32_t *lo = pl.lo + (int64_t)strand * pl.W + cwoff[rec.contig];
                const uint32_t *hi = pl.hi + (int64_t)strand * pl.W + cwoff[rec.contig];
                const StartRec *st = starts_m[mi] + rec.base;
                Acc4 *sums = sums_m[mi] + rec.base;
                // the starts of my pair, from the one closest to the stop outwards, one record prefetched ahead
                int j = it.last;
                StartRec cur = st[j];
                auto full_of = [&](const StartRec &sr) {
                    return (int)(sr.sraw & 0x7fffffffu) + (((sr.sraw >> 31) && from_first) ? 0 : 3) + kc;
                };
                int target = full_of(cur);
                StartRec nxt = cur;
                if (cur.link >= 0) nxt = st[cur.link];
                Acc4 acc = {0, 0, 0, 0};
                int i = t - 1;
                const int ph = (((-1 - kc) % 3) + 3) % 3;   // codon position of the window's first letter at i = t-1
                const double2 *tcp = tc0 + ph * pstep;       // the coding table of that codon position (my replica)
                // the (kc+1)-letter window ending at i starts at bit `sh` of plane word `w`; the word pair (w, w+1) sits in registers
                int w = (i - kc) >> 5, sh = (i - kc) & 31;
                // plane words come in 16-byte groups (4 words = 128 letters) held in registers: one sector request per
                // group instead of one per word (the L1 is too small for 512 concurrent item streams per SM)
                const uint4 *lo4 = reinterpret_cast<const uint4 *>(lo), *hi4 = reinterpret_cast<const uint4 *>(hi);
                uint4 gl = lo4[w >> 2], gh = hi4[w >> 2];
                auto pick = [](const uint4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); };
                uint32_t l0 = pick(gl, w & 3), h0 = pick(gh, w & 3);
                uint32_t l1 = (w & 3) == 3 ? lo[w + 1] : pick(gl, (w & 3) + 1), h1 = (w & 3) == 3 ? hi[w + 1] : pick(gh, (w & 3) + 1);
                const int lo_pos = it.lo_pos;
                while (i >= lo_pos) {
                    // a run ends at the next target, at lo_pos, or where the window start leaves the current plane word
                    const int run = min(i - max(target, lo_pos) + 1, sh + 1);
                    if (!careful) {
                        // software pipeline: the two lookups of position r+1 are in flight while position r is accumulated
                        auto lookup = [&](double2 &c2, double2 &n2) {
                            const uint32_t lb = __funnelshift_r(l0, l1, sh) & wmask;
                            const uint32_t hb = __funnelshift_r(h0, h1, sh) & wmask;
                            const uint32_t idn = ((hb >> nsh) << (kn + 1)) | (lb >> nsh);
                            c2 = tcp[((hb << (kc + 1)) | lb) << 1];
                            n2 = tn8[idn << 3];
                            sh--;
                            tcp = tcp == tc0 ? tc0 + 2 * pstep : tcp - pstep;   // codon position of the window's first letter steps 0 -> 2 -> 1 -> 0
                        };
                        double2 c2, n2;
                        lookup(c2, n2);
                        for (int r = 1; r < run; r++) {
                            double2 c3, n3;
                            lookup(c3, n3);
                            acc.cB += c2.x; acc.cA += c2.y; acc.oB += n2.x; acc.oA += n2.y;   // oB/oA hold the non-coding sums here
                            c2 = c3; n2 = n3;
                        }
                        acc.cB += c2.x; acc.cA += c2.y; acc.oB += n2.x; acc.oA += n2.y;
                    } else {
                        // a window with an ambiguous letter contributes ln COD_P_N to the coding side and ln NON_P_N to the
                        // non-coding side of each domain (so ln(COD_P_N / NON_P_N) to the log-odds)
                        const uint32_t *nmp = pl.nm + (int64_t)strand * pl.W + cwoff[rec.contig];
                        const uint32_t a0 = nmp[w], a1 = nmp[w + 1];
                        for (int r = 0; r < run; r++) {
                            const uint32_t lb = __funnelshift_r(l0, l1, sh) & wmask;
                            const uint32_t hb = __funnelshift_r(h0, h1, sh) & wmask;
                            const uint32_t idn = ((hb >> nsh) << (kn + 1)) | (lb >> nsh);
                            double2 c2 = tcp[((hb << (kc + 1)) | lb) << 1];
                            double2 n2 = tn8[idn << 3];
                            if (__funnelshift_r(a0, a1, sh) & wmask) { c2 = amb_c; n2 = amb_n; }
                            acc.cB += c2.x; acc.cA += c2.y; acc.oB += n2.x; acc.oA += n2.y;
                            sh--;
                            tcp = tcp == tc0 ? tc0 + 2 * pstep : tcp - pstep;
                        }
                    }
                    i -= run;
                    if (sh < 0 && w > 0) {
                        w--;
                        sh = 31;
                        l1 = l0; h1 = h0;
                        if ((w & 3) == 3) { gl = lo4[w >> 2]; gh = hi4[w >> 2]; }
                        l0 = pick(gl, w & 3); h0 = pick(gh, w & 3);
                    }
                    while (i + 1 == target) {   // a virtual and a real start may share their first full-order position
                        Acc4 out;
                        out.cB = acc.cB; out.oB = acc.cB - acc.oB; out.cA = acc.cA; out.oA = acc.cA - acc.oA;
                        sums[j] = out;
                        j = cur.link;
                        if (j < 0) { target = -1; break; }
                        cur = nxt;
                        target = full_of(cur);
                        if (cur.link >= 0) nxt = st[cur.link];
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// the pair walk: two positions per table lookup
// ------------------------------------------------------------------------------------------------
// Same items, same sums as k_walk, but the shared-memory tables are indexed by the (kc+2)-mer and hold the SUM of the two
// positions that window covers ({B, A} coding log-likelihoods of positions i-1 and i), the non-coding table likewise by
// the (kn+2)-mer: one coding and one non-coding lookup per TWO positions.  That halves the shared-memory wavefronts,
// which bound k_walk (profiles/r1j_k_walk_1Gbp.txt), and the instruction count with them.  Order 4/2: 3 x 4096 x 16 B =
// 192 KB coding (one copy) + 256 x 16 B x 8 replicas = 32 KB non-coding: one CTA of 512 threads owns the SM.
//
// A walk therefore advances in steps of two from the letter before the stop codon.  The first full-order position F of a
// start lies t - F positions below the stop; when that count is odd the walk delivers the sum down to F + 1 and k_final
// adds position F itself from the single-position table (phase 0 by construction: F - kc is the first summed letter).
// The loop is flat — one pair step per iteration for every lane, plane-word refills and sum deliveries as short
// predicated blocks inside it — so the lanes of a warp stay together until their items (sorted by length) end.
// Items whose range touches an ambiguous letter take walk_careful() (single positions, tables behind L1/L2).
struct WalkItemCtx {
    const uint32_t *lo, *hi, *nm;
    const StartRec *st;
    Acc4 *sums;
    int t, kc, kn, lo2, j;
    bool from_first;
};

__device__ __forceinline__ int pair_target(uint32_t sraw, bool from_first, int kc, int t) {
    const int F = (int)(sraw & 0x7fffffffu) + (((sraw >> 31) && from_first) ? 0 : 3) + kc;
    return F + ((t - F) & 1);
}

__device__ __noinline__ void walk_careful(const WalkItemCtx &cx, const double2 *gq, int nfull, double2 amb_c, double2 amb_n, int lane) {
    // single positions from the global image of k_walk's table (coding part 2 replicas, non-coding part 8)
    const int kc = cx.kc, kn = cx.kn, t = cx.t;
    const double2 *tc0 = gq + (lane & 1), *tn8 = gq + (size_t)3 * nfull * 2 + (lane & 7);
    const int pstep = 2 * nfull, nsh = kc - kn;
    const uint32_t wmask = (1u << (kc + 1)) - 1u;
    int j = cx.j;
    StartRec cur = cx.st[j];
    int target = pair_target(cur.sraw, cx.from_first, kc, t);
    Acc4 acc = {0, 0, 0, 0};
    int ph = (((-1 - kc) % 3) + 3) % 3;
    for (int i = t - 1;; i--) {
        while (i + 1 == target) {
            Acc4 out;
            out.cB = acc.cB; out.oB = acc.cB - acc.oB; out.cA = acc.cA; out.oA = acc.cA - acc.oA;
            cx.sums[j] = out;
            j = cur.link;
            if (j < 0) return;
            cur = cx.st[j];
            target = pair_target(cur.sraw, cx.from_first, kc, t);
        }
        uint32_t hb, lb, ab, dummy;
        window_bits(cx.lo, cx.hi, i - kc, kc + 1, hb, lb);
        window_bits(cx.nm, cx.nm, i - kc, kc + 1, ab, dummy);
        hb &= wmask; lb &= wmask;
        double2 c2 = amb_c, n2 = amb_n;
        if (!ab) {
            c2 = __ldg(tc0 + ph * pstep + (((hb << (kc + 1)) | lb) << 1));
            n2 = __ldg(tn8 + ((((hb >> nsh) << (kn + 1)) | (lb >> nsh)) << 3));
        }
        acc.cB += c2.x; acc.cA += c2.y; acc.oB += n2.x; acc.oA += n2.y;
        ph = ph == 0 ? 2 : ph - 1;
    }
}

__device__ __forceinline__ double2 lds_d2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// KC, KN >= 0: the orders are compile-time constants (all shifts and masks of a step fold); KC < 0: read from the model.
template <int KC, int KN>
__global__ void __launch_bounds__(WALK2_THREADS, 1) k_walk_pair(Planes pl, const DevModel *__restrict__ models, int n_models,
                                                                const StartRec *const *starts_m, const ItemRec *const *items_m,
                                                                const uint32_t *const *item_off_m, const uint32_t *const *unit_off_m,
                                                                const uint32_t *__restrict__ unit_totals, Acc4 *const *sums_m,
                                                                unsigned int *next_unit) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *tbl = reinterpret_cast<double2 *>(smem_raw);
    __shared__ int s_model, s_pair;
    __shared__ uint32_t s_end;
    __shared__ unsigned int s_next;
    __shared__ __align__(8) unsigned long long s_mbar;   // completion of the table's bulk copies (transaction bytes)
    const int lane = lane_id();
    int loaded_model = -1, loaded_pair = -1;
    const uint32_t tbl_addr = (uint32_t)__cvta_generic_to_shared(tbl);
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    uint32_t tbl_phase = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    while (true) {
        if (threadIdx.x == 0) {
            unsigned int u = atomicAdd(next_unit, 1u);
            int m = 0;
            while (m < n_models && u >= unit_totals[m]) { u -= unit_totals[m]; m++; }
            s_model = m;
            if (m < n_models) {
                const uint32_t *uoff = unit_off_m[m];
                int lo = 0, hi = models[m].npair * pl.n_tiles;  // last (tile, pair) with uoff[.] <= u (uoff has one more entry)
                while (lo + 1 < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (uoff[mid] <= u) lo = mid; else hi = mid;
                }
                s_pair = lo;
                const uint32_t *ioff = item_off_m[m];
                const uint32_t first = ioff[lo * WALK_CLASSES] + (u - uoff[lo]) * WALK2_CHUNK;
                s_end = min(first + (uint32_t)WALK2_CHUNK, ioff[(lo + 1) * WALK_CLASSES]);
                s_next = first;
            }
        }
        __syncthreads();
        const int mi = s_model;
        if (mi >= n_models) break;
        const DevModel &dm = models[mi];
        const int pair = s_pair % dm.npair;   // s_pair counts (tile, pair) combinations, tile-major
        const int kc = KC >= 0 ? KC : dm.kc, kn = KC >= 0 ? KN : dm.kn;
        if (mi != loaded_model || pair != loaded_pair) {
            // The unit's tables (up to 224 KB, the SM's whole shared memory: nothing to double-buffer into) come in by bulk
            // copies of the async proxy (cp.async.bulk, 32 KB each) that one thread issues; everybody waits on the mbarrier
            // they complete.  Nobody still reads the previous tables: all threads passed the barrier that ended the last unit.
            const uint32_t bytes = (uint32_t)dm.wtab2_n16 * 16u;
            if (threadIdx.x == 0) {
                const char *src = reinterpret_cast<const char *>(dm.wtab2 + (size_t)pair * dm.wtab2_n16);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
                for (uint32_t at = 0; at < bytes; at += 32768u) {
                    const uint32_t n = min(32768u, bytes - at);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tbl_addr + at),
                                 "l"(src + at), "r"(n), "r"(mbar)
                                 : "memory");
                }
            }
            uint32_t done = 0;
            for (uint32_t spins = 0; !done; spins++) {
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done)
                             : "r"(mbar), "r"(tbl_phase)
                             : "memory");
                if (spins > (1u << 26)) __trap();   // a copy that never completes must not hang the device
            }
            tbl_phase ^= 1u;
            loaded_model = mi;
            loaded_pair = pair;
        }
        // shared-memory byte addresses: coding entry of (codon position p, window index x) at tc + p * pbytes + (x << (rcs + 4)),
        // my replica folded into tc; non-coding entry of index y at tn + (y << 7) (8 replicas of 16 bytes)
        const int rcs = KC == 4 ? 0 : 31 - __clz(dm.rc2);        // log2 of the coding replicas (order 4: one copy fills the SM)
        const uint32_t pbytes = (uint32_t)(1u << (2 * (kc + 2))) << (rcs + 4);
        const uint32_t tc = tbl_addr + ((uint32_t)(lane & ((1 << rcs) - 1)) << 4);
        const uint32_t tn = tbl_addr + 3 * pbytes + ((uint32_t)(lane & 7) << 4);
        const int nsh = kc - kn;
        const uint32_t wmask = (1u << (kc + 2)) - 1u;
        const int ph0 = (((-2 - kc) % 3) + 3) % 3;               // codon position of the first letter of the topmost pair window
        const uint32_t end = s_end;
        const ItemRec *items = items_m[mi];
        const StartRec *starts = starts_m[mi];
        Acc4 *sums_all = sums_m[mi];
        while (true) {
            uint32_t w_first = 0;
            if (lane == 0) w_first = atomicAdd(&s_next, 32u);
            w_first = __shfl_sync(0xffffffffu, w_first, 0);
            if (w_first >= end) break;
            if (w_first + lane >= end) continue;
            const int4 ia = __ldg(reinterpret_cast<const int4 *>(items + w_first + lane));
            const int4 ib = __ldg(reinterpret_cast<const int4 *>(items + w_first + lane) + 1);
            const int t = ib.x;
            const bool from_first = ((uint32_t)ib.w & 0x200u) != 0;
            const uint32_t *lo = pl.lo + (uint32_t)ib.z, *hi = pl.hi + (uint32_t)ib.z;
            const StartRec *st = starts + (uint32_t)ib.y;
            Acc4 *sums = sums_all + (uint32_t)ib.y;
            int j = ia.w;
            const int lo2 = ia.y + ((t - ia.y) & 1);
            if ((uint32_t)ia.x >> 31) {   // the walk touches an ambiguous letter
                WalkItemCtx cx;
                cx.lo = lo; cx.hi = hi; cx.nm = pl.nm + (uint32_t)ib.z; cx.st = st; cx.sums = sums;
                cx.t = t; cx.kc = kc; cx.kn = kn; cx.lo2 = lo2; cx.j = j; cx.from_first = from_first;
                walk_careful(cx, dm.wtab + (size_t)pair * dm.wtab_n16, dm.nfull,
                             make_double2(dm.pair_amb[pair * 4 + 0], dm.pair_amb[pair * 4 + 1]),
                             make_double2(dm.pair_amb[pair * 4 + 2], dm.pair_amb[pair * 4 + 3]), lane);
                continue;
            }
            // the starts of my pair, from the one closest to the stop outwards; the next two records of the chain are kept
            // loaded (starts of a gene often sit a few codons apart: one record ahead is not enough to hide the load)
            StartRec cur = st[j];
            int target = pair_target(cur.sraw, from_first, kc, t);
            StartRec nxt = cur, nx2 = cur;
            if (cur.link >= 0) {
                nxt = st[cur.link];
                nx2 = nxt;
            }
            bool nx2_pending = cur.link >= 0;           // nx2 still has to be loaded (needs nxt.link, which is in flight)
            int steps = (t - lo2) >> 1;                 // pair steps of this walk
            // Rounds of 16 pair steps, the same for every lane of the warp.  The window of step r of a round starts 2r letters
            // below the round's first window; a 64-letter register pair per plane, whose bit 0 is the letter 30 below that first
            // window, holds everything the round reads, so step r extracts its (kc+2)-mer with the CONSTANT shift 30 - 2r.
            // Between rounds the register pair moves down by one 32-letter word — one new (unaligned: same bit offset `off`
            // for the whole walk) word per plane, loaded a round ahead.  No lane-dependent refill inside the steps.
            const int base = t - 2 - kc - 30;           // letter index of bit 0 in the first round (may be negative: unused letters)
            int k = base >> 5;                          // word holding that letter (arithmetic shift: floor)
            const int off = base & 31;
            // plane words travel as 16-byte groups (4 words = 128 letters, one sector request per group and plane instead of
            // one per word: the L1 left beside 224 KB of tables does not keep 2 x 512 word streams, and the outstanding
            // sector misses of an SM are what bounds this kernel once the table lookups are halved)
            const uint4 *lo4 = reinterpret_cast<const uint4 *>(lo), *hi4 = reinterpret_cast<const uint4 *>(hi);
            auto pick = [](const uint4 &v, int q) { return q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w)); };
            const int k0 = k < 0 ? 0 : k;               // words below the contig hold no letter a window reads
            uint4 gl = __ldg(lo4 + (k0 >> 2)), gh = __ldg(hi4 + (k0 >> 2));
            uint32_t wl = pick(gl, k0 & 3), wh = pick(gh, k0 & 3);                       // W[k]
            uint32_t ul, uh, vl, vh;                                                     // W[k+1], W[k+2]
            {
                const int k1 = k + 1 < 0 ? 0 : k + 1, k2 = k + 2 < 0 ? 0 : k + 2;
                uint4 g2l = gl, g2h = gh;
                if ((k2 >> 2) != (k0 >> 2)) { g2l = __ldg(lo4 + (k2 >> 2)); g2h = __ldg(hi4 + (k2 >> 2)); }
                ul = (k1 >> 2) == (k0 >> 2) ? pick(gl, k1 & 3) : pick(g2l, k1 & 3);
                uh = (k1 >> 2) == (k0 >> 2) ? pick(gh, k1 & 3) : pick(g2h, k1 & 3);
                vl = (k2 >> 2) == (k0 >> 2) ? pick(gl, k2 & 3) : pick(g2l, k2 & 3);
                vh = (k2 >> 2) == (k0 >> 2) ? pick(gh, k2 & 3) : pick(g2h, k2 & 3);
            }
            uint32_t low_l = __funnelshift_r(wl, ul, off), low_h = __funnelshift_r(wh, uh, off);
            uint32_t high_l = __funnelshift_r(ul, vl, off), high_h = __funnelshift_r(uh, vh, off);
            // coding tables of the three codon positions in the order this round's steps use them: step r reads tp[r % 3];
            // a round has 16 steps, so the triple rotates by one between rounds
            uint32_t tp0 = tc + (uint32_t)ph0 * pbytes, tp1 = tc + (uint32_t)((ph0 + 1) % 3) * pbytes, tp2 = tc + (uint32_t)((ph0 + 2) % 3) * pbytes;
            double cB = 0, cA = 0, nB = 0, nA = 0;
            int i = t - 1;                              // topmost position not yet summed
            auto deliver = [&]() {
                do {                                    // a virtual and a real start may share their first full-order position
                    Acc4 out;
                    out.cB = cB; out.oB = cB - nB; out.cA = cA; out.oA = cA - nA;
                    sums[j] = out;
                    j = cur.link;
                    if (j < 0) { target = -1 << 30; break; }
                    if (nx2_pending) { nx2 = nxt; if (nxt.link >= 0) nx2 = st[nxt.link]; nx2_pending = false; }
                    cur = nxt;
                    nxt = nx2;
                    target = pair_target(cur.sraw, from_first, kc, t);
                    if (nxt.link >= 0) nx2 = st[nxt.link];
                } while (i + 1 == target);
            };
            if (i + 1 == target) deliver();             // starts with an empty (or one-position) full-order region
            while (steps > 0) {
                const int n_this = steps < 16 ? steps : 16;
                steps -= n_this;
                // next round's word: when it opens a new group, the group is on its way while this round runs
                k--;
                if (k >= 0 && (k & 3) == 3) { gl = __ldg(lo4 + (k >> 2)); gh = __ldg(hi4 + (k >> 2)); }
                if (nx2_pending) { nx2 = nxt; if (nxt.link >= 0) nx2 = st[nxt.link]; nx2_pending = false; }
                double2 c2, n2;
                auto lookup = [&](int shift, uint32_t tp, double2 &c, double2 &n) {
                    const uint32_t lb = __funnelshift_r(low_l, high_l, shift) & wmask;
                    const uint32_t hb = __funnelshift_r(low_h, high_h, shift) & wmask;
                    c = lds_d2(tp + (((hb << (kc + 2)) | lb) << (rcs + 4)));
                    n = lds_d2(tn + ((((hb >> nsh) << (kn + 2)) | (lb >> nsh)) << 7));
                };
                lookup(30, tp0, c2, n2);
#pragma unroll
                for (int r = 0; r < 16; r++) {
                    if (r < n_this) {
                        double2 c3 = c2, n3 = n2;
                        if (r + 1 < 16 && r + 1 < n_this)   // in flight while this pair is added
                            lookup(28 - 2 * r, (r + 1) % 3 == 0 ? tp0 : ((r + 1) % 3 == 1 ? tp1 : tp2), c3, n3);
                        cB += c2.x; cA += c2.y; nB += n2.x; nA += n2.y;
                        i -= 2;
                        if (i + 1 == target) deliver();
                        c2 = c3; n2 = n3;
                    }
                }
                const uint32_t nl = pick(gl, k & 3), nh = pick(gh, k & 3);   // k < 0: letters below the contig, never read
                high_l = low_l; high_h = low_h;
                low_l = __funnelshift_r(nl, wl, off); low_h = __funnelshift_r(nh, wh, off);
                wl = nl; wh = nh;
                { const uint32_t x = tp0; tp0 = tp1; tp1 = tp2; tp2 = x; }   // 16 steps further: codon position + 16 = + 1
            }
        }
        __syncthreads();   // nobody still reads s_next / s_end of this unit when thread 0 fetches the next one
    }
}

// ------------------------------------------------------------------------------------------------
// per-start totals: order ramp, domain choice, duration / start / stop terms, pruning, bucket histogram
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) k_final(Planes pl, const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                               const DevModel *__restrict__ models, const OrfRec *const *orfs_m,
                                               const StartRec *const *starts_m, const Acc4 *const *sums_m,
                                               const unsigned long long *counters_m, Cand *const *surv_m, uint32_t *const *bcount_m,
                                               uint32_t *const *cover_m, unsigned long long *surv_counters, uint32_t cap_starts,
                                               uint32_t cap_surv, int pair_walk) {
    const int mi = blockIdx.y;
    const DevModel &dm = models[mi];
    const int lane = lane_id();
    const unsigned long long n_starts = counters_m[(size_t)mi * 8 + 1];
    if (n_starts > cap_starts) return;   // overflowed, partly written start array: nothing to score, the host grows it and decodes again
    const OrfRec *orfs = orfs_m[mi];
    const StartRec *starts = starts_m[mi];
    const Acc4 *sums = sums_m[mi];
    Cand *surv = surv_m[mi];
    uint32_t *bcount = bcount_m[mi];
    uint32_t *cover = cover_m[mi];
    const int kc = dm.kc, nramp = dm.nramp;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // Survivor slots are reserved per warp in chunks of SURV_CHUNK (one atomic per ~30 rounds instead of one per round: the
    // single hot counter otherwise decides the speed of this latency-bound kernel); slots a warp does not fill
    // carry contig = -1 and are skipped by k_scatter.  Small inputs reserve exactly (the waste would not fit their buffers).
    const int chunk = n_starts >= (1ull << 22) ? SURV_CHUNK : 0;
    unsigned long long chunk_base = 0;
    int chunk_left = 0;
    // whole warps iterate together (ballots below); the loop bound is rounded up to a multiple of 32
    for (int64_t k0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31LL; k0 < (int64_t)n_starts; k0 += stride) {
        const int64_t k = k0 + lane;
        const bool has = k < (int64_t)n_starts;
        bool keep = false;
        Cand cd;
        cd.right = 0; cd.contig = 0;
        uint32_t orf_id = 0xffffffffu;
        int s = 0, strand = 0;
        bool virt = false;
        int64_t w0 = 0;
        uint32_t orf_base = 0, orf_n = 0;
        if (has) {
            const StartRec sr = starts[k];
            const OrfRec rec = orfs[sr.orf];
            orf_id = sr.orf; orf_base = rec.base; orf_n = rec.n;
            const int c = rec.contig, t = rec.t;
            strand = rec.sf & 1;
            const bool pstop = (rec.sf & 0x100u) != 0;
            const int L = clen[c];
            w0 = cwoff[c];
            const uint32_t *lo = pl.lo + (int64_t)strand * pl.W + w0;
            const uint32_t *hi = pl.hi + (int64_t)strand * pl.W + w0;
            virt = (sr.sraw >> 31) != 0;
            s = (int)(sr.sraw & 0x7fffffffu);
            const int len = t + 3 - s;
            const int gc = (int)(sr.pg >> 16), pair = (int)(sr.pg & 0xffffu);
            const int secB = dm.map[0][gc], secA = dm.map[1][gc];
            const int first = s + ((virt && (rec.sf & 0x200u)) ? 0 : 3);
            const int fullstart = first + kc;
            Acc4 acc = {0, 0, 0, 0};
            if (fullstart <= t - 1) acc = sums[k];
            // the pair walk advances two positions at a time from the stop: with an odd number of full-order positions it
            // stopped one above `fullstart`, and that position is added below from the single-position table
            const bool odd_one = pair_walk && fullstart <= t - 1 && ((t - fullstart) & 1);
            // Early pruning: with the largest possible ramp / start / stop terms of either domain the total stays <= 0 for most
            // starts of spurious ORFs; those never touch the planes or the ramp tables.  (Exact otherwise: only candidates
            // that the full computation would prune as well are dropped, with a 1e-6 guard band.)
            const double dl = (double)len;
            // 2 ln(len) for the bound only: the single-precision logarithm plus a margin far above its error (conversion 6e-8,
            // __logf 3 ulp ~ 6e-6 at ln 1e7) — the double-precision log of all 145 M starts was 12 % of this kernel's
            // instructions; survivors of the bound take the exact one below
            const double lgt = 2.0 * ((double)__logf((float)len) + 1e-4);
            bool maybe = false;
            {
                const double *ub = dm.pair_ub + pair * 8;
                const double ub1B = odd_one ? ub[6] : 0.0, ub1A = odd_one ? ub[7] : 0.0;
                if (secB >= 0) {
                    const SectionScalars *sb = dm.sc + secB;
                    maybe |= acc.oB + ub1B + ub[0] + lgt + sb->tr + dl * sb->dur_slope + sb->dur_const + (virt ? 0.0 : ub[1]) + (pstop ? 0.0 : ub[2]) > -1e-6;
                }
                if (secA >= 0) {
                    const SectionScalars *sa = dm.sc + secA;
                    maybe |= acc.oA + ub1A + ub[3] + lgt + sa->tr + dl * sa->dur_slope + sa->dur_const + (virt ? 0.0 : ub[4]) + (pstop ? 0.0 : ub[5]) > -1e-6;
                }
            }
            if (maybe) {
            const uint32_t *nmp = pl.dflags[1] ? pl.nm + (int64_t)strand * pl.W + w0 : nullptr;
            // order ramp: orders 0..kc-1 at positions first .. first+kc-1 (window starts on a codon boundary)
            const Quad *rq = dm.rquad + (size_t)pair * dm.kr * nramp;
            // The ramp shares one ambiguity test: an ambiguous letter anywhere in the first full-order window (the kc+1
            // letters the sum starts with) silences all kc reduced-order positions (probed against the reference).
            bool ramp_bad = false;
            if (nmp) {
                uint32_t nb_, dummy;
                window_bits(nmp, nmp, first, min(kc + 1, t + 3 - first), nb_, dummy);
                ramp_bad = nb_ != 0;
            }
            if (odd_one) {   // position `fullstart`: window of kc+1 letters starting at `first`, codon position 0
                uint32_t hb, lb;
                window_bits(lo, hi, first, kc + 1, hb, lb);
                if (ramp_bad) {   // the same window as the ramp's ambiguity test
                    const double *pa = dm.pair_amb + pair * 4;
                    acc.cB += pa[0]; acc.oB += pa[0] - pa[2]; acc.cA += pa[1]; acc.oA += pa[1] - pa[3];
                } else {
                    const Quad *q = dm.quad + (size_t)pair * 3 * dm.nfull + ((hb << (kc + 1)) | lb);
                    const double2 x = __ldg(reinterpret_cast<const double2 *>(q));
                    const double2 y = __ldg(reinterpret_cast<const double2 *>(q) + 1);
                    acc.cB += x.x; acc.oB += x.y; acc.cA += y.x; acc.oA += y.y;
                }
            }
            uint32_t rhb = 0, rlb = 0;   // the kc letters at `first`, read once; order j uses the first j+1 of them
            if (kc > 0 && !ramp_bad) window_bits(lo, hi, first, kc, rhb, rlb);
            for (int j = 0; j < kc && first + j <= t - 1 && !ramp_bad; j++) {
                const uint32_t jm = (2u << j) - 1u, hb = rhb & jm, lb = rlb & jm;
                const Quad *q = rq + (size_t)j * nramp + ((hb << (j + 1)) | lb);
                const double2 x = __ldg(reinterpret_cast<const double2 *>(q));
                const double2 y = __ldg(reinterpret_cast<const double2 *>(q) + 1);
                acc.cB += x.x; acc.oB += x.y; acc.cA += y.x; acc.oA += y.y;
            }
            if (secB >= 0 || secA >= 0) {
                // the domain with the larger coding log-likelihood plus its transition term ln(to_mgm * to_atypical) — the terms
                // are equal (ln 0.25) unless the parameter file sets $TO_ATYPICAL_*; ties -> bacteria
                const bool arc = secB < 0 || (secA >= 0 && acc.cA + dm.sc[secA].tr > acc.cB + dm.sc[secB].tr);
                const SectionScalars *sc = dm.sc + (arc ? secA : secB);
                double total = (arc ? acc.oA : acc.oB) + 2.0 * log(dl) + dl * sc->dur_slope + sc->dur_const + sc->tr;
                if (!virt) {
                    uint32_t hb, lb;
                    window_bits(lo, hi, s, 1, hb, lb);
                    total += sc->start_lo[hb ? (lb ? 2 : 1) : 0];  // A -> ATG, G -> GTG, T -> TTG
                }
                if (!pstop) {
                    uint32_t hb, lb;
                    window_bits(lo, hi, t, 3, hb, lb);
                    // TAA: A at +1,+2 ; TAG: G at +2 ; TGA: G at +1
                    const bool g1 = ((hb >> 1) & 1u) && !((lb >> 1) & 1u);
                    const bool g2 = ((hb >> 2) & 1u) && !((lb >> 2) & 1u);
                    total += sc->stop_lo[g1 ? 2 : (g2 ? 1 : 0)];
                }
                keep = total > 0.0;
                cd.contig = c;
                cd.left = strand ? (L - 1 - (t + 2)) : s;
                cd.right = strand ? (L - 1 - s) : (t + 2);
                cd.orf = sr.orf;
                cd.total = total;
                cd.gc = (int16_t)gc;
                cd.strand = strand ? -1 : 1;
                cd.flags = (uint8_t)((arc ? 1 : 0) | (virt ? 2 : 0) | (pstop ? 4 : 0));
                cd.len = len;
            }
            }
        }
        // reverse strand: the survivor with the largest s has the smallest forward right end, so the chain meets it
        // first among the starts of its ORF and may skip the lookup of the ORF's shared back-pointer.  Decidable here
        // only when all starts of the ORF sit in this warp's 32 slots (starts of an ORF are contiguous, ascending in s).
        const uint32_t km = __ballot_sync(0xffffffffu, keep);
        const uint32_t same = __match_any_sync(0xffffffffu, orf_id);   // lanes holding starts of my ORF
        const uint32_t vm = __ballot_sync(0xffffffffu, has && virt);
        if (keep && strand) {
            const bool whole = (int64_t)orf_base >= k0 && (int64_t)orf_base + orf_n <= k0 + 32;
            const bool last = (((same & km) >> lane) >> 1) == 0;        // no kept start of this ORF in a higher lane
            if (whole && last && !(vm & same & km)) cd.flags |= 8;
        }
        if (km) {
            const int cnt = __popc(km);
            unsigned long long base = 0;
            if (chunk) {
                if (cnt > chunk_left) {   // retire what is left of the old chunk, take a new one
                    if (lane < chunk_left && chunk_base + lane < cap_surv) surv[chunk_base + lane].contig = -1;
                    if (lane == 0) chunk_base = atomicAdd(&surv_counters[mi], (unsigned long long)chunk);
                    chunk_base = __shfl_sync(0xffffffffu, chunk_base, 0);
                    chunk_left = chunk;
                }
                base = chunk_base;
                chunk_base += cnt;
                chunk_left -= cnt;
            } else {
                if (lane == 0) base = atomicAdd(&surv_counters[mi], (unsigned long long)cnt);
                base = __shfl_sync(0xffffffffu, base, 0);
            }
            if (keep) {
                const unsigned long long kk = base + __popc(km & ((1u << lane) - 1u));
                if (kk < cap_surv) {
                    surv[kk] = cd;
                    const uint32_t br = (uint32_t)((w0 * 32 + cd.right) >> 7), bl = (uint32_t)((w0 * 32 + cd.left) >> 7);
                    atomicAdd(&bcount[br], 1u);
                    // a LONG candidate (spanning >= PIECE_BUCKETS bucket boundaries, > 8 kb: rare) marks the boundaries it spans
                    // in a difference array; the plan kernel finds all shorter ones by looking PIECE_BUCKETS buckets ahead
                    if (br >= bl + PIECE_BUCKETS) { atomicAdd(&cover[bl + 1], 1u); atomicAdd(&cover[br + 1], 0xffffffffu); }
                }
            }
        }
    }
    for (int i = lane; i < chunk_left; i += 32)   // unused tail of the warp's last chunk
        if (chunk_base + i < cap_surv) surv[chunk_base + i].contig = -1;
}

// survivors -> position buckets; boff holds exclusive offsets and is advanced to the bucket ends
__global__ void __launch_bounds__(256) k_scatter(const Cand *const *surv_m, Cand *const *sorted_m, uint32_t *const *boff_m,
                                                 const unsigned long long *surv_counters, const int64_t *__restrict__ cwoff,
                                                 uint32_t cap_surv) {
    const int mi = blockIdx.y;
    unsigned long long n = surv_counters[mi];
    if (n > cap_surv) n = cap_surv;
    const Cand *surv = surv_m[mi];
    Cand *sorted = sorted_m[mi];
    uint32_t *boff = boff_m[mi];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < (int64_t)n; k += (int64_t)gridDim.x * blockDim.x) {
        Cand cd = surv[k];
        if (cd.contig < 0) continue;   // slot reserved but not filled (k_final)
        uint32_t pos = atomicAdd(&boff[(cwoff[cd.contig] * 32 + cd.right) >> 7], 1u);
        sorted[pos] = cd;
    }
}

// ------------------------------------------------------------------------------------------------
// chain DP: one warp per (contig, model)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double overlap_penalty(int left_g, int len_g, int strand_g, int right_h, int len_h, int strand_h) {
    int ov = right_h - left_g + 1;
    if (ov <= 0) return 0.0;
    double k;
    if (strand_h == strand_g) {
        int up = strand_h > 0 ? len_h : len_g, down = strand_h > 0 ? len_g : len_h;
        k = up >= down ? 1.0 : 0.4;
    } else k = strand_h > 0 ? 1.0 : 2.0;
    double o = (double)ov;
    return o * log(1.0 + 0.5 * k * (o / (double)len_h + o / (double)len_g));
}


// The most recent CHAIN_RING predecessor states of the warp's contig live in shared memory; older ones
// (rarely needed: only genes spanning more than CHAIN_RING states) are read back from global memory.
// Candidates of one 128-nt bucket sit in registers (one per lane); all members of a group (same right
// end) are evaluated in parallel, each lane walking the predecessor window of its own member.
constexpr int CHAIN_RING = 64;
constexpr int GP_CACHE = 8;
struct ChainRing {
    int right[CHAIN_RING], left[CHAIN_RING], ls[CHAIN_RING], idx[CHAIN_RING], parg[CHAIN_RING];
    double cum[CHAIN_RING], pmax[CHAIN_RING];
    uint32_t gp_orf[GP_CACHE];  // small cache of the shared back-pointers of the reverse-strand ORFs in flight
    int gp_val[GP_CACHE];
};

struct ChainState {
    int na;          // states visible as predecessors (all have right < current group's right end)
    int head;        // states written
    double pm_val;   // running prefix maximum over written states
    int pm_arg;
    int gp_next;     // round-robin slot of the pointer cache
};

// Appends one predecessor state.  `owner` holds the values; cum is broadcast so that every lane tracks the running
// prefix maximum in registers.
__device__ __forceinline__ void chain_append(const ChainBufs &B, ChainRing &R, uint32_t cs, ChainState &S, int owner, int gi,
                                             double my_cum, int right, int left, int ls) {
    const double cum = __shfl_sync(0xffffffffu, my_cum, owner);
    if (S.head == 0 || cum > S.pm_val) { S.pm_val = cum; S.pm_arg = gi; }
    if (lane_id() == owner) {
        const int k = S.head, r = k & (CHAIN_RING - 1);
        R.right[r] = right; R.left[r] = left; R.ls[r] = ls; R.idx[r] = gi; R.cum[r] = cum; R.pmax[r] = S.pm_val; R.parg[r] = S.pm_arg;
        B.av_right[cs + k] = right; B.av_left[cs + k] = left; B.av_ls[cs + k] = ls; B.av_idx[cs + k] = gi;
        B.av_cum[cs + k] = cum; B.av_pmax[cs + k] = S.pm_val; B.av_parg[cs + k] = S.pm_arg;
    }
    S.head++;
}

// Per-lane evaluation of one candidate against the visible predecessor states (k < S.na).
//   bp_non : best non-overlapping predecessor (sorted index, -1 none)     b_non : its chain score (0 if none)
//   bp_ov  : best overlapping predecessor that beats it strictly, else -1  b_ov  : its value
__device__ __forceinline__ void chain_eval_lane(const ChainBufs &B, const ChainRing &R, uint32_t cs, const ChainState &S, int gl,
                                                int glen, int gstrand, double &b_non, int &bp_non, double &b_ov, int &bp_ov) {
    const int ring_lo = S.head - CHAIN_RING;  // states k >= ring_lo are in shared memory
    double best_v = 0.0;
    int best_p = -1;
    int k = S.na - 1;
    for (; k >= 0; k--) {
        const bool inr = k >= ring_lo;
        const int r = k & (CHAIN_RING - 1);
        const int rh = inr ? R.right[r] : B.av_right[cs + k];
        if (rh < gl) break;
        const int lh = inr ? R.left[r] : B.av_left[cs + k];
        if (lh < gl) {
            const int ls = inr ? R.ls[r] : B.av_ls[cs + k];
            const double ch = inr ? R.cum[r] : B.av_cum[cs + k];
            const double v = ch - overlap_penalty(gl, glen, gstrand, rh, ls >> 1, (ls & 1) ? -1 : 1);
            if (v >= best_v && v > 0.0) { best_v = v; best_p = k; }  // descending k: ties go to the earliest state
        }
    }
    b_non = 0.0;
    bp_non = -1;
    if (k >= 0) {
        const bool inr = k >= ring_lo;
        const double pm = inr ? R.pmax[k & (CHAIN_RING - 1)] : B.av_pmax[cs + k];
        if (pm > 0.0) { b_non = pm; bp_non = inr ? R.parg[k & (CHAIN_RING - 1)] : B.av_parg[cs + k]; }
    }
    b_ov = 0.0;
    bp_ov = -1;
    if (best_p >= 0 && best_v > b_non) {  // an overlap wins only if strictly better
        b_ov = best_v;
        bp_ov = best_p >= ring_lo ? R.idx[best_p & (CHAIN_RING - 1)] : B.av_idx[cs + best_p];
    }
}

// The same evaluation for ONE candidate with the whole warp: lane l looks at predecessor state top - l.  The states
// are ordered by right end, so the overlapping ones (right >= gl) are the last few; the first state below them gives
// the best non-overlapping predecessor through its prefix maximum.  All arguments and results are warp-uniform.
__device__ __forceinline__ void chain_eval_warp(const ChainBufs &B, const ChainRing &R, uint32_t cs, const ChainState &S, int gl,
                                                int glen, int gstrand, double &b_non, int &bp_non, double &b_ov, int &bp_ov) {
    const int lane = lane_id();
    const int ring_lo = S.head - CHAIN_RING;  // states k >= ring_lo are in shared memory
    double best_v = 0.0;
    int best_k = 0x7fffffff;
    int k_non = -1;
    for (int top = S.na - 1; top >= 0; top -= 32) {
        const int k = top - lane;
        const bool valid = k >= 0;
        const bool inr = k >= ring_lo;
        const int r = k & (CHAIN_RING - 1);
        int rh = 0x7fffffff, lh = 0x7fffffff;
        if (valid) {
            rh = inr ? R.right[r] : B.av_right[cs + k];
            lh = inr ? R.left[r] : B.av_left[cs + k];
        }
        if (valid && rh >= gl && lh < gl) {
            const int ls = inr ? R.ls[r] : B.av_ls[cs + k];
            const double ch = inr ? R.cum[r] : B.av_cum[cs + k];
            const double v = ch - overlap_penalty(gl, glen, gstrand, rh, ls >> 1, (ls & 1) ? -1 : 1);
            if (v >= best_v && v > 0.0) { best_v = v; best_k = k; }  // a lane meets its states in descending k: ties go to the earliest
        }
        const uint32_t below = __ballot_sync(0xffffffffu, valid && rh < gl);
        if (below) { k_non = top - (__ffs(below) - 1); break; }
    }
    // best overlapping predecessor over the lanes: largest v, ties to the smallest k (v > 0: bit order = value order)
    const unsigned long long bits = (unsigned long long)__double_as_longlong(best_v);
    const uint32_t hi = __reduce_max_sync(0xffffffffu, (uint32_t)(bits >> 32));
    const bool in2 = (uint32_t)(bits >> 32) == hi;
    const uint32_t lo = __reduce_max_sync(0xffffffffu, in2 ? (uint32_t)bits : 0u);
    const bool in3 = in2 && (uint32_t)bits == lo;
    const uint32_t kmin = __reduce_min_sync(0xffffffffu, in3 ? (uint32_t)best_k : 0xffffffffu);
    const double vmax = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    b_non = 0.0;
    bp_non = -1;
    if (k_non >= 0) {
        const bool inr = k_non >= ring_lo;
        const double pm = inr ? R.pmax[k_non & (CHAIN_RING - 1)] : B.av_pmax[cs + k_non];
        if (pm > 0.0) { b_non = pm; bp_non = inr ? R.parg[k_non & (CHAIN_RING - 1)] : B.av_parg[cs + k_non]; }
    }
    b_ov = 0.0;
    bp_ov = -1;
    if (vmax > 0.0 && kmin != 0x7fffffffu && vmax > b_non) {  // an overlap wins only if strictly better
        b_ov = vmax;
        bp_ov = (int)kmin >= ring_lo ? R.idx[kmin & (CHAIN_RING - 1)] : B.av_idx[cs + kmin];
    }
}

// current shared pointer of a reverse-strand ORF (warp-uniform call)
__device__ __forceinline__ int gp_lookup(const ChainBufs &B, const ChainRing &R, uint32_t orf, int flags) {
    if (flags & 8) return GPTR_UNSET;  // first surviving start of its ORF in chain order: nothing stored yet
    const int lane = lane_id();
    const uint32_t hit = __ballot_sync(0xffffffffu, lane < GP_CACHE && R.gp_orf[lane & (GP_CACHE - 1)] == orf);
    if (hit) return R.gp_val[__ffs(hit) - 1];
    return B.gptr[orf];
}

__device__ __forceinline__ void gp_store(const ChainBufs &B, ChainRing &R, ChainState &S, uint32_t orf, int val) {
    const int lane = lane_id();
    const uint32_t hit = __ballot_sync(0xffffffffu, lane < GP_CACHE && R.gp_orf[lane & (GP_CACHE - 1)] == orf);
    int slot;
    if (hit) slot = __ffs(hit) - 1;
    else { slot = S.gp_next; S.gp_next = (S.gp_next + 1) & (GP_CACHE - 1); }
    if (lane == 0) {
        R.gp_orf[slot] = orf;
        R.gp_val[slot] = val;
        B.gptr[orf] = val;
    }
    __syncwarp();
}

// Processes the candidates held one per lane (valid lanes: `pending`) in chain order.  All candidates of a bucket end
// within the same 128 positions, so (right & 127) orders the groups; warp-wide minima are single redux instructions.
__device__ __forceinline__ void chain_bucket(const ChainBufs &B, ChainRing &R, uint32_t cs, ChainState &S, uint32_t bs, const Cand &me,
                                             bool pending) {
    const int lane = lane_id();
    const uint32_t rr = (uint32_t)me.right & 127u;
    while (true) {
        const uint32_t rmin = __reduce_min_sync(0xffffffffu, pending ? rr : 0xffffffffu);
        if (rmin == 0xffffffffu) break;
        const bool ing = pending && rr == rmin;
        pending = pending && !ing;
        double b_non = 0, b_ov = 0;
        int bp_non = -1, bp_ov = -1;
        for (uint32_t gm = __ballot_sync(0xffffffffu, ing); gm; gm &= gm - 1) {  // the members one by one, 32 states at a time
            const int o = __ffs(gm) - 1;
            double r_bn, r_bo;
            int r_pn, r_po;
            chain_eval_warp(B, R, cs, S, __shfl_sync(0xffffffffu, me.left, o), __shfl_sync(0xffffffffu, me.len, o),
                            __shfl_sync(0xffffffffu, (int)me.strand, o), r_bn, r_pn, r_bo, r_po);
            if (lane == o) { b_non = r_bn; bp_non = r_pn; b_ov = r_bo; bp_ov = r_po; }
        }
        const double cumg = me.total + (bp_ov >= 0 ? b_ov : b_non);
        if (ing) {
            B.cum[bs + lane] = cumg;
            B.prev[bs + lane] = bp_ov >= 0 ? bp_ov : bp_non;
        }
        // reverse-strand members: shared pointer of their ORF, then append (ascending left; a real start before a
        // virtual one at the same position)
        uint32_t rev = __ballot_sync(0xffffffffu, ing && me.strand < 0);
        const uint32_t dir = __ballot_sync(0xffffffffu, ing && me.strand > 0);
        while (rev) {
            int o = __ffs(rev) - 1;
            if (rev & (rev - 1)) {
                const bool in = (rev >> lane) & 1u;
                const uint32_t key = in ? (((uint32_t)me.left << 1) | ((me.flags >> 1) & 1u)) : 0xffffffffu;
                const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
                o = __ffs(__ballot_sync(0xffffffffu, in && key == kmin)) - 1;
            }
            rev &= ~(1u << o);
            const uint32_t gorf = __shfl_sync(0xffffffffu, me.orf, o);
            const int gfl = __shfl_sync(0xffffffffu, (int)me.flags, o);
            int gp = gp_lookup(B, R, gorf, gfl);
            // the owner resolves the pointer, everybody learns it
            if (gp == GPTR_UNSET) gp = bp_non;
            if (bp_ov >= 0) gp = bp_ov;
            gp = __shfl_sync(0xffffffffu, gp, o);
            gp_store(B, R, S, gorf, gp);
            chain_append(B, R, cs, S, o, (int)(bs + o), cumg, me.right, me.left, me.len * 2 + 1);
        }
        // direct-strand members: the stop keeps one state, its best start (ties: smallest left)
        if (dir) {
            int o = __ffs(dir) - 1;
            if (dir & (dir - 1)) {
                const bool in = (dir >> lane) & 1u;
                const unsigned long long bits = (unsigned long long)__double_as_longlong(cumg);  // cumg > 0: bit order = value order
                const uint32_t hi = __reduce_max_sync(0xffffffffu, in ? (uint32_t)(bits >> 32) : 0u);
                const bool in2 = in && (uint32_t)(bits >> 32) == hi;
                const uint32_t lo = __reduce_max_sync(0xffffffffu, in2 ? (uint32_t)bits : 0u);
                const bool in3 = in2 && (uint32_t)bits == lo;
                const uint32_t lmin = __reduce_min_sync(0xffffffffu, in3 ? (uint32_t)me.left : 0xffffffffu);
                o = __ffs(__ballot_sync(0xffffffffu, in3 && (uint32_t)me.left == lmin)) - 1;
            }
            chain_append(B, R, cs, S, o, (int)(bs + o), cumg, me.right, me.left, me.len * 2);
        }
        S.na = S.head;
        __syncwarp();
    }
}

// Cut points.  Candidates are bucketed by right end; the boundary in front of a bucket is safe when no candidate spans
// it: nothing on its right can then overlap, or share a reverse-strand ORF with, anything on its left, and the only
// coupling left is the additive best score of the prefix.  A candidate ending in bucket b spans boundary k <= b iff its
// left end lies before the boundary, so with minleft[b] = smallest left end among the candidates ending in b,
//   boundary k is safe  <=>  min(minleft[k .. k+H)) >= position of k   and no LONG candidate (>= H boundaries) spans k,
// the long ones being counted by k_final in a prefix-summed difference array (ChainBufs::cover).  H = PIECE_BUCKETS.
// k_scatter leaves a bucket unordered.  Thread per candidate: its rank inside its bucket in chain order (right end, direct strand
// first, left end, real start before a virtual one at the same position) = the number of members with a smaller key; the
// record moves to its final place in `out` (ChainBufs::sorted).  The member with rank 0 also stores the bucket's smallest left end
// (empty buckets keep a stale value: k_chain_plan tests for emptiness first).  Neighbouring lanes hold the members of one bucket,
// so the member loop runs over lines the warp has just loaded.
__global__ void __launch_bounds__(256) k_bucket_order(const ChainBufs *bufs_m, const Cand *const *tmp_m, Cand *const *out_m,
                                                      const int64_t *__restrict__ cwoff, int64_t n_buckets) {
    const int mi = blockIdx.y;
    const ChainBufs B = bufs_m[mi];
    const Cand *__restrict__ tmp = tmp_m[mi];
    Cand *__restrict__ out = out_m[mi];
    const int64_t n = B.bend[n_buckets - 1];
    auto key_of = [](int right, int strand, int left, int flags) {
        return ((unsigned long long)((uint32_t)right & 127u) << 40) | ((unsigned long long)(strand < 0) << 39) |
               (((unsigned long long)(uint32_t)left << 1) | (((uint32_t)flags >> 1) & 1u));
    };
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const Cand me = tmp[k];
        const int64_t b = (cwoff[me.contig] * 32 + me.right) >> 7;
        const uint32_t s = b ? B.bend[b - 1] : 0u, e = B.bend[b];
        uint32_t rank = 0, minl = (uint32_t)me.left;
        if (e - s > 1) {
            const unsigned long long mine = key_of(me.right, me.strand, me.left, me.flags);
            for (uint32_t j = s; j < e; j++) {
                const int4 a = *reinterpret_cast<const int4 *>(&tmp[j]);            // contig, left, right, orf
                const int4 c = *(reinterpret_cast<const int4 *>(&tmp[j]) + 1);        // total (2 words), gc | strand | flags, len
                const unsigned long long kj = key_of(a.z, (int)(int8_t)((uint32_t)c.z >> 16), a.y, (int)(((uint32_t)c.z >> 24) & 0xffu));
                rank += kj < mine;
                minl = min(minl, (uint32_t)a.y);
            }
        }
        out[s + rank] = me;
        if (rank == 0) B.minleft[b] = minl;
    }
}

// work list of a model (k_chain_plan): wc[0] long pieces at the front of `work`, wc[2] others from its back, wc[1] the consumers' cursor
constexpr int CHAIN_BIG_PIECE = 96;
__device__ __forceinline__ int work_count(const int *wc) { return wc[0] + wc[2]; }
__device__ __forceinline__ int work_slot(const ChainBufs &B, const int *wc, int item, int n_slots) {
    const int nb = wc[0];
    return item < nb ? B.work[item] : B.work[n_slots - 1 - (item - nb)];
}

// Thread per piece slot = aligned window of PIECE_BUCKETS buckets of a contig; a window contributes its first safe
// boundary as a cut (window 0: the contig start), and the piece runs up to the next window's cut.  Fully parallel, also
// inside one long contig.
__global__ void __launch_bounds__(128) k_chain_plan(const ChainBufs *bufs_m, const int64_t *__restrict__ cwoff,
                                                    const int32_t *__restrict__ clen, const int32_t *__restrict__ slotbase,
                                                    const int32_t *__restrict__ slot_contig, int32_t n_slots, int *work_counters) {
    const int mi = blockIdx.y;
    const ChainBufs B = bufs_m[mi];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int c = slot_contig[s];
    const int sb = slotbase[c], m = s - sb, nwin = slotbase[c + 1] - sb;
    const int L = clen[c];
    const int64_t b_first = cwoff[c] >> 2;  // slots start on 128-nt boundaries
    const int nbk = (max(L - 1, 0) >> 7) + 1;
    const uint32_t cs = b_first ? B.bend[b_first - 1] : 0u, ce = B.bend[b_first + nbk - 1];
    if (ce == cs) return;
    const uint32_t *cov = B.cover + b_first + 1;  // cov[rb]: LONG candidates spanning the boundary in front of relative bucket rb
    const uint32_t *ml = B.minleft + b_first;     // ml[rb]: smallest (contig-relative) left end among the candidates ending in rb
    const uint32_t *be = B.bend + b_first;        // a bucket is empty when its end equals its predecessor's (its minleft entry is stale then)
    auto ml_at = [&](int rb) { return be[rb] == ((b_first + rb) ? be[rb - 1] : 0u) ? 0xffffffffu : ml[rb]; };
    auto cut_of = [&](int win) {  // first safe boundary of the window, -1 if none
        if (win == 0) return 0;
        const int lo = win * PIECE_BUCKETS, hi = min(lo + PIECE_BUCKETS - 1, nbk - 1);
        uint32_t run = 0xffffffffu;  // min(minleft[rb .. hi + H)), built from the far end
        for (int rb = min(hi + PIECE_BUCKETS - 1, nbk - 1); rb > hi; rb--) run = min(run, ml_at(rb));
        int best = -1;
        for (int rb = hi; rb >= lo; rb--) {
            run = min(run, ml_at(rb));
            if (run >= (uint32_t)rb * 128u && cov[rb] == 0u) best = rb;
        }
        return best;
    };
    const int mine = cut_of(m);
    if (mine < 0) return;
    int last = nbk - 1;
    for (int win = m + 1; win < nwin; win++) {
        const int nx = cut_of(win);
        if (nx >= 0) { last = nx - 1; break; }
    }
    const uint32_t ps = (b_first + mine) ? B.bend[b_first + mine - 1] : 0u, pe = B.bend[b_first + last];
    if (pe == ps) return;  // nothing ends inside this piece
    B.piece_b0[s] = (int32_t)(b_first + mine);
    B.piece_b1[s] = (int32_t)(b_first + last);
    // pieces in processing order: the long ones from the front of `work` (taken first, so that the longest serial chains start
    // at once), the others from its back
    if (pe - ps >= (uint32_t)CHAIN_BIG_PIECE) B.work[atomicAdd(&work_counters[4 * mi], 1)] = s;
    else B.work[n_slots - 1 - atomicAdd(&work_counters[4 * mi + 2], 1)] = s;
}

__global__ void __launch_bounds__(128, 8) k_chain(const ChainBufs *bufs_m, int *work_counters, int n_slots, int long_only) {
    __shared__ ChainRing rings[4];
    const int mi = blockIdx.y;
    const ChainBufs B = bufs_m[mi];
    const int lane = lane_id();
    ChainRing &R = rings[threadIdx.x >> 5];
    // long_only: the long pieces at the front of the work list (their own cursor); k_chain_flat takes the others
    const int n_work = long_only ? work_counters[4 * mi] : work_count(work_counters + 4 * mi);
    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&work_counters[4 * mi + (long_only ? 3 : 1)], 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_work) break;
        const int slot = work_slot(B, work_counters + 4 * mi, item, n_slots);
        const int64_t b_first = B.piece_b0[slot], b_last = B.piece_b1[slot];
        const uint32_t cs = b_first ? B.bend[b_first - 1] : 0u;
        const uint32_t ce = B.bend[b_last];
        ChainState S;
        S.na = 0; S.head = 0; S.pm_val = 0.0; S.pm_arg = -1; S.gp_next = 0;
        if (lane < GP_CACHE) R.gp_orf[lane] = 0xffffffffu;
        __syncwarp();
        // the piece's candidates are contiguous in `sorted`; walk them in register-sized blocks that end on bucket boundaries
        uint32_t prev_end = cs;
        uint32_t nxt_end = b_first <= b_last ? (b_first + lane <= b_last ? B.bend[b_first + lane] : ce) : ce;
        for (int64_t bb = b_first; bb <= b_last && prev_end < ce; bb += 32) {
            const uint32_t my_end = nxt_end;
            if (bb + 32 <= b_last) nxt_end = bb + 32 + lane <= b_last ? B.bend[bb + 32 + lane] : ce;  // prefetch next block of bucket ends
            uint32_t my_start = __shfl_up_sync(0xffffffffu, my_end, 1);
            if (lane == 0) my_start = prev_end;
            prev_end = __shfl_sync(0xffffffffu, my_end, 31);
            uint32_t nonempty = __ballot_sync(0xffffffffu, my_end > my_start);
            // software pipeline: the next non-empty bucket is loaded while the current one is processed
            Cand nxt;
            nxt.right = 0;
            uint32_t nbs = 0, nn = 0;
            if (nonempty) {
                const int bl = __ffs(nonempty) - 1;
                nbs = __shfl_sync(0xffffffffu, my_start, bl);
                nn = __shfl_sync(0xffffffffu, my_end, bl) - nbs;
                if ((uint32_t)lane < nn && nn <= 32) nxt = B.sorted[nbs + lane];
            }
            while (nonempty) {
                nonempty &= nonempty - 1;
                const Cand me = nxt;
                const uint32_t bs = nbs, n = nn;
                if (nonempty) {
                    const int bl = __ffs(nonempty) - 1;
                    nbs = __shfl_sync(0xffffffffu, my_start, bl);
                    nn = __shfl_sync(0xffffffffu, my_end, bl) - nbs;
                    if ((uint32_t)lane < nn && nn <= 32) nxt = B.sorted[nbs + lane];
                }
                if (n <= 32) {
                    chain_bucket(B, R, cs, S, bs, me, (uint32_t)lane < n);
                } else {
                    // rare: more than 32 candidates in one 128-nt bucket.  Their right ends are handled in ascending order,
                    // 32 members of a group at a time (a group larger than 32 falls back to member-at-a-time semantics
                    // only in the choice of the kept direct-strand state, which is re-resolved below).
                    int cur = -1;
                    while (true) {
                        int r = 0x7fffffff;
                        for (uint32_t k = bs + lane; k < bs + n; k += 32) {
                            const int rr = B.sorted[k].right;
                            if (rr > cur) r = min(r, rr);
                        }
                        r = warp_min_i(r);
                        if (r == 0x7fffffff) break;
                        cur = r;
                        // evaluate all members of this group; append reverse members immediately, track the best direct one
                        double bc = -1e300;
                        int bleft = 0x7fffffff, bidx = -1, blen = 0;
                        const ChainState S0 = S;  // members see the states before the group
                        for (uint32_t k0 = bs; k0 < bs + n; k0 += 32) {
                            const uint32_t k = k0 + lane;
                            Cand g;
                            g.right = -1;
                            if (k < bs + n) g = B.sorted[k];
                            const bool ing = g.right == r;
                            double b_non = 0, b_ov = 0;
                            int bp_non = -1, bp_ov = -1;
                            ChainState Sv = S0;
                            Sv.head = S.head;
                            if (ing) chain_eval_lane(B, R, cs, Sv, g.left, g.len, g.strand, b_non, bp_non, b_ov, bp_ov);
                            const double cumg = g.total + (bp_ov >= 0 ? b_ov : b_non);
                            if (ing) { B.cum[k] = cumg; B.prev[k] = bp_ov >= 0 ? bp_ov : bp_non; }
                            uint32_t rev = __ballot_sync(0xffffffffu, ing && g.strand < 0);
                            while (rev) {
                                const int o = __ffs(rev) - 1;
                                rev &= rev - 1;
                                const uint32_t gorf = __shfl_sync(0xffffffffu, g.orf, o);
                                const int o_non = __shfl_sync(0xffffffffu, bp_non, o), o_ov = __shfl_sync(0xffffffffu, bp_ov, o);
                                int gp = gp_lookup(B, R, gorf, 0);
                                if (gp == GPTR_UNSET) gp = o_non;
                                if (o_ov >= 0) gp = o_ov;
                                gp_store(B, R, S, gorf, gp);
                                chain_append(B, R, cs, S, o, (int)(k0 + o), cumg, g.right, g.left, g.len * 2 + 1);
                            }
                            const bool dirm = ing && g.strand > 0;
                            double c2 = dirm ? cumg : -1e300;
                            int l2 = dirm ? g.left : 0x7fffffff, o2 = lane;
#pragma unroll
                            for (int o = 16; o; o >>= 1) {
                                const double oc = __shfl_xor_sync(0xffffffffu, c2, o);
                                const int ol = __shfl_xor_sync(0xffffffffu, l2, o), oo = __shfl_xor_sync(0xffffffffu, o2, o);
                                if (oc > c2 || (oc == c2 && ol < l2)) { c2 = oc; l2 = ol; o2 = oo; }
                            }
                            if (c2 > bc || (c2 == bc && l2 < bleft)) {
                                bc = c2; bleft = l2; bidx = (int)(k0 + o2); blen = __shfl_sync(0xffffffffu, g.len, o2);
                            }
                        }
                        if (bidx >= 0 && bc > -1e299) chain_append(B, R, cs, S, 0, bidx, bc, r, bleft, blen * 2);
                        S.na = S.head;
                        __syncwarp();
                    }
                }
            }
        }
        // the traceback is a chain of dependent loads: it runs in k_chain_trace, a thread per piece, instead of on lane 0
        // of this warp with 31 lanes idle.  gene_count[slot] carries the end of the best chain until then.
        if (lane == 0) {
            B.gene_count[slot] = S.head > 0 ? S.pm_arg : -1;
            B.piece_total[slot] = S.head > 0 ? S.pm_val : 0.0;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// chain DP, thread per piece (round 2).  The warp-per-piece kernel above spends ~300 warp instructions per candidate on
// warp-wide bookkeeping (minima, ballots, broadcasts) for buckets that hold 1.2 candidates on average.  Here every thread
// runs the plain serial recurrence of its piece, in the order of oracle/gmhmm_oracle.c orc_chain (k_bucket_order has
// sorted every bucket: right end, direct strand first, left end, real start before virtual): groups of equal right end see
// the states written before the group; the direct-strand members' best becomes the stop's one state, appended before the
// group's reverse-strand members.
// Written as nested loops (pieces x candidates x states) the 32 pieces of a warp would sit at 32 different program points
// and every instruction would be issued for 3-4 lanes (measured).  Instead each lane is a small state machine that WANTS
// one of four sections — fetch the next candidate (and, at the end of a piece, the next piece), scan predecessor states,
// evaluate one overlap penalty (two divisions and a log: the expensive one), finish the candidate — and every iteration the
// warp runs the ONE section most lanes want; the others wait a turn.  The sections are executed with a third to a half of
// the lanes instead of a tenth.
// The scan over predecessor states is the bulk of the work (24 states per candidate on average, most of them nested inside
// the candidate and skipped): it reads 8-byte {right, left} records, four per 32-byte sector, the group below requested ahead.
// Pieces are handed out one at a time per THREAD: a lane that finishes its piece takes the next while its neighbours go on.
// ------------------------------------------------------------------------------------------------
#ifndef CHAIN_MINB
#define CHAIN_MINB 5
#endif
enum { W_FETCH = 0, W_SCAN = 1, W_EVAL = 2, W_FINAL = 3, W_DONE = 4 };

__device__ __forceinline__ void load_group(const int2 *rl, int gb, int2 (&out)[4]) {   // four states' {right, left}: one aligned 32-byte sector
    const int4 a = *reinterpret_cast<const int4 *>(rl + gb), b = *reinterpret_cast<const int4 *>(rl + gb + 2);
    out[0] = make_int2(a.x, a.y); out[1] = make_int2(a.z, a.w); out[2] = make_int2(b.x, b.y); out[3] = make_int2(b.z, b.w);
}

__global__ void __launch_bounds__(128, CHAIN_MINB) k_chain_flat(const ChainBufs *bufs_m, int *work_counters, int n_slots, int skip_long) {
    const int mi = blockIdx.y;
    const ChainBufs B = bufs_m[mi];
    const Cand *__restrict__ C = B.sorted;
    const int *wc = work_counters + 4 * mi;
    const int n_work = work_count(wc);
    const int first_item = skip_long ? wc[0] : 0;   // the long pieces at the front of the list go to the warp-per-piece kernel
    // States of a piece live at the piece's candidate range [a0, ...) of two arrays: rl = {right, left} (8 bytes: what the scan
    // reads, four states per 32-byte sector) and det = {len*2+rev, index, index of the prefix maximum | cum, prefix maximum}.
    int2 *__restrict__ rl = B.st_rl;
    ChainSt *__restrict__ det = B.st;
    // piece
    int slot = 0;
    uint32_t k = 0, ce = 0;              // next candidate, end of the piece
    int a0 = 0, na = 0, head = 0, pm_arg = -1;   // first state (absolute index), visible / written states
    double pm_val = 0.0;
    bool have_piece = false;
    // group (equal right end): its pending direct-strand state
    int grp_right = -1, bidx = -1, bleft = 0, blen = 0;
    double bc = 0.0;
    // candidate under evaluation
    Cand g, gn;                          // gn: the next candidate of the piece, requested while g is evaluated
    g.left = 0; g.right = 0; g.len = 1; g.strand = 1; g.flags = 0; g.orf = 0; g.total = 0.0;
    gn = g;
    uint32_t gi = 0;
    int gp_pre = GPTR_UNSET;             // shared back-pointer of g's ORF (reverse strand), requested at fetch time
    double fin_pm = 0.0;                 // prefix maximum / its index at the state the scan stopped at, requested when it stops
    int fin_parg = -1;
    int ak = -1;                         // absolute index of the state under examination (< a0: none left)
    int2 cur[4], nxt[4];                 // the aligned group of four states holding ak, and the group below it (requested ahead)
#pragma unroll
    for (int u = 0; u < 4; u++) cur[u] = nxt[u] = make_int2(0, 0);
    int ev_right = 0, best_idx = -1;
    double best_v = 0.0;
    int want = W_FETCH;

    auto append = [&](int idx, double cum, int right, int left, int ls) {
        if (head == 0 || cum > pm_val) { pm_val = cum; pm_arg = idx; }
        rl[a0 + head] = make_int2(right, left);
        det[a0 + head].i = make_int4(ls, idx, pm_arg, 0);
        det[a0 + head].d = make_double2(cum, pm_val);
        head++;
    };
    auto advance = [&]() {   // the scan moves to the next older state
        const bool cross = (ak & 3) == 0;
        ak--;
        if (cross) {
#pragma unroll
            for (int u = 0; u < 4; u++) cur[u] = nxt[u];
            const int gb = (ak & ~3) - 4;
            if (gb >= 0 && gb + 3 >= a0) load_group(rl, gb, nxt);
        }
    };

    while (true) {
        const uint32_t m_fetch = __ballot_sync(0xffffffffu, want == W_FETCH);
        const uint32_t m_scan = __ballot_sync(0xffffffffu, want == W_SCAN);
        const uint32_t m_eval = __ballot_sync(0xffffffffu, want == W_EVAL);
        const uint32_t m_final = __ballot_sync(0xffffffffu, want == W_FINAL);
        if (!(m_fetch | m_scan | m_eval | m_final)) break;
        int pick = W_FETCH, most = __popc(m_fetch);
        if (__popc(m_scan) > most) { pick = W_SCAN; most = __popc(m_scan); }
        if (__popc(m_final) > most) { pick = W_FINAL; most = __popc(m_final); }
        if (__popc(m_eval) > most) pick = W_EVAL;   // (a section nobody wants must never be picked: the loop would spin)

        if (pick == W_FETCH) {
            if (want == W_FETCH) {
                if (have_piece && k >= ce) {   // piece finished
                    if (bidx >= 0) append(bidx, bc, grp_right, bleft, blen * 2);
                    B.gene_count[slot] = head > 0 ? pm_arg : -1;   // end of the best chain, until k_chain_trace
                    B.piece_total[slot] = head > 0 ? pm_val : 0.0;
                    have_piece = false;
                }
                if (!have_piece) {
                    const int item = first_item + atomicAdd(&work_counters[4 * mi + 1], 1);
                    if (item >= n_work) want = W_DONE;
                    else {
                        slot = work_slot(B, wc, item, n_slots);
                        const int64_t b_first = B.piece_b0[slot], b_last = B.piece_b1[slot];
                        const uint32_t cs = b_first ? B.bend[b_first - 1] : 0u;
                        ce = B.bend[b_last];   // > cs: k_chain_plan lists no empty piece
                        k = cs;
                        a0 = (int)cs;
                        na = 0; head = 0; pm_val = 0.0; pm_arg = -1;
                        grp_right = -1; bidx = -1;
                        have_piece = true;
                        gn = C[k];
                    }
                }
                if (want != W_DONE) {
                    g = gn;
                    gi = k++;
                    if (k < ce) gn = C[k];
                    // only this thread writes the pointer of the ORFs of its piece, and not before this candidate is finished
                    if (g.strand < 0 && !(g.flags & 8)) gp_pre = B.gptr[g.orf];
                    if (g.right != grp_right) {
                        if (bidx >= 0) { append(bidx, bc, grp_right, bleft, blen * 2); bidx = -1; }
                        na = head;
                        grp_right = g.right;
                    } else if (g.strand < 0 && bidx >= 0) {
                        append(bidx, bc, grp_right, bleft, blen * 2);
                        bidx = -1;
                    }
                    ak = a0 + na - 1;
                    best_v = 0.0;
                    best_idx = -1;
                    if (na > 0) {
                        const int gb = ak & ~3;
                        load_group(rl, gb, cur);
                        if (gb - 4 >= 0 && gb - 1 >= a0) load_group(rl, gb - 4, nxt);
                    }
                    want = W_SCAN;
                }
            }
        } else if (pick == W_SCAN) {
            // States whose right end lies inside the candidate, newest first, a group of four per turn: one comparison mask for
            // "ends before the candidate" (closes the scan) and one for "starts before it" (needs its penalty); the highest flagged
            // entry at or below ak decides, a group without flags is skipped whole (most states are nested inside the candidate).
            if (want == W_SCAN) {
                if (ak < a0) want = W_FINAL;
                else {
                    const int gb = ak & ~3;
                    uint32_t ends = 0, ovl = 0;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        ends |= (uint32_t)(cur[u].x < g.left) << u;
                        ovl |= (uint32_t)(cur[u].y < g.left) << u;
                    }
                    uint32_t valid = (2u << (ak & 3)) - 1u;                   // entries at or below ak ...
                    if (a0 > gb) valid &= ~((1u << (a0 - gb)) - 1u);          // ... that belong to this piece
                    const uint32_t ev = (ends | ovl) & valid;
                    if (!ev) {   // nothing here: on to the group below
                        ak = max(gb, a0) - 1;
                        if (ak >= a0) {
#pragma unroll
                            for (int u = 0; u < 4; u++) cur[u] = nxt[u];
                            const int nb = gb - 8;
                            if (nb >= 0 && nb + 3 >= a0) load_group(rl, nb, nxt);
                        } else want = W_FINAL;
                    } else {
                        const int u = 31 - __clz(ev);
                        ak = gb + u;
                        if ((ends >> u) & 1u) {
                            want = W_FINAL;
                            fin_pm = det[ak].d.y;
                            fin_parg = det[ak].i.z;
                        } else {
                            want = W_EVAL;
                            ev_right = u == 0 ? cur[0].x : (u == 1 ? cur[1].x : (u == 2 ? cur[2].x : cur[3].x));
                        }
                    }
                }
            }
        } else if (pick == W_EVAL) {
            if (want == W_EVAL) {
                const int4 di = det[ak].i;
                const double v = det[ak].d.x - overlap_penalty(g.left, g.len, g.strand, ev_right, di.x >> 1, (di.x & 1) ? -1 : 1);
                if (v >= best_v && v > 0.0) { best_v = v; best_idx = di.y; }   // descending index: ties go to the earliest state
                advance();
                want = W_SCAN;
            }
        } else {
            if (want == W_FINAL) {
                // best non-overlapping predecessor = prefix maximum at the first state ending before the candidate
                double b = 0.0;
                int bp = -1;
                if (ak >= a0 && fin_pm > 0.0) { b = fin_pm; bp = fin_parg; }
                const int bp_non = bp;
                const bool over = best_idx >= 0 && best_v > b;   // an overlap wins only if strictly better
                if (over) { b = best_v; bp = best_idx; }
                const double cumg = g.total + b;
                B.cum[gi] = cumg;
                B.prev[gi] = bp;
                if (g.strand > 0) {
                    if (bidx < 0 || cumg > bc) { bc = cumg; bidx = (int)gi; bleft = g.left; blen = g.len; }   // sorted by left: ties keep the smallest
                } else {
                    int gp = (g.flags & 8) ? GPTR_UNSET : gp_pre;   // the ORF's shared back-pointer (rule 6)
                    if (gp == GPTR_UNSET) gp = bp_non;
                    if (over) gp = bp;
                    B.gptr[g.orf] = gp;
                    append((int)gi, cumg, g.right, g.left, g.len * 2 + 1);
                }
                want = W_FETCH;
            }
        }
    }
}

// traceback of every piece (reverse-strand genes follow their ORF's shared pointer): thread per piece, the candidate and
// back-pointer loads of half a million pieces in flight at once
__global__ void __launch_bounds__(128) k_chain_trace(const ChainBufs *bufs_m, const int *work_counters, int n_slots) {
    const int mi = blockIdx.y;
    const ChainBufs B = bufs_m[mi];
    const int n_work = work_count(work_counters + 4 * mi);
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_work) return;
    const int slot = work_slot(B, work_counters + 4 * mi, item, n_slots);
    const int64_t b_first = B.piece_b0[slot], b_last = B.piece_b1[slot];
    const uint32_t cs = b_first ? B.bend[b_first - 1] : 0u;
    const uint32_t ce = B.bend[b_last];
    int g = B.gene_count[slot];
    int ng = 0;
    const int limit = (int)(ce - cs);
    int k = limit;
    while (g >= 0 && k > 0) {
        B.chain[cs + (--k)] = g;
        const Cand cd = B.sorted[g];
        g = cd.strand < 0 ? B.gptr[cd.orf] : B.prev[g];
        ng++;
    }
    for (int q = 0; q < ng; q++) B.chain[cs + q] = B.chain[cs + k + q];
    B.gene_count[slot] = ng;
}

// genes of all contigs, contiguous, ordered by (contig, left); per-model sum of round(logodd, 2) in cents
__global__ void __launch_bounds__(128) k_gather(const ChainBufs *bufs_m, mgc_gene *const *genes_m, const int32_t *__restrict__ slot_contig,
                                                int32_t n_slots, long long *model_cents) {
    const int mi = blockIdx.y;
    const ChainBufs B = bufs_m[mi];
    mgc_gene *genes = genes_m[mi];
    const int lane = lane_id();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    long long cents = 0;
    for (int64_t slot = warp; slot < n_slots; slot += nw) {
        const int ng = B.gene_count[slot];
        if (!ng) continue;
        const int64_t b_first = B.piece_b0[slot];
        const uint32_t cs = b_first ? B.bend[b_first - 1] : 0u;
        const uint32_t off = B.slot_gene_off[slot];
        const int32_t c = slot_contig[slot];
        for (int q = lane; q < ng; q += 32) {
            const Cand cd = B.sorted[B.chain[cs + q]];
            mgc_gene g;
            g.contig = c;
            g.left = cd.left + 1;
            g.right = cd.right + 1;
            g.strand = cd.strand;
            g.domain = cd.flags & 1;
            const int ps = (cd.flags >> 1) & 1, pe = (cd.flags >> 2) & 1;
            g.partial_left = cd.strand > 0 ? ps : pe;
            g.partial_right = cd.strand > 0 ? pe : ps;
            g.gc = cd.gc;
            g.logodd = cd.total;
            genes[off + q] = g;
            cents += llrint(cd.total * 100.0);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cents += __shfl_xor_sync(0xffffffffu, cents, o);
    if (lane == 0 && cents) atomicAdd((unsigned long long *)&model_cents[mi], (unsigned long long)cents);
}

// per contig: offset of its first gene and the chain total = sum of its pieces' totals, in order
__global__ void __launch_bounds__(256) k_contig_finish(const ChainBufs *bufs_m, const int32_t *__restrict__ slotbase, int32_t n_contigs) {
    const ChainBufs B = bufs_m[blockIdx.y];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_contigs) return;
    const int s0 = slotbase[c];
    B.gene_off[c] = B.slot_gene_off[s0];
    if (c == n_contigs) return;
    const int s1 = slotbase[c + 1];
    double t = 0.0;
    for (int s = s0; s < s1; s++) t += B.piece_total[s];
    B.contig_total[c] = t;
}

// ------------------------------------------------------------------------------------------------
// window sums and labels (reference bin/pipeline.py:157-385); thread per (window, model) -> warp per window batch
// ------------------------------------------------------------------------------------------------
// cents[w*M + m] = sum over genes of model m on the contig with start <= gene.left(1-based) < end; if that is 0 and the
// last gene starting before the window ends after it, that gene's value (pipeline.py:290-321).
__global__ void __launch_bounds__(128) k_windows(const mgc_gene *const *genes_m, const uint32_t *const *gene_off_m, int n_models,
                                                 const int32_t *__restrict__ clen, const int32_t *__restrict__ stride,
                                                 const int64_t *__restrict__ win_off, int32_t n_contigs, int32_t window,
                                                 int32_t consecutive, int short_contigs, int32_t *labels, long long *cents_out,
                                                 unsigned long long *label_hist /* [MGC_MAX_MODELS] or null */) {
    const int64_t c = (int64_t)blockIdx.x;
    if (c >= n_contigs) return;
    const int L = clen[c], st = stride[c];
    const int64_t wbase = win_off[c];
    const int nwin = (int)(win_off[c + 1] - wbase);
    // rule 1: contigs of at most 5000 bp are not classified unless --short_contigs (pipeline.py:196-225)
    const bool unclassified = L <= 5000 && !short_contigs;
    // rule 2: contigs too short for `consecutive` windows get the whole-contig argmax replicated (pipeline.py:227-275)
    const bool whole = !unclassified && (long long)L <= (long long)window + (long long)(consecutive - 1) * st;
    __shared__ long long tot[MGC_MAX_MODELS];
    __shared__ unsigned int hist[MGC_MAX_MODELS];
    if (threadIdx.x < MGC_MAX_MODELS) { tot[threadIdx.x] = 0; hist[threadIdx.x] = 0; }
    __syncthreads();
    if (whole) {
        for (int m = 0; m < n_models; m++) {
            const uint32_t a = gene_off_m[m][c], b = gene_off_m[m][c + 1];
            long long s = 0;
            for (uint32_t k = a + threadIdx.x; k < b; k += blockDim.x) s += llrint(genes_m[m][k].logodd * 100.0);
            if (s) atomicAdd((unsigned long long *)&tot[m], (unsigned long long)s);
        }
        __syncthreads();
    }
    for (int w = threadIdx.x; w < nwin; w += blockDim.x) {
        long long best = 0;
        int lab = 0;
        for (int m = 0; m < n_models; m++) {
            long long s = 0;
            if (unclassified) s = 0;
            else if (whole) s = tot[m];
            else {
                const long long start = (long long)w * st, end = start + window;
                const mgc_gene *g = genes_m[m];
                const uint32_t a = gene_off_m[m][c], b = gene_off_m[m][c + 1];
                // first gene with left >= start (genes ascend in left)
                uint32_t lo = a, hi = b;
                while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (g[mid].left < start) lo = mid + 1; else hi = mid; }
                for (uint32_t k = lo; k < b && g[k].left < end; k++) s += llrint(g[k].logodd * 100.0);
                if (s == 0 && lo > a && g[lo - 1].right > end) s = llrint(g[lo - 1].logodd * 100.0);
            }
            cents_out[(wbase + w) * n_models + m] = s;
            if (m == 0 || s > best) { best = s; lab = m; }
        }
        labels[wbase + w] = unclassified ? 0 : lab;
        if (label_hist) atomicAdd(&hist[unclassified ? 0 : lab], 1u);
    }
    if (label_hist) {   // labels of the whole input: the smoothing treats all but the two most frequent as noise
        __syncthreads();
        if (threadIdx.x < MGC_MAX_MODELS && hist[threadIdx.x]) atomicAdd(&label_hist[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
    }
}

// noise labels of Pipeline.find_best_sequence_of_labels (bin/pipeline.py:1084-1086): the labels that occur, ordered as
// np.argsort(counts)[::-1] orders them (ascending insertion sort is stable, so after the reversal equal counts put the
// higher label first), minus the first two
__global__ void k_noise_mask(const unsigned long long *label_hist, int n_models, uint32_t *mask_out) {
    if (threadIdx.x || blockIdx.x) return;
    int order[MGC_MAX_MODELS], n = 0;
    for (int l = 0; l < n_models; l++) {
        if (!label_hist[l]) continue;
        int k = n++;
        while (k > 0 && (label_hist[order[k - 1]] < label_hist[l] || (label_hist[order[k - 1]] == label_hist[l] && order[k - 1] < l))) {
            order[k] = order[k - 1];
            k--;
        }
        order[k] = l;
    }
    uint32_t mask = 0;
    for (int k = 2; k < n; k++) mask |= 1u << order[k];
    *mask_out = mask;
}

// ------------------------------------------------------------------------------------------------
// label smoothing (reference Pipeline.find_best_sequence_of_labels, bin/pipeline.py:1064-1168): thread per contig.
// Sequential per contig and tiny (one label per window); runs on the device so that labels never leave HBM between
// the window kernel and the segmentation inputs.  status[c] = 1 where the reference would raise IndexError (it walks
// off the array when a contig starts with a noise label).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_merge_labels(const int32_t *__restrict__ labels, const int64_t *__restrict__ win_off,
                                                     int32_t n_contigs, int32_t consecutive, uint32_t noise_mask_value,
                                                     const uint32_t *__restrict__ noise_mask_dev, int32_t *out, int32_t *run,
                                                     int32_t *status) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    const uint32_t noise_mask = noise_mask_dev ? *noise_mask_dev : noise_mask_value;
    const int64_t a = win_off[c];
    const int n = (int)(win_off[c + 1] - a);
    const int32_t *lab = labels + a;
    int32_t *o = out + a, *r = run + a;
    status[c] = 0;
    bool uniform = true;
    for (int i = 0; i < n; i++) {
        o[i] = lab[i];
        uniform &= lab[i] == lab[0];
    }
    if (n == 0 || uniform) return;
    auto noise = [&](int l) { return l >= 0 && l < 32 && ((noise_mask >> l) & 1u); };
    r[0] = 1;
    for (int i = 1; i < n; i++) r[i] = (lab[i] == lab[i - 1] || noise(lab[i])) ? r[i - 1] + 1 : 1;
    const int k = consecutive;
    while (true) {
        int seed = -1, best = 0;
        for (int i = 0; i < n; i++)
            if (r[i] > best) { best = r[i]; seed = i; }   // first maximum, as np.argmax
        if (seed < 0) break;
        int lo = seed - best, hi = seed + 1;
        int ref_pos = seed;
        int ref = o[ref_pos];
        while (noise(ref)) {               // a noise label never names a block
            ref_pos--;
            if (ref_pos < -n) { status[c] = 1; return; }
            ref = o[ref_pos < 0 ? ref_pos + n : ref_pos];   // negative indices wrap, as in numpy
        }
        int miss = 0;
        while (lo >= 0 && miss < k) { miss = o[lo] != ref ? miss + 1 : 0; lo--; }
        const bool left_exhausted = lo < 0;
        miss = 0;
        while (hi <= n - 1 && miss < k) { miss = o[hi] != ref ? miss + 1 : 0; hi++; }
        lo = left_exhausted ? 0 : lo + k + 1;
        hi = hi > n - 1 ? n : hi - k;
        if (hi - lo < k) {
            const int li = lo - 1;
            const int left_label = o[li < 0 ? li + n : li];
            const int right_label = hi < n ? o[hi] : left_label;
            if (left_label == right_label) ref = left_label;
            else {
                int i = lo - 1, left_len = 0;
                while (i >= 0 && o[i] == left_label) { i--; left_len++; }
                int j = hi, right_len = 0;
                while (j < n && o[j] == right_label) { j++; right_len++; }
                ref = left_len > right_len ? left_label : right_label;
            }
        }
        for (int i = max(lo, 0); i < hi; i++) { r[i] = 0; o[i] = ref; }
        if (hi <= lo) { status[c] = 2; return; }  // no progress (cannot happen for well-formed input)
    }
}

void launch_merge_labels(const int32_t *labels, const int64_t *win_off, int32_t n_contigs, int32_t consecutive, uint32_t noise_mask,
                         const uint32_t *noise_mask_dev, int32_t *out, int32_t *run, int32_t *status, cudaStream_t st) {
    k_merge_labels<<<(unsigned)((n_contigs + 63) / 64), 64, 0, st>>>(labels, win_off, n_contigs, consecutive, noise_mask, noise_mask_dev,
                                                                     out, run, status);
}

void launch_noise_mask(const unsigned long long *label_hist, int n_models, uint32_t *mask_out, cudaStream_t st) {
    k_noise_mask<<<1, 32, 0, st>>>(label_hist, n_models, mask_out);
}

// ------------------------------------------------------------------------------------------------
// in-frame stop codons inside genes (reference bin/results.py:253-294, Results.find_recoded_stops / make_seqrecord):
// codon indices k < ncodons - 1 whose codon is TAA / TAG / TGA, per gene, ascending.  One warp per gene over the
// packed planes of the resident batch.  FILL = false counts, FILL = true writes the indices at pos_off[g*3 + t].
// ------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(128) k_recoded_stops(Planes pl, const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                                       const GeneRef *__restrict__ genes, int64_t n_genes, int32_t *counts,
                                                       const int64_t *__restrict__ pos_off, int32_t *pos_out) {
    const int lane = lane_id();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp; g < n_genes; g += nw) {
        const GeneRef gr = genes[g];
        const FindCtx cx = make_ctx(pl, gr.strand, cwoff[gr.slot], clen[gr.slot], 7, 0);
        const int s = gr.start, last = gr.start + gr.n - 4;  // codon starts p with s <= p <= last, p = s mod 3: all but the final codon
        int base[3] = {0, 0, 0};
        if (last >= s) {
            const int wa = s >> 5, wb = last >> 5;
            for (int wblk = wa; wblk <= wb; wblk += 32) {
                const int w = wblk + lane;
                uint32_t bits[3] = {0u, 0u, 0u};
                if (w <= wb) {
                    const CodonMasks m = masks_at(cx, w);
                    const int r = ((s - 32 * w) % 3 + 3) % 3;
                    const uint32_t sel = frame_pattern(r) & bit_range(s - 32 * w, last - 32 * w);
                    bits[0] = m.taa & sel; bits[1] = m.tag & sel; bits[2] = m.tga & sel;
                }
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const int c = __popc(bits[t]);
                    int inc = c;  // inclusive scan over lanes
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += v;
                    }
                    if (FILL) {
                        int64_t at = pos_off[g * 3 + t] + base[t] + inc - c;
                        uint32_t b = bits[t];
                        while (b) {
                            const int bit = __ffs(b) - 1;
                            b &= b - 1;
                            pos_out[at++] = (32 * w + bit - s) / 3;
                        }
                    }
                    base[t] += __shfl_sync(0xffffffffu, inc, 31);
                }
            }
        }
        if (!FILL && lane == 0) { counts[g * 3] = base[0]; counts[g * 3 + 1] = base[1]; counts[g * 3 + 2] = base[2]; }
    }
}

void launch_recoded_stops(Planes pl, const int64_t *cwoff, const int32_t *clen, const GeneRef *genes, int64_t n_genes, int32_t *counts,
                          const int64_t *pos_off, int32_t *pos_out, bool fill, int n_sm, cudaStream_t st) {
    int64_t blocks = (n_genes + 3) / 4;
    if (blocks > (int64_t)n_sm * 16) blocks = (int64_t)n_sm * 16;
    if (blocks < 1) blocks = 1;
    if (fill) k_recoded_stops<true><<<(unsigned)blocks, 128, 0, st>>>(pl, cwoff, clen, genes, n_genes, counts, pos_off, pos_out);
    else k_recoded_stops<false><<<(unsigned)blocks, 128, 0, st>>>(pl, cwoff, clen, genes, n_genes, counts, pos_off, pos_out);
}

// ------------------------------------------------------------------------------------------------
// runs of >= min_run ambiguous letters per contig (only launched when the batch has ambiguous letters)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_find_gaps(const uint32_t *__restrict__ nm, const int64_t *__restrict__ cwoff,
                                                   const int32_t *__restrict__ clen, int32_t n_contigs, int min_run, int32_t *gaps,
                                                   unsigned int *gap_count, unsigned int cap) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    const uint32_t *m = nm + cwoff[c];
    const int L = clen[c];
    int run = 0;
    auto close = [&](int end) {  // a run of `run` letters ended just before position `end`
        if (run >= min_run) {
            const unsigned int k = atomicAdd(gap_count, 1u);
            if (k < cap) { gaps[k * 3] = c; gaps[k * 3 + 1] = end - run; gaps[k * 3 + 2] = run; }
        }
        run = 0;
    };
    const int nwords = (L + 31) >> 5;
    for (int w = 0; w < nwords; w++) {
        const uint32_t x = m[w];
        const int nb = min(32, L - 32 * w);
        if (x == 0u) { if (run) close(32 * w); continue; }
        if (nb == 32 && x == 0xffffffffu) { run += 32; continue; }
        for (int b = 0; b < nb; b++) {
            if ((x >> b) & 1u) run++;
            else if (run) close(32 * w + b);
        }
    }
    if (run) close(L);
}

void launch_find_gaps(const uint32_t *nm, const int64_t *cwoff, const int32_t *clen, int32_t n_contigs, int min_run, int32_t *gaps,
                      unsigned int *gap_count, unsigned int cap, cudaStream_t st) {
    k_find_gaps<<<(unsigned)((n_contigs + 127) / 128), 128, 0, st>>>(nm, cwoff, clen, n_contigs, min_run, gaps, gap_count, cap);
}

// ------------------------------------------------------------------------------------------------
// host-callable launch wrappers (keep <<< >>> in this translation unit)
// ------------------------------------------------------------------------------------------------
void launch_groups(const int64_t *cwoff, int32_t n_contigs, int64_t n_groups, int32_t *grp2contig, cudaStream_t st) {
    k_groups<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(cwoff, n_contigs, n_groups, grp2contig);
}

void launch_pack(const char *ascii, const int64_t *aoff, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig,
                 int64_t W, uint32_t *lo, uint32_t *hi, uint32_t *nm, uint32_t *gccount, uint32_t *ncount, int *dflags,
                 cudaStream_t st) {
    dim3 grid((unsigned)((W + 255) / 256), 2);
    k_pack<<<grid, 256, 0, st>>>(ascii, aoff, cwoff, clen, grp2contig, W, lo, hi, nm, gccount, ncount, dflags);
}

void launch_codon_planes(Planes pl, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig, uint32_t *codon, cudaStream_t st) {
    k_codon_planes<<<dim3((unsigned)((pl.W + 255) / 256), 2), 256, 0, st>>>(pl, cwoff, clen, grp2contig, codon);
}

int launch_laststop(Planes pl, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig, const DevModel *models,
                    int n_models, uint32_t *const *ls_m, uint32_t *const *ls_tile_m, cudaStream_t st) {
    const int64_t nt = (2 * pl.W + LS_THREADS - 1) / LS_THREADS;
    dim3 grid((unsigned)nt, n_models);
    k_laststop_tiles<<<grid, LS_THREADS, 0, st>>>(pl, cwoff, clen, grp2contig, models, ls_m, ls_tile_m);
    k_laststop_sums<<<n_models, 1024, 0, st>>>(ls_tile_m, nt);
    return 2;  // the carry into each tile is applied by the consumer (k_orf_find)
}

void launch_orf_find(Planes pl, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig, const DevModel *models,
                     int n_models, const uint32_t *const *ls_m, const uint32_t *const *ls_tile_m, OrfRec *const *orfs_m,
                     unsigned long long *counters_m, uint32_t cap_orf, cudaStream_t st) {
    dim3 grid((unsigned)((2 * pl.W + 127) / 128), n_models);
    k_orf_find<<<grid, 128, 0, st>>>(pl, cwoff, clen, grp2contig, models, ls_m, ls_tile_m, orfs_m, counters_m, cap_orf);
}

int launch_orf_prep(Planes pl, const uint32_t *gcpre, const int64_t *cwoff, const int32_t *clen, const DevModel *models, int n_models,
                    OrfRec *const *orfs_m, StartRec *const *starts_m, ItemRec *const *items_m, uint32_t *const *hist_m,
                    unsigned long long *counters_m, const ChainBufs *bufs_m, uint32_t cap_orf, uint32_t cap_starts, uint32_t cap_items,
                    int n_sm, cudaStream_t st) {
    k_orf_expand<<<dim3(n_sm * 8, n_models), EXPAND_ORFS, 0, st>>>(pl, gcpre, cwoff, clen, models, orfs_m, starts_m, counters_m, bufs_m,
                                                                 cap_orf, cap_starts);
    k_start_link<<<dim3(n_sm * 8, n_models), 256, 0, st>>>(pl, cwoff, clen, models, orfs_m, starts_m, items_m, hist_m, counters_m, cap_starts,
                                                         cap_items);
    return 2;
}

int launch_item_sort(const ItemRec *const *items_m, ItemRec *const *sorted_m, uint32_t *const *hist_m, uint32_t *const *item_off_m,
                     uint32_t *const *unit_off_m, const DevModel *models, int n_models, uint32_t *unit_totals,
                     const unsigned long long *counters_m, uint32_t cap_items, int n_tiles, int n_sm, bool pair_walk, cudaStream_t st) {
    k_item_plan<<<n_models, 1024, 0, st>>>(hist_m, item_off_m, unit_off_m, models, unit_totals, const_cast<unsigned long long *>(counters_m), n_tiles,
                                           pair_walk ? (uint32_t)WALK2_CHUNK : (uint32_t)WALK_CHUNK);
    dim3 grid(n_sm * 4, n_models);
    k_item_scatter<<<grid, 256, 0, st>>>(items_m, sorted_m, hist_m, counters_m, cap_items);
    return 2;
}

int walk_smem_bytes(int cod_order, int non_order) { return (3 * (1 << (2 * (cod_order + 1))) * 2 + (1 << (2 * (non_order + 1))) * 8) * 16; }

int launch_walk(Planes pl, const int64_t *cwoff, const DevModel *models, int n_models, int max_cod_order, int max_non_order,
                const OrfRec *const *orfs_m,
                const StartRec *const *starts_m, const ItemRec *const *items_m, const uint32_t *const *item_off_m,
                const uint32_t *const *unit_off_m, const uint32_t *unit_totals, void *const *sums_m, unsigned int *next_unit, int n_sm,
                int pair_smem, bool pair_orders_uniform, cudaStream_t st) {
    const int smem = walk_smem_bytes(max_cod_order, max_non_order);
    Acc4 *const *sums = reinterpret_cast<Acc4 *const *>(sums_m);
    if (pair_smem > 0) {
        // pair tables (two positions per lookup): one CTA per SM with the whole shared memory
        if (max_cod_order == 4 && max_non_order == 2 && pair_orders_uniform) {
            cudaFuncSetAttribute(k_walk_pair<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WALK2_SMEM_MAX);
            k_walk_pair<4, 2><<<n_sm, WALK2_THREADS, pair_smem, st>>>(pl, models, n_models, starts_m, items_m, item_off_m, unit_off_m,
                                                                      unit_totals, sums, next_unit);
        } else {
            cudaFuncSetAttribute(k_walk_pair<-1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WALK2_SMEM_MAX);
            k_walk_pair<-1, -1><<<n_sm, WALK2_THREADS, pair_smem, st>>>(pl, models, n_models, starts_m, items_m, item_off_m, unit_off_m,
                                                                        unit_totals, sums, next_unit);
        }
    } else if (smem <= 110 * 1024) {
        cudaFuncSetAttribute(k_walk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        k_walk<true><<<n_sm * 2, WALK_UNIT, smem, st>>>(pl, cwoff, models, n_models, orfs_m, starts_m, items_m, item_off_m, unit_off_m,
                                                        unit_totals, sums, next_unit);
    } else {
        // tables of order > 4 do not fit next to a second CTA: gather from global memory (L1/L2) instead
        k_walk<false><<<n_sm * 4, WALK_UNIT, 0, st>>>(pl, cwoff, models, n_models, orfs_m, starts_m, items_m, item_off_m, unit_off_m,
                                                      unit_totals, sums, next_unit);
    }
    return 1;
}

void launch_final(Planes pl, const int64_t *cwoff, const int32_t *clen, const DevModel *models, int n_models, const OrfRec *const *orfs_m,
                  const StartRec *const *starts_m, void *const *sums_m, const unsigned long long *counters_m, Cand *const *surv_m,
                  uint32_t *const *bcount_m, uint32_t *const *cover_m, unsigned long long *surv_counters, uint32_t cap_starts, uint32_t cap_surv, int n_sm,
                  bool pair_walk, cudaStream_t st) {
    dim3 grid(n_sm * 8, n_models);
    k_final<<<grid, 256, 0, st>>>(pl, cwoff, clen, models, orfs_m, starts_m, reinterpret_cast<const Acc4 *const *>(sums_m), counters_m,
                                  surv_m, bcount_m, cover_m, surv_counters, cap_starts, cap_surv, pair_walk ? 1 : 0);
}

void launch_scatter(const Cand *const *surv_m, Cand *const *sorted_m, uint32_t *const *boff_m,
                    const unsigned long long *surv_counters, const int64_t *cwoff, uint32_t cap_surv, int n_models, int n_sm,
                    cudaStream_t st) {
    dim3 grid(n_sm * 4, n_models);
    k_scatter<<<grid, 256, 0, st>>>(surv_m, sorted_m, boff_m, surv_counters, cwoff, cap_surv);
}

// All per-decode clears in ONE launch.  cudaMemsetAsync runs on a copy engine: at the start of a step it queued behind the previous
// result's device-to-host copy on the copy-out stream and the whole step started 3 ms late (measured: back-to-back decodes
// 69.0 ms, 66.1 ms with a pause between them).  Regions are whole 4-byte words.
__global__ void __launch_bounds__(256) k_clear(ClearList list) {
    for (int r = 0; r < list.n; r++) {
        uint32_t *p = static_cast<uint32_t *>(list.ptr[r]);
        const uint32_t v = list.value[r];
        const size_t words = list.words[r];
        uint4 *p4 = reinterpret_cast<uint4 *>(p);   // cudaMalloc'ed bases: 16-byte aligned
        const size_t n4 = words / 4;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) p4[i] = make_uint4(v, v, v, v);
        for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
    }
}
void launch_clear(const ClearList &list, int n_sm, cudaStream_t st) {
    if (list.n) k_clear<<<(unsigned)n_sm * 4, 256, 0, st>>>(list);
}

// diagnostic (MGCOD_B200_PIECE_STATS): log2 histogram of candidates per chain piece of model 0; hist[32] = largest piece, hist[33] = pieces
__global__ void k_piece_stats(const ChainBufs *bufs_m, const int *work_counters, int n_slots, unsigned int *hist) {
    const ChainBufs B = bufs_m[0];
    const int n_work = work_count(work_counters);
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_work) return;
    const int slot = work_slot(B, work_counters, item, n_slots);
    const int64_t b_first = B.piece_b0[slot], b_last = B.piece_b1[slot];
    const uint32_t n = B.bend[b_last] - (b_first ? B.bend[b_first - 1] : 0u);
    atomicAdd(&hist[n ? 32 - __clz(n) : 0], 1u);
    atomicMax(&hist[32], n);
    atomicAdd(&hist[33], 1u);
}
void launch_piece_stats(const ChainBufs *bufs_m, const int *work_counters, int32_t n_slots, unsigned int *hist, cudaStream_t st) {
    k_piece_stats<<<(unsigned)((n_slots + 127) / 128), 128, 0, st>>>(bufs_m, work_counters, n_slots, hist);
}

int launch_chain(const ChainBufs *bufs_m, const Cand *const *scattered_m, Cand *const *ordered_m, const int64_t *cwoff, const int32_t *clen,
                 const int32_t *slotbase, const int32_t *slot_contig, int32_t n_slots, int64_t n_buckets, int *work_counters, int n_models,
                 int n_sm, int chain_mode, cudaStream_t st) {
    k_bucket_order<<<dim3((unsigned)n_sm * 8, n_models), 256, 0, st>>>(bufs_m, scattered_m, ordered_m, cwoff, n_buckets);
    k_chain_plan<<<dim3((unsigned)((n_slots + 127) / 128), n_models), 128, 0, st>>>(bufs_m, cwoff, clen, slotbase, slot_contig, n_slots,
                                                                                     work_counters);
    // grids sized so that the blocks of ALL models are resident together: a model's pieces then share the machine for the whole
    // kernel instead of four runs one after the other, each ending on its longest piece
    const unsigned per_model = (unsigned)std::max(1, n_sm * 8 / std::max(1, n_models));
    // Pieces of CHAIN_BIG_PIECE candidates and more (3 % of the pieces, 7 % of the candidates) go to the warp-per-piece kernel: a
    // thread needs ~9 us per candidate of its piece (a dozen dependent sections), so the longest pieces set the floor of
    // k_chain_flat — 3 ms of a 125 Mbp batch's 3.3 ms chain phase (the per-rank share of 1 Gbp over 8 GPUs); a warp does them in a
    // tenth of that.  mode 1: everything through the warp kernel, mode 2: everything through the thread kernel (measurements).
    if (chain_mode == 1) k_chain<<<dim3(per_model * 2, n_models), 128, 0, st>>>(bufs_m, work_counters, n_slots, 0);
    else if (chain_mode == 2) k_chain_flat<<<dim3(per_model, n_models), 128, 0, st>>>(bufs_m, work_counters, n_slots, 0);
    else {
        k_chain<<<dim3(per_model, n_models), 128, 0, st>>>(bufs_m, work_counters, n_slots, 1);
        k_chain_flat<<<dim3(per_model, n_models), 128, 0, st>>>(bufs_m, work_counters, n_slots, 1);
    }
    k_chain_trace<<<dim3((unsigned)((n_slots + 127) / 128), n_models), 128, 0, st>>>(bufs_m, work_counters, n_slots);
    return chain_mode == 0 ? 5 : 4;
}

int launch_gather(const ChainBufs *bufs_m, mgc_gene *const *genes_m, const int32_t *slot_contig, int32_t n_slots, const int32_t *slotbase,
                  int32_t n_contigs, long long *model_cents, int n_models, int n_sm, cudaStream_t st) {
    int64_t blocks = ((int64_t)n_slots + 3) / 4;
    if (blocks > (int64_t)n_sm * 16) blocks = (int64_t)n_sm * 16;
    k_gather<<<dim3((unsigned)blocks, n_models), 128, 0, st>>>(bufs_m, genes_m, slot_contig, n_slots, model_cents);
    k_contig_finish<<<dim3((unsigned)(n_contigs / 256 + 1), n_models), 256, 0, st>>>(bufs_m, slotbase, n_contigs);
    return 2;
}

void launch_windows(const mgc_gene *const *genes_m, const uint32_t *const *gene_off_m, int n_models, const int32_t *clen,
                    const int32_t *stride, const int64_t *win_off, int32_t n_contigs, int32_t window, int32_t consecutive,
                    int short_contigs, int32_t *labels, long long *cents, unsigned long long *label_hist, cudaStream_t st) {
    k_windows<<<(unsigned)n_contigs, 128, 0, st>>>(genes_m, gene_off_m, n_models, clen, stride, win_off, n_contigs, window,
                                                    consecutive, short_contigs, labels, cents, label_hist);
}

}  // namespace mgc
