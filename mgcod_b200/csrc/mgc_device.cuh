// mgc_device.cuh — device-side data layout and small bit-parallel helpers (sm_100a).
//
// Sequence layout in HBM ("bit planes"): every contig owns a slot of `slot_words` 32-bit words per
// plane (slot_words = ceil(L/32)+1 rounded up to a multiple of 4, so slots start on 16-byte
// boundaries and a trailing zero word always exists).  Letter code = 2*hi + lo with A=0 C=1 G=2 T=3;
// bit b of word w is the letter at contig position 32*w + b.  The reverse-complement strand is stored
// the same way in a second region (strand 1) at the same slot offset: rc position j = L-1-j forward.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mgc_model.hpp"

namespace mgc {

// four per-position numbers of one (bacterial, archaeal) section pair: coding log-likelihood and
// coding - non-coding log-odds of each domain; one 32-byte sector per gather
struct __align__(32) Quad {
    double cB, oB, cA, oA;
};

struct DevModel {
    int gcode, kc, kn, min_len, nsec, kr, knr;
    int stop_sel;   // bit0 TAA, bit1 TAG, bit2 TGA are stops under this genetic code
    int atg_only;   // genetic code 1
    int nfull, nnon, nramp, nnramp;
    const double *lc_full, *ln_full, *lc_ramp, *ln_ramp;
    const SectionScalars *sc;
    const Quad *quad;   // [npair][3][4^(kc+1)], plane-major k-mer index
    const double2 *wtab; // [npair][wtab_n16]: the single-position walk table image (coding part 2x replicated, non-coding part 8x replicated)
    int wtab_n16;
    const double2 *wtab2; // [npair][wtab2_n16]: the pair walk's table image: coding {B, A} sums of two consecutive positions per
                          // (phase, (kc+2)-mer), rc2 replicas side by side; non-coding sums per (kn+2)-mer, 8 replicas (k_walk_pair)
    int wtab2_n16;
    int rc2;              // replicas of the coding part (power of two, 1 for order 4); 0 = the tables do not fit one SM
    const Quad *rquad;  // [npair][kr][4^kr]: order ramp (window starts on a codon boundary, phase 0)
    int npair;
    int map[2][101];
    int pair[101];      // gc -> index of its (secB, secA) pair in quad
    const double *pair_amb;  // [npair][4]: ln COD_P_N (B, A), ln NON_P_N (B, A) — terms of a window with an ambiguous letter
    const double *pair_ub;   // [npair][8]: upper bounds per domain of the ramp, start and stop terms ([dom*3 + 0..2]) and of one
                             // full-order position ([6 + dom], pair walk only) — early pruning in k_final
};

// one stop-delimited ORF that may have start candidates of sufficient length
struct __align__(8) OrfRec {
    int32_t contig;
    int32_t t;          // strand-local position of the first nt of the (real or virtual) stop codon
    int32_t s_lo;       // first position a start codon may occupy (previous in-frame stop + 3, or the frame's first codon)
    uint32_t base;      // first entry in the starts array (filled by k_orf_prep)
    uint32_t n;         // number of starts              (filled by k_orf_prep)
    uint32_t sf;        // bit0: strand (1 = reverse complement); bit1: open at the left edge (virtual start);
                        // bit8: stop codon missing (ORF runs into the right edge); bit9: the virtual start is summed from
                        // its first letter (every case except a reverse-strand ORF cut by the true end of the record)
};

// one start candidate of an ORF
struct __align__(16) StartRec {
    uint32_t sraw;      // strand-local position of the start codon; bit31: virtual start at the contig edge
    uint32_t orf;       // index of its OrfRec
    uint32_t pg;        // section pair (low 16 bits) | gc << 16
    int32_t link;       // index (within the ORF) of the previous start of the same section pair that has a full-order region, -1 none
};

// one 32-letter word of an ORF's admissible start range

// one walk of an ORF under one section pair: serves every start of that pair
// (32 bytes: the second half repeats what the walk needs of the ORF record, so that a walk starts with its planes and start
// records one load after the coalesced item load instead of item -> ORF -> contig offset -> planes)
struct __align__(16) ItemRec {
    uint32_t orf;       // bit 31: the walk touches an ambiguous letter
    int32_t lo_pos;     // lowest position of the walk (first full-order position of the farthest start of this pair)
    uint32_t key;       // (tile * npair + pair) * WALK_CLASSES + length class (descending)
    int32_t last;       // index (within the ORF) of the start of this pair closest to the stop
    int32_t t;          // OrfRec.t
    uint32_t base;      // OrfRec.base
    uint32_t woff;      // word offset of the contig's slot inside the planes of the ORF's strand (strand * W + cwoff[contig])
    uint32_t sf;        // OrfRec.sf
};
static_assert(sizeof(ItemRec) == 32, "ItemRec must be 32 bytes");

// Walk items are ordered by (position tile, section pair, length class): all items of a tile are walked before the next
// tile, so the plane words of a tile (both strands, lo + hi = 16 B per word) stay in the 126 MB L2 while every section
// pair and length class passes over them.  A tile is WALK_TILE_WORDS words (32 MB of planes; 16 tiles for 1 Gbp) unless that would give more
// than WALK_TILES_MAX tiles; a small input is one tile.
constexpr int WALK_TILES_MAX = 32;
constexpr long long WALK_TILE_WORDS = 2ll << 20;
struct WalkTiles {
    int n;              // tiles in this batch
    long long words;    // words per tile
};
inline WalkTiles walk_tiles_for(long long W, long long tile_words = WALK_TILE_WORDS) {
    WalkTiles t;
    if (tile_words < 1) tile_words = WALK_TILE_WORDS;
    long long n = (W + tile_words - 1) / tile_words;
    if (n < 1) n = 1;
    if (n > WALK_TILES_MAX) n = WALK_TILES_MAX;
    t.n = (int)n;
    t.words = (W + n - 1) / n;
    if (t.words < 1) t.words = 1;
    return t;
}
constexpr int WALK_CLASSES = 64;   // length classes of 32 nt; the last one collects everything longer
constexpr int WALK_UNIT = 256;     // threads per CTA of k_walk
constexpr int WALK_CHUNK = 2048;   // items per work unit (one section pair; warps take 32 at a time)
// pair walk: one CTA per SM owns the whole shared memory (192 KB coding + 32 KB non-coding for orders 4/2)
constexpr int WALK2_THREADS = 768;
constexpr int WALK2_CHUNK = 8192;  // larger units: the tail of a unit (its shortest items) idles a smaller share of the CTA
constexpr int WALK2_SMEM_MAX = 227 * 1024 - 256;

// a start/stop pair with total > 0, forward coordinates, 0-based inclusive
struct __align__(16) Cand {
    int32_t contig;
    int32_t left, right;
    uint32_t orf;       // id of the ORF record (groups reverse-strand starts sharing a stop)
    double total;
    int16_t gc;
    int8_t strand;      // +1 / -1
    uint8_t flags;      // bit0 domain (1 = archaea), bit1 partial start, bit2 partial stop
    int32_t len;
};
static_assert(sizeof(Cand) == 32, "Cand must be 32 bytes");

constexpr int32_t GPTR_UNSET = (int32_t)0x80808080;

// predecessor state of the thread-per-piece chain
struct __align__(16) ChainSt {
    int4 i;      // len * 2 + (strand < 0), sorted index, sorted index of the prefix maximum, unused
    double2 d;   // cum, prefix maximum of cum
};

struct ChainBufs {
    const Cand *sorted;
    const uint32_t *bend;   // bucket ends (bucket b spans [b ? bend[b-1] : 0, bend[b]))
    const uint32_t *cover;  // exclusive prefix of the boundary difference array: cover[k + 1] = LONG candidates spanning the boundary in front of bucket k
    uint32_t *minleft;      // per bucket: smallest left end among the candidates ending in it (0xffffffff if none)
    double *cum;            // per sorted entry
    int32_t *prev;          // per sorted entry (index into sorted, -1 none)
    int32_t *gptr;          // per ORF record: shared back-pointer of reverse-strand ORFs (GPTR_UNSET until first use)
    // available predecessor states, appended in right-end order, stored at the contig's sorted range
    int32_t *av_right, *av_left, *av_ls;  // av_ls = len * 2 + (strand < 0)
    int32_t *av_idx;
    double *av_cum, *av_pmax;
    int32_t *av_parg;
    ChainSt *st;            // thread-per-piece chain (k_chain_flat): one record per state
    int2 *st_rl;            //   and its {right, left}, what the scan over predecessor states reads
    int32_t *chain;         // traceback output, piece-local, stored at the piece's sorted range
    // chain pieces: a contig is cut where no candidate spans a bucket boundary; pieces are independent chains.
    // Slot of the piece starting at relative bucket rb of contig c: slotbase[c] + rb / PIECE_BUCKETS.
    int32_t *piece_b0, *piece_b1;   // per slot: first / last bucket (absolute index)
    int32_t *work;                  // slots in processing order (k_chain_plan)
    int32_t *gene_count;            // per slot (+1 terminator for the scan)
    double *piece_total;            // per slot
    const uint32_t *slot_gene_off;  // exclusive scan of gene_count
    uint32_t *gene_off;             // per contig (+1)
    double *contig_total;           // per contig
};
// a gene located in the packed planes: slot, strand region (0 forward, 1 reverse complement), first letter and length there
struct __align__(16) GeneRef {
    int32_t slot, strand, start, n;
};
constexpr int SURV_CHUNK = 256;     // survivor slots a warp of k_final reserves at a time
constexpr int PIECE_BUCKETS = 64;
constexpr int PACK_SLACK = 64;      // bytes of slack before and after the letters in the device ASCII buffer (k_pack)   // cuts are at least this many 128-nt buckets apart

struct Planes {
    const uint32_t *lo, *hi;  // [2][W + 4]
    const uint32_t *nm;       // [2][W + 4] ambiguity mask: bit set = letter outside ACGT (its code in lo/hi is 0)
    const uint32_t *npre;     // [W + 4] exclusive prefix of ambiguous letters per word, forward region
    const uint32_t *codon;    // [5][2W] codon planes: start (ATG|GTG|TTG), ATG, TAA, TAG, TGA (k_codon_planes)
    const int *dflags;        // device flags written by k_pack: [0] letter outside the IUPAC alphabet, [1] any ambiguous letter
    const uint8_t *cflags;    // per contig slot: bit0 = its right edge is the end of the record (not a gap)
    int64_t W;                // words per strand region
    int n_tiles;              // walk_tiles_for(W): position tiles of the walk-item order
    long long tile_words;
};

__device__ __forceinline__ uint32_t frame_pattern(int r) {
    // bits b with b % 3 == r
    return 0x49249249u << r;   // r in {0, 1, 2}: 0x49249249, 0x92492492, 0x24924924
}

// letters of 32 consecutive positions starting at word w (needs w+1 for the two look-ahead letters)
struct CodonMasks {
    uint32_t start, taa, tag, tga;
};

__device__ __forceinline__ CodonMasks codon_masks(uint32_t l0, uint32_t h0, uint32_t l1n, uint32_t h1n, int atg_only) {
    uint32_t l1 = __funnelshift_r(l0, l1n, 1), h1 = __funnelshift_r(h0, h1n, 1);
    uint32_t l2 = __funnelshift_r(l0, l1n, 2), h2 = __funnelshift_r(h0, h1n, 2);
    uint32_t T0 = h0 & l0, A0 = ~h0 & ~l0, C0 = ~h0 & l0;
    uint32_t T1 = h1 & l1, A1 = ~h1 & ~l1, G1 = h1 & ~l1;
    uint32_t A2 = ~h2 & ~l2, G2 = h2 & ~l2;
    CodonMasks m;
    m.start = T1 & G2 & (atg_only ? A0 : ~C0);
    m.taa = T0 & A1 & A2;
    m.tag = T0 & A1 & G2;
    m.tga = T0 & G1 & A2;
    return m;
}

__device__ __forceinline__ uint32_t stop_mask(const CodonMasks &m, int stop_sel) {
    return ((stop_sel & 1) ? m.taa : 0u) | ((stop_sel & 2) ? m.tag : 0u) | ((stop_sel & 4) ? m.tga : 0u);
}

// bits [from, to] (inclusive, 0..31) set
__device__ __forceinline__ uint32_t bit_range(int from, int to) {
    if (from > to) return 0u;
    uint32_t hi = to >= 31 ? 0xffffffffu : ((1u << (to + 1)) - 1u);
    uint32_t lo = from <= 0 ? 0u : ((1u << from) - 1u);
    return hi & ~lo;
}

}  // namespace mgc
