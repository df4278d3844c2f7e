// mgc_kernels.cu — the CUDA path (sm_100a) of the gene finder.
//
// Pipeline for one batch of contigs and M parameter files (all launches on one stream, no host
// synchronisation until the final counters are read):
//
//   k_groups      word-group -> contig map                                     (once per upload)
//   k_pack        ASCII -> 2-bit planes for both strands, GC count per word    (reads 1 B/nt twice)
//   scan          exclusive prefix of GC counts
//   k_orf_find    per model: stop-delimited ORFs and their start codons        (bit-parallel on planes)
//   k_orf_score   per model: one warp per ORF; 3-periodic Markov log-likelihoods of every start as
//                 warp suffix scans, GC-bin section choice, domain choice, duration/start/stop terms,
//                 pruning (total > 0), bucket histogram of right ends
//   scan + k_scatter   survivors into position buckets (128 nt)  -> right-end order per contig
//   k_chain       per (contig, model): one warp; max-plus chain with overlap penalties, traceback
//   scan + k_gather    genes ordered by (contig, left), per-model sums in 1/100 units
//   k_windows     per (contig, window): sums, spanning-gene fallback, argmax label
//
// Arithmetic follows SURVEY.md §A.4-A.6 as pinned by oracle/ (never linked here).
#include <cstdint>
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

// number of G/C letters in strand-0 positions [a, b] of contig slot starting at word `w0`
__device__ __forceinline__ int gc_count_fwd(const uint32_t *lo, const uint32_t *hi, const uint32_t *gcpre, int64_t w0,
                                            int a, int b) {
    // pre(x) = number of GC among positions < x
    int64_t wa = w0 + (a >> 5), wb = w0 + ((b + 1) >> 5);
    uint32_t ma = (1u << (a & 31)) - 1u, mb = (1u << ((b + 1) & 31)) - 1u;
    int pa = (int)gcpre[wa] + __popc((lo[wa] ^ hi[wa]) & ma);
    int pb = (int)gcpre[wb] + __popc((lo[wb] ^ hi[wb]) & mb);
    return pb - pa;
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

// one warp packs 32 words (1024 letters) of one strand: lane = bit position, ballot = plane word
__global__ void __launch_bounds__(256) k_pack(const char *__restrict__ ascii, const int64_t *__restrict__ aoff,
                                              const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                              const int32_t *__restrict__ grp2contig, int64_t W, uint32_t *lo,
                                              uint32_t *hi, uint32_t *gccount, int *bad_letter) {
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = lane_id();
    int strand = blockIdx.y;
    int64_t w0 = warp * 32;
    if (w0 >= W) return;
    uint32_t mylo = 0, myhi = 0, mygc = 0;
    int bad = 0;
    for (int r = 0; r < 32; r++) {
        int64_t w = w0 + r;
        uint32_t code = 0;
        bool valid = false;
        if (w < W) {
            int c = grp2contig[w >> 2];
            int L = clen[c];
            int64_t i = (w - cwoff[c]) * 32 + lane;
            if (i < L) {
                valid = true;
                unsigned char ch = (unsigned char)ascii[aoff[c] + (strand ? (L - 1 - i) : i)];
                unsigned char up = ch & 0xDF;  // upper case
                if (!(up == 'A' || up == 'C' || up == 'G' || up == 'T' || up == 'U')) bad = 1;
                code = (ch >> 1) & 3u;         // A0 C1 G3 T2 (also U -> 2)
                code ^= code >> 1;             // A0 C1 G2 T3
                if (strand) code = 3u - code;
            }
        }
        uint32_t bl = __ballot_sync(0xffffffffu, code & 1u);
        uint32_t bh = __ballot_sync(0xffffffffu, (code >> 1) & 1u);
        uint32_t bv = __ballot_sync(0xffffffffu, valid);
        if (lane == r) {
            mylo = bl;
            myhi = bh;
            mygc = __popc((bl ^ bh) & bv);
        }
    }
    if (w0 + lane < W) {
        lo[(int64_t)strand * W + w0 + lane] = mylo;
        hi[(int64_t)strand * W + w0 + lane] = myhi;
        if (!strand) gccount[w0 + lane] = mygc;
    }
    if (!strand && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(bad_letter, 1);
}

// ------------------------------------------------------------------------------------------------
// ORF discovery
// ------------------------------------------------------------------------------------------------
struct FindCtx {
    const uint32_t *lo, *hi;  // strand region base + contig slot
    int L;                    // contig length
    int nwords;               // ceil(L/32)
    int stop_sel, atg_only;
};

__device__ __forceinline__ CodonMasks masks_at(const FindCtx &cx, int w) {
    CodonMasks m = codon_masks(cx.lo[w], cx.hi[w], cx.lo[w + 1], cx.hi[w + 1], cx.atg_only);
    // codon start positions must satisfy i <= L-3
    int last = cx.L - 3 - 32 * w;
    uint32_t valid = last < 0 ? 0u : (last >= 31 ? 0xffffffffu : ((1u << (last + 1)) - 1u));
    m.start &= valid; m.taa &= valid; m.tag &= valid; m.tga &= valid;
    return m;
}

// Handles the ORF that ends at stop position t (first nt of the stop codon, or of the last full codon
// when `virt`): finds the previous in-frame stop, counts / writes the start candidates.
__device__ void handle_stop(const FindCtx &cx, const DevModel &dm, int contig, int strand, int t, bool virt,
                            OrfRec *orfs, uint32_t *starts, unsigned long long *counters, uint32_t cap_orf,
                            uint32_t cap_starts) {
    const int f = t % 3;
    // latest start position that still gives len >= min_len and at least one codon between start and stop
    int s_hi = min(t - 3, t + 3 - dm.min_len);
    if (s_hi < f) return;
    // previous in-frame stop
    int p_prev = -1;
    {
        int w = t >> 5;
        uint32_t below = (t & 31) ? ((1u << (t & 31)) - 1u) : 0u;
        for (; w >= 0; w--) {
            CodonMasks m = masks_at(cx, w);
            int r = ((f - (w * 32) % 3) % 3 + 3) % 3;
            uint32_t sm = stop_mask(m, cx.stop_sel) & frame_pattern(r) & below;
            if (sm) { p_prev = w * 32 + 31 - __clz(sm); break; }
            below = 0xffffffffu;
        }
    }
    const bool open_left = p_prev < 0;
    int s_lo = open_left ? f : p_prev + 3;
    if (s_lo > s_hi) return;
    // pass 1: count
    uint32_t n = open_left ? 1u : 0u;
    for (int w = s_lo >> 5; w <= (s_hi >> 5); w++) {
        CodonMasks m = masks_at(cx, w);
        int r = ((f - (w * 32) % 3) % 3 + 3) % 3;
        uint32_t mm = m.start & frame_pattern(r) & bit_range(s_lo - 32 * w, s_hi - 32 * w);
        n += __popc(mm);
    }
    if (!n) return;
    unsigned long long oid = atomicAdd(&counters[0], 1ull);
    unsigned long long base = atomicAdd(&counters[1], (unsigned long long)n);
    if (oid >= cap_orf || base + n > cap_starts) return;  // overflow is detected on the host from the counters
    OrfRec rec;
    rec.contig = contig;
    rec.t = t;
    rec.base = (uint32_t)base;
    rec.n = n;
    rec.sf = (uint32_t)strand | (virt ? 0x100u : 0u);
    orfs[oid] = rec;
    // pass 2: write (virtual start first; bit31 marks it)
    uint32_t k = (uint32_t)base;
    if (open_left) starts[k++] = 0x80000000u | (uint32_t)f;
    for (int w = s_lo >> 5; w <= (s_hi >> 5); w++) {
        CodonMasks m = masks_at(cx, w);
        int r = ((f - (w * 32) % 3) % 3 + 3) % 3;
        uint32_t mm = m.start & frame_pattern(r) & bit_range(s_lo - 32 * w, s_hi - 32 * w);
        while (mm) {
            int b = __ffs(mm) - 1;
            mm &= mm - 1;
            starts[k++] = (uint32_t)(w * 32 + b);
        }
    }
}

// thread per word of the slot space; grid.y = strand, grid.z = model
__global__ void __launch_bounds__(128) k_orf_find(Planes pl, const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                                  const int32_t *__restrict__ grp2contig, const DevModel *__restrict__ models,
                                                  OrfRec *const *orfs_m, uint32_t *const *starts_m,
                                                  unsigned long long *counters_m, uint32_t cap_orf, uint32_t cap_starts) {
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= pl.W) return;
    const int strand = blockIdx.y, mi = blockIdx.z;
    const DevModel &dm = models[mi];
    int c = grp2contig[w >> 2];
    int L = clen[c];
    int wl = (int)(w - cwoff[c]);
    if (L < 6 || wl * 32 > L - 3) return;  // no codon starts in this word (virtual stops sit at t <= L-3 too)
    FindCtx cx;
    cx.lo = pl.lo + (int64_t)strand * pl.W + cwoff[c];
    cx.hi = pl.hi + (int64_t)strand * pl.W + cwoff[c];
    cx.L = L;
    cx.nwords = (L + 31) >> 5;
    cx.stop_sel = dm.stop_sel;
    cx.atg_only = dm.atg_only;
    OrfRec *orfs = orfs_m[mi];
    uint32_t *starts = starts_m[mi];
    unsigned long long *counters = counters_m + (size_t)mi * 8;
    CodonMasks m = masks_at(cx, wl);
    uint32_t sm = stop_mask(m, cx.stop_sel);
    while (sm) {
        int b = __ffs(sm) - 1;
        sm &= sm - 1;
        handle_stop(cx, dm, c, strand, wl * 32 + b, false, orfs, starts, counters, cap_orf, cap_starts);
    }
    // right edge: in each frame the last full codon closes an open ORF unless it is a real stop
    for (int f = 0; f < 3; f++) {
        int tv = L - 3 - ((L - 3 - f) % 3 + 3) % 3;  // largest t <= L-3 with t % 3 == f
        if (tv < 0 || (tv >> 5) != wl) continue;
        uint32_t all = stop_mask(m, cx.stop_sel);
        if ((all >> (tv & 31)) & 1u) continue;
        handle_stop(cx, dm, c, strand, tv, true, orfs, starts, counters, cap_orf, cap_starts);
    }
}

// ------------------------------------------------------------------------------------------------
// ORF scoring: one warp per ORF record
// ------------------------------------------------------------------------------------------------
// Per position the kernel needs four numbers under the (bacterial, archaeal) section pair of a GC bin:
// coding log-likelihood and coding-minus-non-coding log-odds for each domain.  They sit in ONE 32-byte
// table entry (Quad), so a position costs one sector-sized gather.  Lanes own positions (stride 32 down
// from the stop) and accumulate privately; a warp reduction happens once per start codon, when the walk
// reaches the position where that start's full-order region begins.
struct Acc4 {
    double cB, oB, cA, oA;
};

__device__ __forceinline__ void window_bits(const uint32_t *lo, const uint32_t *hi, int first, int n, uint32_t &hb,
                                            uint32_t &lb) {
    int w = first >> 5, sh = first & 31;
    uint32_t mask = (1u << n) - 1u;
    lb = __funnelshift_r(lo[w], lo[w + 1], sh) & mask;
    hb = __funnelshift_r(hi[w], hi[w + 1], sh) & mask;
}

__device__ __forceinline__ Acc4 warp_sum4(Acc4 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        v.cB += __shfl_xor_sync(0xffffffffu, v.cB, o);
        v.oB += __shfl_xor_sync(0xffffffffu, v.oB, o);
        v.cA += __shfl_xor_sync(0xffffffffu, v.cA, o);
        v.oA += __shfl_xor_sync(0xffffffffu, v.oA, o);
    }
    return v;
}

__global__ void __launch_bounds__(256, 3) k_orf_score(Planes pl, const uint32_t *__restrict__ gcpre, const int64_t *__restrict__ cwoff,
                                                      const int32_t *__restrict__ clen, const DevModel *__restrict__ models,
                                                      OrfRec *const *orfs_m, uint32_t *const *starts_m,
                                                      const unsigned long long *counters_m, Cand *const *surv_m,
                                                      uint32_t *const *bcount_m, unsigned long long *surv_counters, uint32_t cap_orf,
                                                      uint32_t cap_surv, const ChainBufs *bufs_m) {
    const int mi = blockIdx.y;
    const DevModel &dm = models[mi];
    const int lane = lane_id();
    const int warps_per_block = blockDim.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
    unsigned long long n_orfs = counters_m[(size_t)mi * 8 + 0];
    if (n_orfs > cap_orf) n_orfs = cap_orf;
    const OrfRec *orfs = orfs_m[mi];
    const uint32_t *starts = starts_m[mi];
    Cand *surv = surv_m[mi];
    uint32_t *bcount = bcount_m[mi];
    int32_t *gptr = bufs_m[mi].gptr;
    const int kc = dm.kc, kn = dm.kn;
    const int nfull = dm.nfull, nnon = dm.nnon, nramp = dm.nramp, nnramp = dm.nnramp;
    const Quad *quad = dm.quad;
    const int ph_lane0 = (((-1 - lane - kc) % 3) + 3) % 3;  // phase of the window's first letter at i = t-1-lane

    for (int64_t oi = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); oi < (int64_t)n_orfs; oi += nwarps) {
        const OrfRec rec = orfs[oi];
        if (lane == 0) gptr[oi] = GPTR_UNSET;
        const int c = rec.contig, t = rec.t, strand = rec.sf & 1;
        const bool pstop = (rec.sf & 0x100u) != 0;
        const int L = clen[c];
        const int64_t w0 = cwoff[c];
        const uint32_t *lo = pl.lo + (int64_t)strand * pl.W + w0;
        const uint32_t *hi = pl.hi + (int64_t)strand * pl.W + w0;
        // type of the stop codon (uniform)
        int stop_type = 0;
        if (!pstop) {
            uint32_t hb, lb;
            window_bits(lo, hi, t, 3, hb, lb);
            // TAA: A at +1,+2 ; TAG: G at +2 ; TGA: G at +1
            bool g1 = ((hb >> 1) & 1u) && !((lb >> 1) & 1u);
            bool g2 = ((hb >> 2) & 1u) && !((lb >> 2) & 1u);
            stop_type = g1 ? 2 : (g2 ? 1 : 0);
        }
        for (uint32_t b0 = 0; b0 < rec.n; b0 += 32) {
            const bool has = b0 + lane < rec.n;
            uint32_t sraw = has ? starts[rec.base + b0 + lane] : 0u;
            const bool virt = (sraw >> 31) != 0;
            const int s = (int)(sraw & 0x7fffffffu);
            const int len = t + 3 - s;
            int gc = 0, secB = -1, secA = -1, pair = 0x7fffffff;
            if (has) {
                int a = strand ? (L - 1 - (t + 2)) : s, b = strand ? (L - 1 - s) : (t + 2);
                int gcn = gc_count_fwd(pl.lo, pl.hi, gcpre, w0, a, b);
                gc = (200 * gcn + len) / (2 * len);
                secB = dm.map[0][gc];
                secA = dm.map[1][gc];
                pair = dm.pair[gc];
            }
            // forward-strand candidates whose start lies beyond the left edge are summed from their first letter
            const int first = s + ((virt && strand == 0) ? 0 : 3);
            const int fullstart = first + kc;
            Acc4 acc = {0, 0, 0, 0};
            // ramp: orders 0..kc-1 at positions first .. first+kc-1 (window always starts on a codon boundary)
            if (has) {
                for (int j = 0; j < kc && first + j <= t - 1; j++) {
                    uint32_t hb, lb;
                    window_bits(lo, hi, first, j + 1, hb, lb);
                    uint32_t idx = (hb << (j + 1)) | lb;
                    uint32_t idn;
                    const double *lnB, *lnA;
                    if (j < kn) {
                        idn = idx;
                        lnB = dm.ln_ramp + ((size_t)max(secB, 0) * dm.knr + j) * nnramp;
                        lnA = dm.ln_ramp + ((size_t)max(secA, 0) * dm.knr + j) * nnramp;
                    } else {
                        idn = ((hb >> (j - kn)) << (kn + 1)) | (lb >> (j - kn));
                        lnB = dm.ln_full + (size_t)max(secB, 0) * nnon;
                        lnA = dm.ln_full + (size_t)max(secA, 0) * nnon;
                    }
                    if (secB >= 0) {
                        double cv = dm.lc_ramp[(((size_t)secB * dm.kr + j) * 3 + 0) * nramp + idx];
                        acc.cB += cv;
                        acc.oB += cv - lnB[idn];
                    }
                    if (secA >= 0) {
                        double cv = dm.lc_ramp[(((size_t)secA * dm.kr + j) * 3 + 0) * nramp + idx];
                        acc.cA += cv;
                        acc.oA += cv - lnA[idn];
                    }
                }
            }
            // full-order part: one walk from t-1 down per distinct section pair among this batch's starts
            bool done = !has || fullstart > t - 1;
            while (true) {
                const int psel = warp_min_i(done ? 0x7fffffff : pair);
                if (psel == 0x7fffffff) break;
                const bool mine = !done && pair == psel;
                const int lo_pos = warp_min_i(mine ? fullstart : 0x7fffffff);
                const Quad *tq = quad + (size_t)psel * 3 * nfull;
                Acc4 run = {0, 0, 0, 0};  // lane-private sum over this lane's positions visited so far
                Acc4 res = {0, 0, 0, 0};
                int ph = ph_lane0;
                for (int hiPos = t - 1; hiPos >= lo_pos; hiPos -= 32) {
                    const int i = hiPos - lane;
                    Acc4 v = {0, 0, 0, 0};
                    if (i >= lo_pos) {
                        uint32_t hb, lb;
                        window_bits(lo, hi, i - kc, kc + 1, hb, lb);
                        const Quad *q = tq + (size_t)ph * nfull + ((hb << (kc + 1)) | lb);
                        const double2 x = __ldg(reinterpret_cast<const double2 *>(q));
                        const double2 y = __ldg(reinterpret_cast<const double2 *>(q) + 1);
                        v.cB = x.x; v.oB = x.y; v.cA = y.x; v.oA = y.y;
                    }
                    // starts whose full-order region begins inside this chunk take their suffix sum now
                    uint32_t hits = __ballot_sync(0xffffffffu, mine && fullstart <= hiPos && fullstart > hiPos - 32);
                    while (hits) {
                        const int o = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const int tl = hiPos - __shfl_sync(0xffffffffu, fullstart, o);
                        Acc4 part = run;
                        if (lane <= tl) { part.cB += v.cB; part.oB += v.oB; part.cA += v.cA; part.oA += v.oA; }
                        part = warp_sum4(part);
                        if (lane == o) res = part;
                    }
                    run.cB += v.cB; run.oB += v.oB; run.cA += v.cA; run.oA += v.oA;
                    ph = ph == 2 ? 0 : ph + 1;
                }
                if (mine) { acc.cB += res.cB; acc.oB += res.oB; acc.cA += res.cA; acc.oA += res.oA; done = true; }
            }
            // totals
            bool keep = false;
            Cand cd;
            if (has && (secB >= 0 || secA >= 0)) {
                const bool arc = secB < 0 || (secA >= 0 && acc.cA > acc.cB);
                const SectionScalars *sc = dm.sc + (arc ? secA : secB);
                double dl = (double)len;
                double total = (arc ? acc.oA : acc.oB) + 2.0 * log(dl) + dl * sc->dur_slope + sc->dur_const - 1.3862943611198906;
                if (!virt) {
                    uint32_t hb, lb;
                    window_bits(lo, hi, s, 1, hb, lb);
                    total += sc->start_lo[hb ? (lb ? 2 : 1) : 0];  // A -> ATG, G -> GTG, T -> TTG
                }
                if (!pstop) total += sc->stop_lo[stop_type];
                keep = total > 0.0;
                cd.contig = c;
                cd.left = strand ? (L - 1 - (t + 2)) : s;
                cd.right = strand ? (L - 1 - s) : (t + 2);
                cd.orf = (uint32_t)oi;
                cd.total = total;
                cd.gc = (int16_t)gc;
                cd.strand = strand ? -1 : 1;
                cd.flags = (uint8_t)((arc ? 1 : 0) | (virt ? 2 : 0) | (pstop ? 4 : 0));
                cd.len = len;
            }
            const uint32_t km = __ballot_sync(0xffffffffu, keep);
            if (km && strand && rec.n <= 32) {
                // reverse strand: the survivor with the largest s has the smallest forward right end, so the chain meets it
                // first among the starts of this ORF and can skip the lookup of the ORF's shared back-pointer
                const int smax = -warp_min_i(keep ? -s : 0x7fffffff);
                const bool any_virt = __any_sync(0xffffffffu, keep && virt);
                if (keep && s == smax && !any_virt) cd.flags |= 8;
            }
            if (km) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&surv_counters[mi], (unsigned long long)__popc(km));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) {
                    unsigned long long k = base + __popc(km & ((1u << lane) - 1u));
                    if (k < cap_surv) {
                        surv[k] = cd;
                        atomicAdd(&bcount[(w0 * 32 + cd.right) >> 7], 1u);
                    }
                }
            }
        }
    }
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

__device__ __forceinline__ void chain_append(const ChainBufs &B, ChainRing &R, uint32_t cs, ChainState &S, int gi, double cum,
                                             int right, int left, int ls) {
    // all lanes hold identical arguments; lane 0 writes
    if (S.head == 0 || cum > S.pm_val) { S.pm_val = cum; S.pm_arg = gi; }
    if (lane_id() == 0) {
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

// Processes the candidates held one per lane (valid lanes: `pending`) in chain order.
__device__ __forceinline__ void chain_bucket(const ChainBufs &B, ChainRing &R, uint32_t cs, ChainState &S, uint32_t bs, const Cand &me,
                                             bool pending) {
    const int lane = lane_id();
    while (true) {
        const int rmin = warp_min_i(pending ? me.right : 0x7fffffff);
        if (rmin == 0x7fffffff) break;
        const bool ing = pending && me.right == rmin;
        pending = pending && !ing;
        double b_non = 0, b_ov = 0;
        int bp_non = -1, bp_ov = -1;
        if (ing) chain_eval_lane(B, R, cs, S, me.left, me.len, me.strand, b_non, bp_non, b_ov, bp_ov);
        double cumg = me.total + (bp_ov >= 0 ? b_ov : b_non);
        int prevg = bp_ov >= 0 ? bp_ov : bp_non;
        if (ing) {
            B.cum[bs + lane] = cumg;
            B.prev[bs + lane] = prevg;
        }
        // reverse-strand members: shared pointer of their ORF, then append (ascending left)
        uint32_t rev = __ballot_sync(0xffffffffu, ing && me.strand < 0);
        while (rev) {
            // smallest left first
            const int lmin = warp_min_i(((rev >> lane) & 1u) ? me.left : 0x7fffffff);
            const int o = __ffs(__ballot_sync(0xffffffffu, ((rev >> lane) & 1u) && me.left == lmin)) - 1;
            rev &= ~(1u << o);
            const uint32_t gorf = __shfl_sync(0xffffffffu, me.orf, o);
            const int gfl = __shfl_sync(0xffffffffu, (int)me.flags, o);
            const int o_non = __shfl_sync(0xffffffffu, bp_non, o), o_ov = __shfl_sync(0xffffffffu, bp_ov, o);
            int gp = gp_lookup(B, R, gorf, gfl);
            if (gp == GPTR_UNSET) gp = o_non;
            if (o_ov >= 0) gp = o_ov;
            gp_store(B, R, S, gorf, gp);
            chain_append(B, R, cs, S, (int)(bs + o), __shfl_sync(0xffffffffu, cumg, o), rmin, lmin,
                         __shfl_sync(0xffffffffu, me.len, o) * 2 + 1);
        }
        // direct-strand members: the stop keeps one state, its best start (ties: smallest left)
        const bool dirm = ing && me.strand > 0;
        if (__any_sync(0xffffffffu, dirm)) {
            double bc = dirm ? cumg : -1e300;
            int bl = dirm ? me.left : 0x7fffffff, bo = lane;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
                const int ol = __shfl_xor_sync(0xffffffffu, bl, o), oo = __shfl_xor_sync(0xffffffffu, bo, o);
                if (oc > bc || (oc == bc && ol < bl)) { bc = oc; bl = ol; bo = oo; }
            }
            chain_append(B, R, cs, S, (int)(bs + bo), bc, rmin, bl, __shfl_sync(0xffffffffu, me.len, bo) * 2);
        }
        S.na = S.head;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) k_chain(const ChainBufs *bufs_m, const int64_t *__restrict__ cwoff,
                                               const int32_t *__restrict__ clen, const int32_t *__restrict__ order,
                                               int32_t n_contigs, int *work_counters) {
    __shared__ ChainRing rings[4];
    const int mi = blockIdx.y;
    const ChainBufs B = bufs_m[mi];
    const int lane = lane_id();
    ChainRing &R = rings[threadIdx.x >> 5];
    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&work_counters[mi], 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_contigs) break;
        const int c = order[item];
        const int L = clen[c];
        const int64_t g0 = cwoff[c] * 32;
        const int64_t b_first = g0 >> 7, b_last = (g0 + max(L - 1, 0)) >> 7;
        const uint32_t cs = b_first ? B.bend[b_first - 1] : 0u;
        const uint32_t ce = B.bend[b_last];
        ChainState S;
        S.na = 0; S.head = 0; S.pm_val = 0.0; S.pm_arg = -1; S.gp_next = 0;
        if (lane < GP_CACHE) R.gp_orf[lane] = 0xffffffffu;
        __syncwarp();
        // the contig's candidates are contiguous in `sorted`; walk them in register-sized pieces that end on bucket boundaries
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
                                chain_append(B, R, cs, S, (int)(k0 + o), __shfl_sync(0xffffffffu, cumg, o), r,
                                             __shfl_sync(0xffffffffu, g.left, o), __shfl_sync(0xffffffffu, g.len, o) * 2 + 1);
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
                        if (bidx >= 0 && bc > -1e299) chain_append(B, R, cs, S, bidx, bc, r, bleft, blen * 2);
                        S.na = S.head;
                        __syncwarp();
                    }
                }
            }
        }
        // traceback (reverse-strand genes follow their ORF's shared pointer)
        if (lane == 0) {
            int ng = 0;
            double tot = 0.0;
            if (S.head > 0) {
                tot = S.pm_val;
                int g = S.pm_arg;
                const int limit = (int)(ce - cs);
                int k = limit;
                while (g >= 0 && k > 0) {
                    B.chain[cs + (--k)] = g;
                    const Cand cd = B.sorted[g];
                    g = cd.strand < 0 ? B.gptr[cd.orf] : B.prev[g];
                    ng++;
                }
                for (int q = 0; q < ng; q++) B.chain[cs + q] = B.chain[cs + k + q];
            }
            B.gene_count[c] = ng;
            B.contig_total[c] = tot;
        }
        __syncwarp();
    }
}

// genes of all contigs, contiguous, ordered by (contig, left); per-model sum of round(logodd, 2) in cents
__global__ void __launch_bounds__(128) k_gather(const ChainBufs *bufs_m, const uint32_t *const *gene_off_m, mgc_gene *const *genes_m,
                                                const int64_t *__restrict__ cwoff, const int32_t *__restrict__ clen,
                                                int32_t n_contigs, long long *model_cents) {
    const int mi = blockIdx.y;
    const ChainBufs B = bufs_m[mi];
    const uint32_t *goff = gene_off_m[mi];
    mgc_gene *genes = genes_m[mi];
    const int lane = lane_id();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    long long cents = 0;
    for (int64_t c = warp; c < n_contigs; c += nw) {
        const int ng = B.gene_count[c];
        if (!ng) continue;
        const int64_t g0 = cwoff[c] * 32;
        const int64_t b_first = g0 >> 7;
        const uint32_t cs = b_first ? B.bend[b_first - 1] : 0u;
        for (int q = lane; q < ng; q += 32) {
            const Cand cd = B.sorted[B.chain[cs + q]];
            mgc_gene g;
            g.contig = (int32_t)c;
            g.left = cd.left + 1;
            g.right = cd.right + 1;
            g.strand = cd.strand;
            g.domain = cd.flags & 1;
            const int ps = (cd.flags >> 1) & 1, pe = (cd.flags >> 2) & 1;
            g.partial_left = cd.strand > 0 ? ps : pe;
            g.partial_right = cd.strand > 0 ? pe : ps;
            g.gc = cd.gc;
            g.logodd = cd.total;
            genes[goff[c] + q] = g;
            cents += llrint(cd.total * 100.0);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cents += __shfl_xor_sync(0xffffffffu, cents, o);
    if (lane == 0 && cents) atomicAdd((unsigned long long *)&model_cents[mi], (unsigned long long)cents);
}

// ------------------------------------------------------------------------------------------------
// window sums and labels (reference bin/pipeline.py:157-385); thread per (window, model) -> warp per window batch
// ------------------------------------------------------------------------------------------------
// cents[w*M + m] = sum over genes of model m on the contig with start <= gene.left(1-based) < end; if that is 0 and the
// last gene starting before the window ends after it, that gene's value (pipeline.py:290-321).
__global__ void __launch_bounds__(128) k_windows(const mgc_gene *const *genes_m, const uint32_t *const *gene_off_m, int n_models,
                                                 const int32_t *__restrict__ clen, const int32_t *__restrict__ stride,
                                                 const int64_t *__restrict__ win_off, int32_t n_contigs, int32_t window,
                                                 int32_t consecutive, int short_contigs, int32_t *labels, long long *cents_out) {
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
    if (threadIdx.x < MGC_MAX_MODELS) tot[threadIdx.x] = 0;
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
    }
}

// ------------------------------------------------------------------------------------------------
// host-callable launch wrappers (keep <<< >>> in this translation unit)
// ------------------------------------------------------------------------------------------------
void launch_groups(const int64_t *cwoff, int32_t n_contigs, int64_t n_groups, int32_t *grp2contig, cudaStream_t st) {
    k_groups<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(cwoff, n_contigs, n_groups, grp2contig);
}

void launch_pack(const char *ascii, const int64_t *aoff, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig,
                 int64_t W, uint32_t *lo, uint32_t *hi, uint32_t *gccount, int *bad_letter, cudaStream_t st) {
    int64_t warps = (W + 31) / 32;
    dim3 grid((unsigned)((warps + 7) / 8), 2);
    k_pack<<<grid, 256, 0, st>>>(ascii, aoff, cwoff, clen, grp2contig, W, lo, hi, gccount, bad_letter);
}

void launch_orf_find(Planes pl, const int64_t *cwoff, const int32_t *clen, const int32_t *grp2contig, const DevModel *models,
                     int n_models, OrfRec *const *orfs_m, uint32_t *const *starts_m, unsigned long long *counters_m,
                     uint32_t cap_orf, uint32_t cap_starts, cudaStream_t st) {
    dim3 grid((unsigned)((pl.W + 127) / 128), 2, n_models);
    k_orf_find<<<grid, 128, 0, st>>>(pl, cwoff, clen, grp2contig, models, orfs_m, starts_m, counters_m, cap_orf, cap_starts);
}

void launch_orf_score(Planes pl, const uint32_t *gcpre, const int64_t *cwoff, const int32_t *clen, const DevModel *models,
                      int n_models, OrfRec *const *orfs_m, uint32_t *const *starts_m, const unsigned long long *counters_m,
                      Cand *const *surv_m, uint32_t *const *bcount_m, unsigned long long *surv_counters, uint32_t cap_orf,
                      uint32_t cap_surv, const ChainBufs *bufs_m, int n_sm, cudaStream_t st) {
    dim3 grid(n_sm * 8, n_models);
    k_orf_score<<<grid, 256, 0, st>>>(pl, gcpre, cwoff, clen, models, orfs_m, starts_m, counters_m, surv_m, bcount_m,
                                      surv_counters, cap_orf, cap_surv, bufs_m);
}

void launch_scatter(const Cand *const *surv_m, Cand *const *sorted_m, uint32_t *const *boff_m,
                    const unsigned long long *surv_counters, const int64_t *cwoff, uint32_t cap_surv, int n_models, int n_sm,
                    cudaStream_t st) {
    dim3 grid(n_sm * 4, n_models);
    k_scatter<<<grid, 256, 0, st>>>(surv_m, sorted_m, boff_m, surv_counters, cwoff, cap_surv);
}

void launch_chain(const ChainBufs *bufs_m, const int64_t *cwoff, const int32_t *clen, const int32_t *order, int32_t n_contigs,
                  int *work_counters, int n_models, int n_sm, cudaStream_t st) {
    int64_t warps = n_contigs;
    int64_t blocks = (warps + 3) / 4;
    if (blocks > (int64_t)n_sm * 16) blocks = (int64_t)n_sm * 16;
    dim3 grid((unsigned)blocks, n_models);
    k_chain<<<grid, 128, 0, st>>>(bufs_m, cwoff, clen, order, n_contigs, work_counters);
}

void launch_gather(const ChainBufs *bufs_m, const uint32_t *const *gene_off_m, mgc_gene *const *genes_m, const int64_t *cwoff,
                   const int32_t *clen, int32_t n_contigs, long long *model_cents, int n_models, int n_sm, cudaStream_t st) {
    int64_t blocks = ((int64_t)n_contigs + 3) / 4;
    if (blocks > (int64_t)n_sm * 16) blocks = (int64_t)n_sm * 16;
    dim3 grid((unsigned)blocks, n_models);
    k_gather<<<grid, 128, 0, st>>>(bufs_m, gene_off_m, genes_m, cwoff, clen, n_contigs, model_cents);
}

void launch_windows(const mgc_gene *const *genes_m, const uint32_t *const *gene_off_m, int n_models, const int32_t *clen,
                    const int32_t *stride, const int64_t *win_off, int32_t n_contigs, int32_t window, int32_t consecutive,
                    int short_contigs, int32_t *labels, long long *cents, cudaStream_t st) {
    k_windows<<<(unsigned)n_contigs, 128, 0, st>>>(genes_m, gene_off_m, n_models, clen, stride, win_off, n_contigs, window,
                                                    consecutive, short_contigs, labels, cents);
}

}  // namespace mgc
