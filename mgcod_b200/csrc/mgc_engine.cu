// mgc_engine.cu — host orchestration of the CUDA path and the C-ABI entry points (include/mgcod_b200.h).
//
// One mgc_engine per process per GPU: device copies of the model tables, the resident batch (planes,
// GC prefix), per-model scratch, one stream.  Nothing here falls back to a CPU path: without a usable
// device mgc_engine_create fails with MGC_E_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "mgc_device.cuh"
#include "mgc_kernels.hpp"
#include "mgc_model.hpp"
#include "mgc_util.hpp"

using namespace mgc;

struct mgc_model {
    Model m;
};

// Page-locked host blocks are expensive to create (cudaHostAlloc pins pages), so results borrow them from a
// pool owned jointly by the engine and its outstanding results.
// completion of a result's device->host copy (shared by the result, the pool's deferred releases and the engine)
// Test knob: MGCOD_B200_WALK_TILE_WORDS=<words> shrinks the position tiles of the walk-item order so that a small input
// runs through the multi-tile code (results do not depend on the tiling).
static long long walk_tile_words_setting() {
    const char *s = std::getenv("MGCOD_B200_WALK_TILE_WORDS");
    return s ? std::atoll(s) : WALK_TILE_WORDS;
}

struct Ready {
    cudaEvent_t ev = nullptr;
    Ready() { cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); }
    ~Ready() { if (ev) cudaEventDestroy(ev); }
    bool done() const { return cudaEventQuery(ev) == cudaSuccess; }
    void wait() const { cudaEventSynchronize(ev); }
};

struct PinnedPool {
    struct Block {
        void *p = nullptr;
        size_t bytes = 0;
    };
    std::mutex mu;
    std::vector<Block> free_blocks;
    std::vector<std::pair<Block, std::shared_ptr<Ready>>> deferred;   // released while their copy was still in flight
    void sweep_locked() {
        for (size_t i = 0; i < deferred.size();)
            if (deferred[i].second->done()) {
                free_blocks.push_back(deferred[i].first);
                deferred[i] = deferred.back();
                deferred.pop_back();
            } else i++;
    }
    Block acquire(size_t need) {
        need = std::max<size_t>(need, 4096);
        {
            std::lock_guard<std::mutex> lk(mu);
            sweep_locked();
            size_t best = free_blocks.size();
            for (size_t i = 0; i < free_blocks.size(); i++)
                if (free_blocks[i].bytes >= need && (best == free_blocks.size() || free_blocks[i].bytes < free_blocks[best].bytes)) best = i;
            if (best != free_blocks.size()) {
                Block b = free_blocks[best];
                free_blocks.erase(free_blocks.begin() + best);
                return b;
            }
        }
        Block b;
        b.bytes = need + need / 4;
        if (cudaHostAlloc(&b.p, b.bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            throw std::bad_alloc();
        }
        return b;
    }
    void release(Block b, const std::shared_ptr<Ready> &ready = nullptr) {
        if (!b.p) return;
        std::lock_guard<std::mutex> lk(mu);
        if (ready && !ready->done()) deferred.emplace_back(b, ready);
        else free_blocks.push_back(b);
    }
    ~PinnedPool() {
        for (auto &d : deferred) { d.second->wait(); cudaFreeHost(d.first.p); }
        for (auto &b : free_blocks) cudaFreeHost(b.p);
    }
};

struct mgc_result {
    int n_models = 0;
    int32_t n_contigs = 0;
    std::shared_ptr<PinnedPool> pool;
    std::vector<PinnedPool::Block> gene_blk, total_blk;
    std::vector<int64_t> n_genes;
    std::vector<int64_t> cents;
    std::vector<int32_t> min_len;   // $GENE_MIN_LENGTH of each model: records shorter than it get no summary line
    bool host_owned = false;   // blocks came from malloc (mgc_result_create), not from the pinned pool
    std::shared_ptr<Ready> ready;   // the genes / totals travel on a copy stream while the engine goes on; readers wait here
    void wait() const { if (ready) ready->wait(); }
    ~mgc_result() {
        if (host_owned) {
            for (auto &b : gene_blk) std::free(b.p);
            for (auto &b : total_blk) std::free(b.p);
        } else if (pool) {
            for (auto &b : gene_blk) pool->release(b, ready);
            for (auto &b : total_blk) pool->release(b, ready);
        }
    }
};

namespace {

#define CUDA_TRY(expr)                                                                                 \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(_e) + " at " #expr);   \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    void ensure(size_t need) {
        if (need <= bytes) return;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA: out of device memory (") + std::to_string(need) + " bytes)");
        bytes = need;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PerModel {
    DevBuf lc_full, ln_full, lc_ramp, ln_ramp, sc, quad, rquad, wtab, wtab2, pair_amb, pair_ub;
    DevBuf ls, ls_tile, orfs, starts, items, items_sorted, hist, item_off, unit_off, sums;
    DevBuf surv, sorted, bcount, cum, prev, gptr;
    DevBuf av_right, av_left, av_ls, av_idx, av_cum, av_pmax, av_parg, st, st_rl, chain, gene_count, gene_off, contig_total, genes;
    DevBuf piece_b0, piece_b1, work, piece_total, slot_gene_off, cover, minleft;
};

constexpr int N_EVENTS = 12;

}  // namespace

struct mgc_engine {
    int device = 0, n_sm = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[N_EVENTS];
    int n_models = 0;
    std::vector<Model> models;
    std::vector<PerModel> pm;
    DevBuf d_models;
    // resident batch
    int32_t n_contigs = 0;
    int32_t n_slots = 0;   // chain piece slots (ChainBufs)
    DevBuf slotbase, slot_contig;
    int64_t total_nt = 0, W = 0;
    std::vector<int64_t> h_aoff, h_cwoff;
    std::vector<int32_t> h_clen;
    // the caller's contigs ("parents") and the slots actually decoded: one slot per contig, or several when runs of
    // >= 40 ambiguous letters split a contig (gmhmmp2 --incomplete_at_gaps 40)
    struct Seg {
        int32_t parent;
        int64_t aoff;      // offset of its first letter in the device ASCII buffer
        int32_t len;
        int32_t poff;      // offset inside the parent
        uint8_t right_end; // its right edge is the end of the record
    };
    std::vector<Seg> segs;
    std::vector<int64_t> h_parent_aoff;
    std::vector<int32_t> h_parent_len;
    int32_t n_parents = 0;
    bool segmented = false;
    DevBuf nm, ncount, cflags, plen, gaps, gap_count, codon;
    DevBuf ascii, aoff, cwoff, clen, grp2contig, lo, hi, gc, tile_sums, flags;
    // next batch on its way to the device while the current one is decoded (mgc_prefetch_batch)
    DevBuf ascii_next;
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t copy_done = nullptr, main_done = nullptr, d2h_begin = nullptr, d2h_end = nullptr;
    std::shared_ptr<Ready> last_ready;   // copy of the previous result: the next gather must not overwrite its source
    bool d2h_timed = true;
    // host side of a batch's slot tables (set_slots), computable before the batch is adopted
    struct SlotTables {
        std::vector<int32_t> clen, slotbase, slot_contig;
        std::vector<int64_t> cwoff, aoff;
        std::vector<uint8_t> cflags;
        int32_t n = 0, n_slots = 0;
        int64_t W = 0, total_nt = 0;
    };
    // what do_upload derives from the caller's offsets, prepared by a background thread while the previous batch is decoded
    struct Prepared {
        std::vector<Seg> segs;
        std::vector<int64_t> parent_aoff;
        std::vector<int32_t> parent_len;
        SlotTables tables;
        std::future<void> ready;
        bool valid = false;
    } prepared;
    struct {
        bool valid = false;
        const char *seq = nullptr;
        std::vector<int64_t> offsets;
    } pending;
    // pointer tables + counters (device)
    DevBuf t_orfs, t_starts, t_surv, t_sorted, t_bcount, t_cover, t_gene_off, t_genes, t_chainbufs;
    DevBuf t_scan_buckets, t_gene_count, t_slot_gene_off;   // pointer tables of the batched scans
    DevBuf t_ls, t_ls_tile, t_items, t_items_sorted, t_hist, t_item_off, t_unit_off, t_sums, unit_totals, next_unit;
    uint32_t cap_items = 0;
    int max_cod_order = 0, max_non_order = 0, max_pairs = 1;
    bool orders_uniform = false;   // all models share one (COD_ORDER, NON_ORDER)
    int pair_smem = 0;     // bytes of the pair walk's shared-memory table; 0 = single-position walk (order 5, or MGCOD_B200_WALK=single)
    DevBuf counters;     // [M][8] unsigned long long: 0 orfs, 1 starts
    DevBuf surv_counters;  // [M]
    PinnedPool::Block table_stage;   // page-locked staging of the slot tables (set_slots)
    DevBuf work_counters;  // [M][4] int: long pieces, consumers' cursor, other pieces, unused
    int chain_mode = -1;      // -1: by batch size (launch_chain's caller); MGCOD_B200_CHAIN=hybrid / warp / thread forces 0 / 1 / 2 (tests, A/B measurements)
    DevBuf model_cents;    // [M]
    DevBuf gene_totals;    // [M] uint32 grand totals of the gene scans
    DevBuf win_stride, win_off, win_labels, win_cents, win_merged, win_run, win_status, win_hist;   // labelling scratch, kept between calls
    PinnedPool::Block win_host;   // page-locked staging of the labelling results
    uint32_t cap_orf = 0, cap_starts = 0, cap_surv = 0;
    int cap_scale = 1;     // multiplier of the scratch capacities, raised on overflow
    std::vector<unsigned char> table_sig;   // contents of the device pointer tables as last uploaded
    bool have_batch = false, have_genes = false;
    std::vector<int64_t> dev_gene_count;   // genes of the last decode per model, as they sit in pm[i].genes on the device
    mgc_stats stats;
    std::shared_ptr<PinnedPool> pool = std::make_shared<PinnedPool>();
};

namespace {

void upload_models(mgc_engine *e) {
    std::vector<DevModel> dm(e->n_models);
    for (int i = 0; i < e->n_models; i++) {
        const Model &m = e->models[i];
        PerModel &p = e->pm[i];
        auto up = [&](DevBuf &b, const void *src, size_t bytes) {
            b.ensure(std::max<size_t>(bytes, 16));
            CUDA_TRY(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
        };
        up(p.lc_full, m.lc_full.data(), m.lc_full.size() * sizeof(double));
        up(p.ln_full, m.ln_full.data(), m.ln_full.size() * sizeof(double));
        up(p.lc_ramp, m.lc_ramp.data(), m.lc_ramp.size() * sizeof(double));
        up(p.ln_ramp, m.ln_ramp.data(), m.ln_ramp.size() * sizeof(double));
        up(p.sc, m.sc.data(), m.sc.size() * sizeof(SectionScalars));
        DevModel &d = dm[i];
        d.gcode = m.gcode;
        d.kc = m.cod_order;
        d.kn = m.non_order;
        d.min_len = std::max(m.min_len, 6);
        d.nsec = m.nsec;
        d.kr = m.cod_order > 0 ? m.cod_order : 1;
        d.knr = m.non_order > 0 ? m.non_order : 1;
        switch (m.gcode) {
            case 4: case 25: d.stop_sel = 1 | 2; break;   // TAA TAG
            case 15: d.stop_sel = 1 | 4; break;           // TAA TGA
            case 101: d.stop_sel = 2 | 4; break;          // TAG TGA
            default: d.stop_sel = 7;
        }
        d.atg_only = m.gcode == 1;
        d.nfull = (int)pow4(d.kc + 1);
        d.nnon = (int)pow4(d.kn + 1);
        d.nramp = (int)pow4(d.kr);
        d.nnramp = (int)pow4(d.knr);
        d.lc_full = p.lc_full.as<double>();
        d.ln_full = p.ln_full.as<double>();
        d.lc_ramp = p.lc_ramp.as<double>();
        d.ln_ramp = p.ln_ramp.as<double>();
        d.sc = p.sc.as<SectionScalars>();
        std::memcpy(d.map, m.map, sizeof(d.map));
        // section pairs: every gc maps to (secB, secA); distinct pairs get one Quad table each
        std::vector<std::pair<int, int>> pairs;
        for (int g = 0; g <= 100; g++) {
            std::pair<int, int> pr(m.map[0][g], m.map[1][g]);
            size_t k = std::find(pairs.begin(), pairs.end(), pr) - pairs.begin();
            if (k == pairs.size()) pairs.push_back(pr);
            d.pair[g] = (int)k;
        }
        d.npair = (int)pairs.size();
        const int kc = d.kc, kn = d.kn;
        const size_t nfull = (size_t)d.nfull, nnon = (size_t)d.nnon;
        std::vector<Quad> q(pairs.size() * 3 * nfull);
        for (size_t pi = 0; pi < pairs.size(); pi++)
            for (int ph = 0; ph < 3; ph++)
                for (size_t idx = 0; idx < nfull; idx++) {
                    // plane-major index: hi bits above lo bits, kc+1 letters; the non-coding context is the last kn+1 letters
                    const uint32_t hb = (uint32_t)(idx >> (kc + 1)), lb = (uint32_t)(idx & ((1u << (kc + 1)) - 1u));
                    const size_t idn = ((size_t)(hb >> (kc - kn)) << (kn + 1)) | (lb >> (kc - kn));
                    Quad &e4 = q[(pi * 3 + ph) * nfull + idx];
                    e4 = Quad{0, 0, 0, 0};
                    if (pairs[pi].first >= 0) {
                        const size_t sB = (size_t)pairs[pi].first;
                        e4.cB = m.lc_full[(sB * 3 + ph) * nfull + idx];
                        e4.oB = e4.cB - m.ln_full[sB * nnon + idn];
                    }
                    if (pairs[pi].second >= 0) {
                        const size_t sA = (size_t)pairs[pi].second;
                        e4.cA = m.lc_full[(sA * 3 + ph) * nfull + idx];
                        e4.oA = e4.cA - m.ln_full[sA * nnon + idn];
                    }
                }
        up(p.quad, q.data(), q.size() * sizeof(Quad));
        d.quad = p.quad.as<Quad>();
        // image of the walk kernel's shared-memory table (see k_walk): per pair, coding {B, A} entries, each stored twice
        // side by side (even / odd 16-byte slots for even / odd lanes), then non-coding {B, A} entries stored eight times
        {
            const size_t n16 = 3 * nfull * 2 + nnon * 8;
            std::vector<double> img(pairs.size() * n16 * 2, 0.0);
            for (size_t pi = 0; pi < pairs.size(); pi++) {
                double *base = img.data() + pi * n16 * 2;
                for (size_t ent = 0; ent < 3 * nfull; ent++) {
                    const Quad &e4 = q[pi * 3 * nfull + ent];
                    for (int rep = 0; rep < 2; rep++) {
                        const size_t o16 = (ent << 1) | (size_t)rep;
                        base[o16 * 2] = e4.cB;
                        base[o16 * 2 + 1] = e4.cA;
                    }
                }
                double *nb = base + 3 * nfull * 2 * 2;
                for (size_t idn = 0; idn < nnon; idn++)
                    for (int rep = 0; rep < 8; rep++) {
                        const size_t o16 = (idn << 3) | (size_t)rep;
                        nb[o16 * 2] = pairs[pi].first >= 0 ? m.ln_full[(size_t)pairs[pi].first * nnon + idn] : 0.0;
                        nb[o16 * 2 + 1] = pairs[pi].second >= 0 ? m.ln_full[(size_t)pairs[pi].second * nnon + idn] : 0.0;
                    }
            }
            up(p.wtab, img.data(), img.size() * sizeof(double));
            d.wtab = p.wtab.as<double2>();
            d.wtab_n16 = (int)n16;
        }
        // image of the pair walk's table (k_walk_pair): per pair and codon position p of the window's first letter, for every
        // (kc+2)-mer a0..a(kc+1) the sum lnP(a_kc | a0.., p) + lnP(a_(kc+1) | a1.., p+1) of both domains, rc2 replicas side by
        // side; then the non-coding sums of the (kn+2)-mers, eight replicas.  rc2 = the most replicas that fit one SM.
        {
            const size_t npc = pow4(kc + 2), npn = pow4(kn + 2);
            int rc = 8;
            while (rc >= 1 && (3 * npc * (size_t)rc + npn * 8) * 16 > (size_t)WALK2_SMEM_MAX) rc >>= 1;
            d.rc2 = rc;
            d.wtab2 = nullptr;
            d.wtab2_n16 = 0;
            if (rc >= 1) {
                const size_t n16 = 3 * npc * rc + npn * 8;
                std::vector<double> img(pairs.size() * n16 * 2, 0.0);
                const uint32_t m1 = (1u << (kc + 1)) - 1u;
                for (size_t pi = 0; pi < pairs.size(); pi++) {
                    double *base = img.data() + pi * n16 * 2;
                    for (int ph = 0; ph < 3; ph++)
                        for (size_t idx = 0; idx < npc; idx++) {
                            const uint32_t hb = (uint32_t)(idx >> (kc + 2)), lb = (uint32_t)(idx & ((1u << (kc + 2)) - 1u));
                            // bit 0 = first letter: the lower position's window is letters 0..kc, the upper one's 1..kc+1
                            const size_t i0 = ((size_t)(hb & m1) << (kc + 1)) | (lb & m1);
                            const size_t i1 = ((size_t)(hb >> 1) << (kc + 1)) | (lb >> 1);
                            const Quad &q0 = q[(pi * 3 + ph) * nfull + i0], &q1 = q[(pi * 3 + (ph + 1) % 3) * nfull + i1];
                            for (int rep = 0; rep < rc; rep++) {
                                const size_t o16 = ((size_t)ph * npc + idx) * rc + rep;
                                base[o16 * 2] = q0.cB + q1.cB;
                                base[o16 * 2 + 1] = q0.cA + q1.cA;
                            }
                        }
                    double *nb = base + 3 * npc * rc * 2;
                    const uint32_t n1 = (1u << (kn + 1)) - 1u;
                    for (size_t idn = 0; idn < npn; idn++) {
                        const uint32_t hb = (uint32_t)(idn >> (kn + 2)), lb = (uint32_t)(idn & ((1u << (kn + 2)) - 1u));
                        const size_t i0 = ((size_t)(hb & n1) << (kn + 1)) | (lb & n1);
                        const size_t i1 = ((size_t)(hb >> 1) << (kn + 1)) | (lb >> 1);
                        for (int dom = 0; dom < 2; dom++) {
                            const int sec = dom ? pairs[pi].second : pairs[pi].first;
                            const double v = sec >= 0 ? m.ln_full[(size_t)sec * nnon + i0] + m.ln_full[(size_t)sec * nnon + i1] : 0.0;
                            for (int rep = 0; rep < 8; rep++) nb[((idn << 3) | (size_t)rep) * 2 + dom] = v;
                        }
                    }
                }
                up(p.wtab2, img.data(), img.size() * sizeof(double));
                d.wtab2 = p.wtab2.as<double2>();
                d.wtab2_n16 = (int)n16;
            }
        }
        // order ramp: order j < kc, window of j+1 letters starting on a codon boundary (phase 0);
        // non-coding order is min(j, kn): its context is the last min(j,kn)+1 letters of the window
        const int kr = d.kr, knr = d.knr;
        const size_t nramp = (size_t)d.nramp, nnramp = (size_t)d.nnramp;
        std::vector<Quad> rq(pairs.size() * kr * nramp, Quad{0, 0, 0, 0});
        for (size_t pi = 0; pi < pairs.size(); pi++)
            for (int j = 0; j < kc; j++)
                for (size_t idx = 0; idx < ((size_t)1 << (2 * (j + 1))); idx++) {
                    const uint32_t hb = (uint32_t)(idx >> (j + 1)), lb = (uint32_t)(idx & ((1u << (j + 1)) - 1u));
                    Quad &e4 = rq[(pi * kr + j) * nramp + idx];
                    for (int dom = 0; dom < 2; dom++) {
                        const int sec = dom ? pairs[pi].second : pairs[pi].first;
                        if (sec < 0) continue;
                        const double cv = m.lc_ramp[(((size_t)sec * kr + j) * 3 + 0) * nramp + idx];
                        double nv;
                        if (j < kn) nv = m.ln_ramp[((size_t)sec * knr + j) * nnramp + idx];
                        else nv = m.ln_full[(size_t)sec * nnon + ((((size_t)(hb >> (j - kn))) << (kn + 1)) | (lb >> (j - kn)))];
                        if (dom) { e4.cA = cv; e4.oA = cv - nv; } else { e4.cB = cv; e4.oB = cv - nv; }
                    }
                }
        {
            std::vector<double> pa(pairs.size() * 4, 0.0);
            for (size_t pi = 0; pi < pairs.size(); pi++)
                for (int dom = 0; dom < 2; dom++) {
                    const int sec = dom ? pairs[pi].second : pairs[pi].first;
                    if (sec < 0) continue;
                    pa[pi * 4 + dom] = m.sc[sec].ln_cpn;                          // ln COD_P_N
                    pa[pi * 4 + 2 + dom] = m.sc[sec].ln_cpn - m.sc[sec].ln_pn;    // ln NON_P_N
                }
            up(p.pair_amb, pa.data(), pa.size() * sizeof(double));
            d.pair_amb = p.pair_amb.as<double>();
        }
        {
            // upper bounds of what k_final adds to a start's walk sum, per pair and domain: order ramp (0 if ambiguous letters
            // silence it), start term, stop term.  A start whose bound stays <= 0 is pruned without touching the planes.
            std::vector<double> ub(pairs.size() * 8, 0.0);
            for (size_t pi = 0; pi < pairs.size(); pi++)
                for (int dom = 0; dom < 2; dom++) {
                    const int sec = dom ? pairs[pi].second : pairs[pi].first;
                    if (sec < 0) continue;
                    double ramp = 0.0;
                    for (int j = 0; j < kc; j++) {
                        double best = -1e300;
                        for (size_t idx = 0; idx < ((size_t)1 << (2 * (j + 1))); idx++) {
                            const Quad &e4 = rq[(pi * kr + j) * nramp + idx];
                            best = std::max(best, dom ? e4.oA : e4.oB);
                        }
                        ramp += std::max(best, 0.0);   // a short candidate may stop before position j
                    }
                    ub[pi * 8 + dom * 3 + 0] = ramp;
                    ub[pi * 8 + dom * 3 + 1] = std::max({m.sc[sec].start_lo[0], m.sc[sec].start_lo[1], m.sc[sec].start_lo[2]});
                    ub[pi * 8 + dom * 3 + 2] = std::max({m.sc[sec].stop_lo[0], m.sc[sec].stop_lo[1], m.sc[sec].stop_lo[2]});
                    // one full-order position at codon position 0 (added by k_final behind the pair walk), or its ambiguous-letter constant
                    double one = std::max(0.0, m.sc[sec].ln_pn);
                    for (size_t idx = 0; idx < nfull; idx++) {
                        const Quad &e4 = q[(pi * 3 + 0) * nfull + idx];
                        one = std::max(one, dom ? e4.oA : e4.oB);
                    }
                    ub[pi * 8 + 6 + dom] = one;
                }
            up(p.pair_ub, ub.data(), ub.size() * sizeof(double));
            d.pair_ub = p.pair_ub.as<double>();
        }
        up(p.rquad, rq.data(), std::max<size_t>(rq.size(), 1) * sizeof(Quad));
        d.rquad = p.rquad.as<Quad>();
        e->max_cod_order = std::max(e->max_cod_order, kc);
        e->max_pairs = std::max(e->max_pairs, d.npair);
        e->max_non_order = std::max(e->max_non_order, kn);
    }
    // the pair walk needs every model's pair tables in one SM's shared memory (orders <= 4); otherwise, or with
    // MGCOD_B200_WALK=single (A/B measurements, tests of the older kernel), the single-position walk runs
    e->pair_smem = 0;
    if (const char *cm = std::getenv("MGCOD_B200_CHAIN")) e->chain_mode = std::strcmp(cm, "warp") == 0 ? 1 : (std::strcmp(cm, "thread") == 0 ? 2 : (std::strcmp(cm, "hybrid") == 0 ? 0 : -1));
    {
        const char *mode = std::getenv("MGCOD_B200_WALK");
        bool all = !(mode && std::strcmp(mode, "single") == 0);
        int bytes = 0;
        for (const DevModel &d : dm) {
            all = all && d.rc2 >= 1;
            bytes = std::max(bytes, d.wtab2_n16 * 16);
        }
        if (all) e->pair_smem = bytes;
        e->orders_uniform = true;
        for (const DevModel &d : dm) e->orders_uniform = e->orders_uniform && d.kc == e->max_cod_order && d.kn == e->max_non_order;
    }
    e->d_models.ensure(sizeof(DevModel) * e->n_models);
    CUDA_TRY(cudaMemcpy(e->d_models.p, dm.data(), sizeof(DevModel) * e->n_models, cudaMemcpyHostToDevice));
}

void ensure_batch_buffers(mgc_engine *e) {
    const int M = e->n_models;
    const int64_t W = e->W, nt = e->total_nt;
    e->lo.ensure((size_t)(2 * W + 8) * 4);
    e->hi.ensure((size_t)(2 * W + 8) * 4);
    e->gc.ensure((size_t)(W + 8) * 4);
    e->nm.ensure((size_t)(2 * W + 8) * 4);
    e->codon.ensure((size_t)(10 * W + 8) * 4);
    e->ncount.ensure((size_t)(W + 8) * 4);
    const int64_t nb = W / 4;
    const int64_t scan_n = std::max<int64_t>(std::max<int64_t>(W, nb), e->n_contigs + 1);
    e->tile_sums.ensure((size_t)(scan_n / 2048 + 8) * 4 * 2 * std::max(M, 1));   // batched scans keep one row of tile sums per array
    e->grp2contig.ensure((size_t)(W / 4 + 8) * 4);
    e->flags.ensure(64);
    // observed on gene-dense input per kb and model: ~11 ORFs, ~36 starts, ~18 walk items, ~10 surviving candidates
    // cap_scale doubles when a batch turns out denser than that (low-complexity sequence such as ATG repeats): do_decode retries
    const int64_t sc = e->cap_scale;
    e->cap_orf = (uint32_t)std::min<int64_t>(nt / 24 * sc + 8192, 0x7fffff00);
    e->cap_starts = (uint32_t)std::min<int64_t>(nt / 8 * sc + 8192, 0x7fffff00);
    e->cap_items = (uint32_t)std::min<int64_t>(nt / 16 * sc + 8192, 0x7fffff00);
    e->cap_surv = (uint32_t)std::min<int64_t>(nt / 24 * sc + 8192, 0x7fffff00);
    std::vector<void *> t_orfs(M), t_starts(M), t_surv(M), t_sorted(M), t_bcount(M), t_cover(M), t_gene_off(M), t_genes(M);
    std::vector<void *> t_scan_buckets(2 * M), t_gene_count(M), t_slot_gene_off(M);
    std::vector<void *> t_ls(M), t_ls_tile(M), t_items(M), t_items_sorted(M), t_hist(M), t_item_off(M), t_unit_off(M), t_sums(M);
    const int64_t ls_tiles = (2 * W + 255) / 256;
    std::vector<ChainBufs> cb(M);
    for (int i = 0; i < M; i++) {
        PerModel &p = e->pm[i];
        p.orfs.ensure((size_t)e->cap_orf * sizeof(OrfRec));
        p.starts.ensure((size_t)e->cap_starts * sizeof(StartRec));
        p.sums.ensure((size_t)e->cap_starts * 32);
        p.items.ensure((size_t)e->cap_items * sizeof(ItemRec));
        p.items_sorted.ensure((size_t)e->cap_items * sizeof(ItemRec));
        const size_t nkeys = (size_t)102 * WALK_CLASSES * WALK_TILES_MAX;   // at most 101 distinct section pairs, per position tile
        p.hist.ensure((nkeys + 8) * 4);
        p.item_off.ensure((nkeys + 8) * 4);
        p.unit_off.ensure((nkeys + 8) * 4);
        p.ls.ensure((size_t)(2 * W + 8) * 3 * 4);
        p.ls_tile.ensure((size_t)(ls_tiles + 8) * 3 * 4);
        t_ls[i] = p.ls.p; t_ls_tile[i] = p.ls_tile.p; t_items[i] = p.items.p; t_items_sorted[i] = p.items_sorted.p;
        t_hist[i] = p.hist.p; t_item_off[i] = p.item_off.p; t_unit_off[i] = p.unit_off.p; t_sums[i] = p.sums.p;
        p.surv.ensure((size_t)e->cap_surv * sizeof(Cand));
        p.sorted.ensure((size_t)e->cap_surv * sizeof(Cand));
        p.bcount.ensure((size_t)(nb + 8) * 4);
        p.cover.ensure((size_t)(nb + 8) * 4);
        p.minleft.ensure((size_t)(nb + 8) * 4);
        p.cum.ensure((size_t)e->cap_surv * 8);
        p.prev.ensure((size_t)e->cap_surv * 4);
        p.gptr.ensure((size_t)e->cap_orf * 4);
        p.av_parg.ensure((size_t)e->cap_surv * 4);
        if (e->chain_mode != 2) {
            for (DevBuf *b : {&p.av_right, &p.av_left, &p.av_ls, &p.av_idx}) b->ensure((size_t)e->cap_surv * 4);
            for (DevBuf *b : {&p.av_cum, &p.av_pmax}) b->ensure((size_t)e->cap_surv * 8);
        }
        if (e->chain_mode != 1) {
            p.st.ensure((size_t)e->cap_surv * sizeof(ChainSt));
            p.st_rl.ensure(((size_t)e->cap_surv + 8) * 8);   // the scan reads aligned groups of four: up to three records past the last
        }
        p.chain.ensure((size_t)e->cap_surv * 4);
        p.gene_count.ensure((size_t)(e->n_slots + 8) * 4);
        p.slot_gene_off.ensure((size_t)(e->n_slots + 8) * 4);
        p.piece_b0.ensure((size_t)(e->n_slots + 8) * 4);
        p.piece_b1.ensure((size_t)(e->n_slots + 8) * 4);
        p.work.ensure((size_t)(e->n_slots + 8) * 4);
        p.piece_total.ensure((size_t)(e->n_slots + 8) * 8);
        p.gene_off.ensure((size_t)(e->n_contigs + 8) * 4);
        p.contig_total.ensure((size_t)(e->n_contigs + 8) * 8);
        p.genes.ensure((size_t)e->cap_surv * sizeof(mgc_gene));
        t_orfs[i] = p.orfs.p; t_starts[i] = p.starts.p; t_surv[i] = p.surv.p; t_sorted[i] = p.sorted.p;
        t_bcount[i] = p.bcount.p; t_cover[i] = p.cover.p; t_gene_off[i] = p.gene_off.p; t_genes[i] = p.genes.p;
        t_scan_buckets[i] = p.bcount.p; t_scan_buckets[M + i] = p.cover.p; t_gene_count[i] = p.gene_count.p; t_slot_gene_off[i] = p.slot_gene_off.p;
        ChainBufs &c = cb[i];
        c.sorted = p.surv.as<Cand>();   // k_final's survivors are scattered into p.sorted by bucket, then ordered inside the buckets back into p.surv
        c.bend = p.bcount.as<uint32_t>();
        c.cover = p.cover.as<uint32_t>();
        c.minleft = p.minleft.as<uint32_t>();
        c.cum = p.cum.as<double>();
        c.prev = p.prev.as<int32_t>();
        c.gptr = p.gptr.as<int32_t>();
        c.av_right = p.av_right.as<int32_t>();
        c.av_left = p.av_left.as<int32_t>();
        c.av_ls = p.av_ls.as<int32_t>();
        c.av_idx = p.av_idx.as<int32_t>();
        c.av_cum = p.av_cum.as<double>();
        c.av_pmax = p.av_pmax.as<double>();
        c.av_parg = p.av_parg.as<int32_t>();
        c.st = p.st.as<ChainSt>();
        c.st_rl = p.st_rl.as<int2>();
        c.chain = p.chain.as<int32_t>();
        c.piece_b0 = p.piece_b0.as<int32_t>();
        c.piece_b1 = p.piece_b1.as<int32_t>();
        c.work = p.work.as<int32_t>();
        c.gene_count = p.gene_count.as<int32_t>();
        c.piece_total = p.piece_total.as<double>();
        c.slot_gene_off = p.slot_gene_off.as<uint32_t>();
        c.gene_off = p.gene_off.as<uint32_t>();
        c.contig_total = p.contig_total.as<double>();
    }
    // the pointer tables only change when a buffer was reallocated: skip their upload for a batch that fits
    std::vector<unsigned char> sig;
    auto add_sig = [&](const void *src, size_t bytes) {
        sig.insert(sig.end(), static_cast<const unsigned char *>(src), static_cast<const unsigned char *>(src) + bytes);
    };
    for (auto *v : {&t_gene_count, &t_slot_gene_off, &t_orfs, &t_starts, &t_surv, &t_sorted, &t_bcount, &t_cover, &t_gene_off, &t_genes, &t_ls, &t_ls_tile, &t_items,
                    &t_items_sorted, &t_hist, &t_item_off, &t_unit_off, &t_sums})
        add_sig(v->data(), M * sizeof(void *));
    add_sig(cb.data(), M * sizeof(ChainBufs));
    add_sig(t_scan_buckets.data(), 2 * M * sizeof(void *));
    const bool tables_current = sig == e->table_sig;
    auto up = [&](DevBuf &b, const void *src, size_t bytes) {
        if (tables_current) return;
        b.ensure(bytes);
        CUDA_TRY(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    };
    up(e->t_orfs, t_orfs.data(), M * sizeof(void *));
    up(e->t_starts, t_starts.data(), M * sizeof(void *));
    up(e->t_surv, t_surv.data(), M * sizeof(void *));
    up(e->t_sorted, t_sorted.data(), M * sizeof(void *));
    up(e->t_bcount, t_bcount.data(), M * sizeof(void *));
    up(e->t_cover, t_cover.data(), M * sizeof(void *));
    up(e->t_scan_buckets, t_scan_buckets.data(), 2 * M * sizeof(void *));
    up(e->t_gene_count, t_gene_count.data(), M * sizeof(void *));
    up(e->t_slot_gene_off, t_slot_gene_off.data(), M * sizeof(void *));
    up(e->t_gene_off, t_gene_off.data(), M * sizeof(void *));
    up(e->t_genes, t_genes.data(), M * sizeof(void *));
    up(e->t_chainbufs, cb.data(), M * sizeof(ChainBufs));
    up(e->t_ls, t_ls.data(), M * sizeof(void *));
    up(e->t_ls_tile, t_ls_tile.data(), M * sizeof(void *));
    up(e->t_items, t_items.data(), M * sizeof(void *));
    up(e->t_items_sorted, t_items_sorted.data(), M * sizeof(void *));
    up(e->t_hist, t_hist.data(), M * sizeof(void *));
    up(e->t_item_off, t_item_off.data(), M * sizeof(void *));
    up(e->t_unit_off, t_unit_off.data(), M * sizeof(void *));
    up(e->t_sums, t_sums.data(), M * sizeof(void *));
    e->table_sig.swap(sig);
    e->unit_totals.ensure((size_t)M * 4 + 16);
    e->next_unit.ensure(16);
    e->counters.ensure((size_t)M * 8 * 8);
    e->surv_counters.ensure((size_t)M * 8);
    e->work_counters.ensure((size_t)M * 16);
    e->model_cents.ensure((size_t)M * 8);
    e->gene_totals.ensure((size_t)M * 4);
}

// (re)builds the slot table from e->segs and uploads it; the ASCII letters stay where they are
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// Up to this many bp x models in a batch the long chain pieces go to the warp-per-piece kernel (measured: 125 Mbp x 4 chain phase
// 3.0 -> 2.2 ms, 1 Gbp x 4 7.5 -> 10.6 ms: there the threads have several pieces each and the long ones, started first, hide)
constexpr int64_t kChainHybridMaxUnits = 1200000000;
constexpr int64_t kChainWarpMaxUnits = 600000000;    // up to 150 Mbp x 4: all pieces to the warp-per-piece kernel (20 Mbp x 4: chain 1.36 -> 0.69 ms, 125 Mbp x 4: 2.19 -> 2.02)
// the caller's contigs as slots, one each (gaps re-slot later: do_decode)
static void segs_from_offsets(const int64_t *offsets, int32_t n, std::vector<mgc_engine::Seg> &segs, std::vector<int32_t> &parent_len,
                              std::vector<int64_t> &parent_aoff) {
    parent_len.resize(n);
    parent_aoff.resize(n + 1);
    segs.resize(n);
    const int64_t a0 = offsets[0], a1 = offsets[n];
    for (int32_t c = 0; c < n; c++) {
        const int64_t L = offsets[c + 1] - offsets[c];
        if (L < 0 || L > 0x7ffffff0) throw std::invalid_argument("contig length out of range");
        parent_len[c] = (int32_t)L;
        parent_aoff[c] = offsets[c] - a0;
        segs[c] = mgc_engine::Seg{c, offsets[c] - a0, (int32_t)L, 0, 1};
    }
    parent_aoff[n] = a1 - a0;
}

// host-only half of set_slots: word offsets of the slots, chain piece slots, totals
static void compute_slot_tables(const std::vector<mgc_engine::Seg> &segs, mgc_engine::SlotTables &T) {
    const int32_t n = (int32_t)segs.size();
    T.n = n;
    T.clen.resize(n);
    T.cwoff.resize(n + 1);
    T.aoff.resize(n + 1);
    T.cflags.resize(n);
    T.slotbase.resize(n + 1);
    T.slot_contig.clear();
    int64_t w = 0, nt = 0;
    for (int32_t c = 0; c < n; c++) {
        const int64_t L = segs[c].len;
        T.slotbase[c] = (int32_t)T.slot_contig.size();
        const int64_t nbk = ((L > 0 ? L - 1 : 0) >> 7) + 1;
        T.slot_contig.insert(T.slot_contig.end(), (size_t)((nbk + PIECE_BUCKETS - 1) / PIECE_BUCKETS), c);
        T.clen[c] = (int32_t)L;
        T.cwoff[c] = w;
        T.aoff[c] = segs[c].aoff;
        T.cflags[c] = segs[c].right_end;
        const int64_t words = (L + 31) / 32 + 1;
        w += (words + 3) / 4 * 4;
        nt += L;
    }
    T.slotbase[n] = (int32_t)T.slot_contig.size();
    T.n_slots = T.slotbase[n];
    T.cwoff[n] = w;
    T.aoff[n] = 0;
    if (w * 64 >= 0xffffffffLL) throw std::invalid_argument("batch too large: at most ~2e9 nt per decode call");
    T.W = w;
    T.total_nt = nt;
}

// device half: buffers for a batch of this size, the six tables through the page-locked staging block
static void apply_slot_tables(mgc_engine *e, mgc_engine::SlotTables &T) {
    const bool trace = std::getenv("MGCOD_TRACE") != nullptr;
    const double t1 = now_ms();
    const int32_t n = T.n;
    e->n_contigs = n;
    e->h_clen.swap(T.clen);
    e->h_cwoff.swap(T.cwoff);
    e->h_aoff.swap(T.aoff);
    e->n_slots = T.n_slots;
    e->W = T.W;
    e->total_nt = T.total_nt;
    ensure_batch_buffers(e);
    const double t2 = now_ms();
    e->aoff.ensure((size_t)(n + 1) * 8);
    e->cwoff.ensure((size_t)(n + 1) * 8);
    e->clen.ensure((size_t)n * 4);
    e->cflags.ensure((size_t)n + 16);
    e->slotbase.ensure((size_t)(n + 1) * 4);
    e->slot_contig.ensure((size_t)e->n_slots * 4 + 16);
    cudaStream_t st = e->stream;
    // the six tables go through one page-locked staging block kept by the engine: a copy from pageable memory waits for the
    // previous result's device-to-host copy (2 ms per step of a stream of batches), a page-locked one is queued at once
    {
        struct Part { void *dst; const void *src; size_t bytes; };
        const Part parts[6] = {{e->slotbase.p, T.slotbase.data(), (size_t)(n + 1) * 4}, {e->slot_contig.p, T.slot_contig.data(), (size_t)e->n_slots * 4},
                               {e->aoff.p, e->h_aoff.data(), (size_t)(n + 1) * 8},       {e->cwoff.p, e->h_cwoff.data(), (size_t)(n + 1) * 8},
                               {e->clen.p, e->h_clen.data(), (size_t)n * 4},             {e->cflags.p, T.cflags.data(), (size_t)n}};
        size_t total = 0;
        for (const Part &q : parts) total += (q.bytes + 63) / 64 * 64;
        if (e->table_stage.bytes < total) {
            e->pool->release(e->table_stage);
            e->table_stage = e->pool->acquire(total);
        }
        size_t at = 0;
        for (const Part &q : parts) {
            if (!q.bytes) continue;
            std::memcpy(static_cast<char *>(e->table_stage.p) + at, q.src, q.bytes);
            CUDA_TRY(cudaMemcpyAsync(q.dst, static_cast<char *>(e->table_stage.p) + at, q.bytes, cudaMemcpyHostToDevice, st));
            at += (q.bytes + 63) / 64 * 64;
        }
    }
    launch_groups(e->cwoff.as<int64_t>(), n, e->W / 4, e->grp2contig.as<int32_t>(), st);
    const double t3 = now_ms();
    CUDA_TRY(cudaStreamSynchronize(st));  // the staging block is reused by the next batch
    CUDA_TRY(cudaGetLastError());
    if (trace) std::fprintf(stderr, "slot tables applied: ensure %.2f ms, uploads %.2f ms, sync %.2f ms\n", t2 - t1, t3 - t2, now_ms() - t3);
}

// (re)builds the slot table from e->segs and uploads it; the ASCII letters stay where they are
void set_slots(mgc_engine *e) {
    mgc_engine::SlotTables T;
    compute_slot_tables(e->segs, T);
    apply_slot_tables(e, T);
}

void do_upload(mgc_engine *e, const char *seq, const int64_t *offsets, int32_t n) {
    if (n <= 0) throw std::invalid_argument("n_contigs must be positive");
    CUDA_TRY(cudaSetDevice(e->device));
    const bool trace = std::getenv("MGCOD_TRACE") != nullptr;
    const double tu0 = now_ms();
    e->have_batch = e->have_genes = false;
    e->segmented = false;
    e->n_parents = n;
    const bool adopted = e->pending.valid && e->pending.seq == seq && (int32_t)e->pending.offsets.size() == n + 1 &&
                         std::memcmp(e->pending.offsets.data(), offsets, (size_t)(n + 1) * 8) == 0;
    e->pending.valid = false;
    // the host tables of an adopted batch were prepared by a background thread while the previous batch was decoded
    bool prepared = false;
    if (e->prepared.valid) {
        e->prepared.valid = false;
        e->prepared.ready.get();   // rethrows what the preparation threw (contig too long, batch too large)
        prepared = adopted;
    }
    if (prepared) {
        e->segs.swap(e->prepared.segs);
        e->h_parent_len.swap(e->prepared.parent_len);
        e->h_parent_aoff.swap(e->prepared.parent_aoff);
    } else {
        segs_from_offsets(offsets, n, e->segs, e->h_parent_len, e->h_parent_aoff);
    }
    const int64_t a0 = offsets[0], a1 = offsets[n];
    cudaStream_t st = e->stream;
    e->plen.ensure((size_t)n * 4 + 16);
    {   // a kernel first: an event recorded as the first operation after the last step's read-back copies would wait for the
        // previous result's device-to-host copy (see run_pipeline)
        ClearList cl;
        e->flags.ensure(64);
        cl.add(e->flags.p, 64);
        launch_clear(cl, 1, st);
    }
    CUDA_TRY(cudaEventRecord(e->ev[0], st));
    if (adopted) {  // the letters are already on their way (or there): take the buffer over
        std::swap(e->ascii, e->ascii_next);
        CUDA_TRY(cudaStreamWaitEvent(st, e->copy_done, 0));
    } else {
        CUDA_TRY(cudaStreamSynchronize(e->copy_stream));  // a prefetch nobody adopted must not outlive its host buffer
        e->ascii.ensure((size_t)(a1 - a0) + 2 * PACK_SLACK);
        CUDA_TRY(cudaMemcpyAsync(e->ascii.as<char>() + PACK_SLACK, seq + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, st));
    }
    const double tu1 = now_ms();
    CUDA_TRY(cudaMemcpyAsync(e->plen.p, e->h_parent_len.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    const double tu2 = now_ms();
    if (prepared) apply_slot_tables(e, e->prepared.tables);
    else set_slots(e);
    const double tu3 = now_ms();
    CUDA_TRY(cudaEventRecord(e->ev[1], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (trace) std::fprintf(stderr, "do_upload: host tables + letters queued %.2f ms, lengths queued %.2f ms, set_slots %.2f ms, sync %.2f ms (adopted %d; set_slots called %.3f, back %.3f)\n",
                            tu1 - tu0, tu2 - tu1, tu3 - tu2, now_ms() - tu3, (int)adopted, tu2, tu3);
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
    e->stats = mgc_stats();
    e->stats.h2d = ms;
    e->have_batch = true;
}

// one pass of the device pipeline over the current slots; *amb_out tells whether any letter was ambiguous
mgc_result *run_pipeline(mgc_engine *e, bool *amb_out) {
    if (!e->have_batch) throw std::invalid_argument("no batch uploaded");
    CUDA_TRY(cudaSetDevice(e->device));
    const int M = e->n_models;
    const int32_t n = e->n_contigs;
    const int64_t W = e->W, nb = W / 4;
    cudaStream_t st = e->stream;
    int launches = 0;
    const bool trace = std::getenv("MGCOD_TRACE") != nullptr;
    const double tp0 = now_ms();
    float h2d = e->stats.h2d;
    e->stats = mgc_stats();
    e->stats.h2d = h2d;
    e->have_genes = false;

    {
        ClearList cl;   // one kernel instead of 30 cudaMemsetAsync calls (which run on a copy engine, behind the previous result's read-back)
        cl.add(e->flags.p, 64);
        cl.add(e->counters.p, (size_t)M * 64);
        cl.add(e->surv_counters.p, (size_t)M * 8);
        cl.add(e->work_counters.p, (size_t)M * 16);
        cl.add(e->model_cents.p, (size_t)M * 8);
        cl.add(e->next_unit.p, 16);
        for (int i = 0; i < M; i++) {
            cl.add(e->pm[i].gene_count.p, (size_t)(e->n_slots + 1) * 4);
            cl.add(e->pm[i].piece_total.p, (size_t)(e->n_slots + 1) * 8);
            cl.add(e->pm[i].piece_b1.p, (size_t)(e->n_slots + 1) * 4, 0xffffffffu);  // -1 = slot not used
            cl.add(e->pm[i].bcount.p, (size_t)(nb + 8) * 4);
            cl.add(e->pm[i].cover.p, (size_t)(nb + 8) * 4);
            cl.add(e->pm[i].hist.p, e->pm[i].hist.bytes);
        }
        if (cl.n >= ClearList::kMax) throw std::runtime_error("clear list too small");
        launch_clear(cl, e->n_sm, st);
        launches += 1;
    }
    // The step's first event is recorded AFTER its first kernel.  The stream's previous operation was the small read-back copy of
    // the last step, and an event record on an otherwise idle stream is queued to the engine the stream used last — the
    // device-to-host copy engine, busy for 3 ms with the previous result's gene tables on the copy-out stream: every kernel of the
    // step then started 3 ms late (measured with cudaEventQuery polling; a clear of 260 MB takes 0.1 ms and is not in `pack`).
    CUDA_TRY(cudaEventRecord(e->ev[0], st));
    // pack + GC prefix
    launch_pack(e->ascii.as<char>() + PACK_SLACK, e->aoff.as<int64_t>(), e->cwoff.as<int64_t>(), e->clen.as<int32_t>(),
                e->grp2contig.as<int32_t>(), W, e->lo.as<uint32_t>(), e->hi.as<uint32_t>(), e->nm.as<uint32_t>(), e->gc.as<uint32_t>(),
                e->ncount.as<uint32_t>(), e->flags.as<int>(), st);
    launches += 1;
    launches += launch_exclusive_scan(e->gc.as<uint32_t>(), e->gc.as<uint32_t>(), e->tile_sums.as<uint32_t>(), W, nullptr, st);
    launches += launch_exclusive_scan(e->ncount.as<uint32_t>(), e->ncount.as<uint32_t>(), e->tile_sums.as<uint32_t>(), W, nullptr, st);
    CUDA_TRY(cudaPeekAtLastError());   // a failed launch must not go unnoticed until the end
    CUDA_TRY(cudaEventRecord(e->ev[1], st));
    Planes pl;
    pl.lo = e->lo.as<uint32_t>();
    pl.hi = e->hi.as<uint32_t>();
    pl.nm = e->nm.as<uint32_t>();
    pl.npre = e->ncount.as<uint32_t>();
    pl.dflags = e->flags.as<int>();
    pl.cflags = e->cflags.as<uint8_t>();
    pl.codon = e->codon.as<uint32_t>();
    pl.W = W;
    { const WalkTiles wt = walk_tiles_for(W, walk_tile_words_setting()); pl.n_tiles = wt.n; pl.tile_words = wt.words; }
    launch_codon_planes(pl, e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), e->grp2contig.as<int32_t>(), e->codon.as<uint32_t>(), st);
    launches += 1;
    launches += launch_laststop(pl, e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), e->grp2contig.as<int32_t>(),
                                e->d_models.as<DevModel>(), M, e->t_ls.as<uint32_t *>(), e->t_ls_tile.as<uint32_t *>(), st);
    launch_orf_find(pl, e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), e->grp2contig.as<int32_t>(), e->d_models.as<DevModel>(), M,
                    e->t_ls.as<const uint32_t *>(), e->t_ls_tile.as<const uint32_t *>(), e->t_orfs.as<OrfRec *>(),
                    e->counters.as<unsigned long long>(), e->cap_orf, st);
    launches += 1;
    CUDA_TRY(cudaPeekAtLastError());   // a failed launch must not go unnoticed until the end
    CUDA_TRY(cudaEventRecord(e->ev[2], st));
    launches += launch_orf_prep(pl, e->gc.as<uint32_t>(), e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), e->d_models.as<DevModel>(), M,
                                e->t_orfs.as<OrfRec *>(), e->t_starts.as<StartRec *>(), e->t_items.as<ItemRec *>(), e->t_hist.as<uint32_t *>(),
                                e->counters.as<unsigned long long>(), e->t_chainbufs.as<ChainBufs>(), e->cap_orf, e->cap_starts, e->cap_items,
                                e->n_sm, st);
    launches += launch_item_sort(e->t_items.as<const ItemRec *>(), e->t_items_sorted.as<ItemRec *>(), e->t_hist.as<uint32_t *>(),
                                 e->t_item_off.as<uint32_t *>(), e->t_unit_off.as<uint32_t *>(), e->d_models.as<DevModel>(), M,
                                 e->unit_totals.as<uint32_t>(), e->counters.as<unsigned long long>(), e->cap_items, pl.n_tiles, e->n_sm,
                                 e->pair_smem > 0, st);
    CUDA_TRY(cudaPeekAtLastError());
    CUDA_TRY(cudaPeekAtLastError());   // a failed launch must not go unnoticed until the end
    CUDA_TRY(cudaEventRecord(e->ev[8], st));
    launches += launch_walk(pl, e->cwoff.as<int64_t>(), e->d_models.as<DevModel>(), M, e->max_cod_order, e->max_non_order,
                            e->t_orfs.as<const OrfRec *>(),
                            e->t_starts.as<const StartRec *>(), e->t_items_sorted.as<const ItemRec *>(),
                            e->t_item_off.as<const uint32_t *>(), e->t_unit_off.as<const uint32_t *>(), e->unit_totals.as<uint32_t>(),
                            e->t_sums.as<void *>(), e->next_unit.as<unsigned int>(), e->n_sm, e->pair_smem, e->orders_uniform, st);
    CUDA_TRY(cudaPeekAtLastError());   // a failed launch must not go unnoticed until the end
    CUDA_TRY(cudaEventRecord(e->ev[9], st));
    launch_final(pl, e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), e->d_models.as<DevModel>(), M, e->t_orfs.as<const OrfRec *>(),
                 e->t_starts.as<const StartRec *>(), e->t_sums.as<void *>(), e->counters.as<unsigned long long>(), e->t_surv.as<Cand *>(),
                 e->t_bcount.as<uint32_t *>(), e->t_cover.as<uint32_t *>(), e->surv_counters.as<unsigned long long>(), e->cap_starts, e->cap_surv,
                 e->n_sm, e->pair_smem > 0, st);
    launches += 1;
    CUDA_TRY(cudaPeekAtLastError());   // a failed launch must not go unnoticed until the end
    CUDA_TRY(cudaEventRecord(e->ev[3], st));
    // bucket offsets (M arrays) and boundary coverage (M arrays) of all models in one batched scan, in place; the two extra
    // entries are zero.  Grand totals: the first M = candidates per model (survivor slots may include unfilled ones), written
    // to the low half of counter 4; the coverage totals land in counter 6 and are not used.
    launches += launch_exclusive_scan_multi(e->t_scan_buckets.as<const uint32_t *>(), e->t_scan_buckets.as<uint32_t *>(), 2 * M,
                                            e->tile_sums.as<uint32_t>(), nb + 2, nullptr, 0, st);
    launch_bucket_totals(e->t_bcount.as<const uint32_t *const>(), nb, e->counters.as<unsigned long long>(), M, st);
    launches += 1;
    launch_scatter(e->t_surv.as<const Cand *>(), e->t_sorted.as<Cand *>(), e->t_bcount.as<uint32_t *>(),
                   e->surv_counters.as<unsigned long long>(), e->cwoff.as<int64_t>(), e->cap_surv, M, e->n_sm, st);
    launches += 1;
    CUDA_TRY(cudaPeekAtLastError());   // a failed launch must not go unnoticed until the end
    CUDA_TRY(cudaEventRecord(e->ev[4], st));
    launches += launch_chain(e->t_chainbufs.as<ChainBufs>(), e->t_sorted.as<const Cand *>(), e->t_surv.as<Cand *>(), e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), e->slotbase.as<int32_t>(),
                             e->slot_contig.as<int32_t>(), e->n_slots, nb, e->work_counters.as<int>(), M, e->n_sm,
                             // small batches (the per-rank shares of one input over several GPUs): long pieces by warps, see launch_chain;
                             // a single genome: every piece by a warp (a thread needs ~9 us per candidate, phage2's longest piece 0.6 ms)
                             e->chain_mode >= 0 ? e->chain_mode
                                                : (e->total_nt * M > kChainHybridMaxUnits ? 2 : (e->total_nt * M <= kChainWarpMaxUnits ? 1 : 0)),
                             st);
    CUDA_TRY(cudaPeekAtLastError());   // a failed launch must not go unnoticed until the end
    CUDA_TRY(cudaEventRecord(e->ev[5], st));
    if (std::getenv("MGCOD_B200_PIECE_STATS")) {   // diagnostic: sizes of the chain pieces of model 0
        unsigned int *d_hist = nullptr, h_hist[40];
        CUDA_TRY(cudaMalloc(&d_hist, sizeof(h_hist)));
        CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(h_hist), st));
        launch_piece_stats(e->t_chainbufs.as<ChainBufs>(), e->work_counters.as<int>(), e->n_slots, d_hist, st);
        CUDA_TRY(cudaMemcpyAsync(h_hist, d_hist, sizeof(h_hist), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(d_hist);
        std::fprintf(stderr, "[mgcod_b200] chain pieces of model 0: %u, largest %u candidates; log2 histogram:", h_hist[33], h_hist[32]);
        for (int i = 0; i < 32; i++) if (h_hist[i]) std::fprintf(stderr, " [<2^%d]=%u", i, h_hist[i]);
        std::fprintf(stderr, "\n");
    }
    // gene_count[n_slots] stays 0 so that the exclusive scan yields the total at slot_gene_off[n_slots]
    launches += launch_exclusive_scan_multi(e->t_gene_count.as<const uint32_t *>(), e->t_slot_gene_off.as<uint32_t *>(), M,
                                            e->tile_sums.as<uint32_t>(), (int64_t)e->n_slots + 1, e->gene_totals.as<uint32_t>(), 1, st);
    if (e->last_ready) CUDA_TRY(cudaStreamWaitEvent(st, e->last_ready->ev, 0));   // the previous result still reads the gene buffers
    launches += launch_gather(e->t_chainbufs.as<ChainBufs>(), e->t_genes.as<mgc_gene *>(), e->slot_contig.as<int32_t>(), e->n_slots,
                              e->slotbase.as<int32_t>(), n, e->model_cents.as<long long>(), M, e->n_sm, st);
    CUDA_TRY(cudaPeekAtLastError());   // a failed launch must not go unnoticed until the end
    CUDA_TRY(cudaEventRecord(e->ev[6], st));

    // small read-backs
    const double tp1 = now_ms();
    double tq0 = 0, tq6 = 0;
    if (trace) {   // when do the first and the last event of the step fire (host clock)?
        while (cudaEventQuery(e->ev[0]) == cudaErrorNotReady) {}
        tq0 = now_ms();
        while (cudaEventQuery(e->ev[6]) == cudaErrorNotReady) {}
        tq6 = now_ms();
    }
    std::vector<unsigned long long> h_counters((size_t)M * 8), h_surv(M);
    std::vector<uint32_t> h_gtot(M);
    std::vector<long long> h_cents(M);
    int h_flags[16];
    CUDA_TRY(cudaMemcpyAsync(h_counters.data(), e->counters.p, (size_t)M * 64, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_surv.data(), e->surv_counters.p, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_gtot.data(), e->gene_totals.p, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_cents.data(), e->model_cents.p, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_flags, e->flags.p, 64, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    const double tp2 = now_ms();
    if (h_flags[0]) throw AlphabetError("error, unexpected letter found in FASTA sequence (accepted: ACGTU and the IUPAC codes RYSWKMBDHVN)");
    *amb_out = h_flags[1] != 0;
    for (int i = 0; i < M; i++) {
        if (h_counters[(size_t)i * 8] > e->cap_orf || h_counters[(size_t)i * 8 + 1] > e->cap_starts ||
            h_counters[(size_t)i * 8 + 2] > e->cap_items || h_surv[i] > e->cap_surv)
            throw mgc::ScratchExhausted("device scratch exhausted (ORFs " + std::to_string(h_counters[(size_t)i * 8]) + ", starts " +
                                     std::to_string(h_counters[(size_t)i * 8 + 1]) + ", candidates " + std::to_string(h_surv[i]) + ")");
        e->stats.n_orfs += (int64_t)h_counters[(size_t)i * 8];
        e->stats.n_starts += (int64_t)h_counters[(size_t)i * 8 + 1];
        e->stats.n_walks += (int64_t)h_counters[(size_t)i * 8 + 5];
        e->stats.n_candidates += (int64_t)(h_counters[(size_t)i * 8 + 4] & 0xffffffffull);
        e->stats.n_genes += h_gtot[i];
    }
    // genes into pooled page-locked blocks
    std::unique_ptr<mgc_result> res(new mgc_result());
    res->n_models = M;
    res->n_contigs = n;
    res->pool = e->pool;
    res->cents.assign(h_cents.begin(), h_cents.end());
    res->n_genes.assign(h_gtot.begin(), h_gtot.end());
    for (int i = 0; i < M; i++) res->min_len.push_back(e->models[i].min_len);
    // genes and totals travel on the copy-out stream; the caller may start the next decode meanwhile (readers of the result wait)
    CUDA_TRY(cudaEventRecord(e->main_done, st));
    CUDA_TRY(cudaStreamWaitEvent(e->d2h_stream, e->main_done, 0));
    CUDA_TRY(cudaEventRecord(e->d2h_begin, e->d2h_stream));
    for (int i = 0; i < M; i++) {
        res->gene_blk.push_back(e->pool->acquire((size_t)h_gtot[i] * sizeof(mgc_gene)));
        res->total_blk.push_back(e->pool->acquire((size_t)n * 8));
        if (h_gtot[i])
            CUDA_TRY(cudaMemcpyAsync(res->gene_blk[i].p, e->pm[i].genes.p, (size_t)h_gtot[i] * sizeof(mgc_gene), cudaMemcpyDeviceToHost,
                                     e->d2h_stream));
        CUDA_TRY(cudaMemcpyAsync(res->total_blk[i].p, e->pm[i].contig_total.p, (size_t)n * 8, cudaMemcpyDeviceToHost, e->d2h_stream));
    }
    CUDA_TRY(cudaEventRecord(e->d2h_end, e->d2h_stream));
    res->ready = std::make_shared<Ready>();
    CUDA_TRY(cudaEventRecord(res->ready->ev, e->d2h_stream));
    e->last_ready = res->ready;
    e->d2h_timed = false;
    float ms;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]); e->stats.pack = ms;
    cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]); e->stats.orf_find = ms;
    cudaEventElapsedTime(&ms, e->ev[2], e->ev[8]); e->stats.orf_prep = ms;
    cudaEventElapsedTime(&ms, e->ev[8], e->ev[9]); e->stats.orf_score = ms;
    cudaEventElapsedTime(&ms, e->ev[9], e->ev[3]); e->stats.orf_final = ms;
    cudaEventElapsedTime(&ms, e->ev[3], e->ev[4]); e->stats.order = ms;
    cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]); e->stats.chain = ms;
    cudaEventElapsedTime(&ms, e->ev[5], e->ev[6]); e->stats.aggregate = ms;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[6]); e->stats.total = ms;   // + d2h once it is known (mgc_engine_stats)
    e->stats.d2h = 0;
    e->stats.kernel_launches = launches;
    e->have_genes = true;
    e->dev_gene_count.assign(h_gtot.begin(), h_gtot.end());
    if (trace) std::fprintf(stderr, "run_pipeline: entered %.3f, all launched +%.2f ms, device done +%.2f ms, result queued +%.2f ms, first event +%.2f ms, last event +%.2f ms (device %.2f ms: pack %.2f find %.2f prep %.2f walk %.2f final %.2f order %.2f chain %.2f aggregate %.2f)\n",
                            tp0, tp1 - tp0, tp2 - tp0, now_ms() - tp0, tq0 - tp0, tq6 - tp0, e->stats.total, e->stats.pack, e->stats.orf_find, e->stats.orf_prep, e->stats.orf_score,
                            e->stats.orf_final, e->stats.order, e->stats.chain, e->stats.aggregate);
    return res.release();
}

// slot-space result -> the caller's contigs: contig index, coordinates, per-contig totals; re-uploads the mapped genes so
// that the window kernel sees parents
void remap_to_parents(mgc_engine *e, mgc_result *res) {
    res->wait();
    const int M = e->n_models;
    const int32_t np = e->n_parents;
    cudaStream_t st = e->stream;
    for (int i = 0; i < M; i++) {
        mgc_gene *g = static_cast<mgc_gene *>(res->gene_blk[i].p);
        const int64_t ng = res->n_genes[i];
        std::vector<uint32_t> off(np + 1, 0);
        for (int64_t k = 0; k < ng; k++) {
            const mgc_engine::Seg &s = e->segs[g[k].contig];
            g[k].contig = s.parent;
            g[k].left += s.poff;
            g[k].right += s.poff;
            off[s.parent + 1]++;
        }
        for (int32_t c = 0; c < np; c++) off[c + 1] += off[c];
        const double *tot = static_cast<const double *>(res->total_blk[i].p);
        PinnedPool::Block nb = e->pool->acquire((size_t)std::max(np, 1) * 8);
        double *pt = static_cast<double *>(nb.p);
        for (int32_t c = 0; c < np; c++) pt[c] = 0.0;
        for (size_t c = 0; c < e->segs.size(); c++) pt[e->segs[c].parent] += tot[c];
        e->pool->release(res->total_blk[i]);
        res->total_blk[i] = nb;
        if (ng) CUDA_TRY(cudaMemcpyAsync(e->pm[i].genes.p, g, (size_t)ng * sizeof(mgc_gene), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->pm[i].gene_off.p, off.data(), (size_t)(np + 1) * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    res->n_contigs = np;
}

// run_pipeline with the scratch capacities doubled (up to 32x) when the batch holds more ORFs / starts / candidates per
// letter than the defaults provide for
mgc_result *run_pipeline_growing(mgc_engine *e, bool *amb_out) {
    while (true) {
        try {
            return run_pipeline(e, amb_out);
        } catch (const mgc::ScratchExhausted &) {
            if (e->cap_scale >= 32) throw;
            e->cap_scale *= 2;
            ensure_batch_buffers(e);
        }
    }
}

mgc_result *do_decode(mgc_engine *e) {
    bool amb = false;
    std::unique_ptr<mgc_result> res(run_pipeline_growing(e, &amb));
    if (amb && !e->segmented) {
        // runs of >= 40 ambiguous letters split a record (gmhmmp2 --incomplete_at_gaps 40): find them on the device,
        // re-slot the batch and decode again
        cudaStream_t st = e->stream;
        const uint32_t cap = 1u << 20;
        e->gaps.ensure((size_t)cap * 12);
        e->gap_count.ensure(16);
        CUDA_TRY(cudaMemsetAsync(e->gap_count.p, 0, 16, st));
        launch_find_gaps(e->nm.as<uint32_t>(), e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), e->n_contigs, 40, e->gaps.as<int32_t>(),
                         e->gap_count.as<unsigned int>(), cap, st);
        unsigned int ngap = 0;
        CUDA_TRY(cudaMemcpyAsync(&ngap, e->gap_count.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (ngap > cap) throw std::runtime_error("too many gaps in one batch");
        if (ngap) {
            std::vector<int32_t> gaps((size_t)ngap * 3);
            CUDA_TRY(cudaMemcpy(gaps.data(), e->gaps.p, (size_t)ngap * 12, cudaMemcpyDeviceToHost));
            std::vector<std::vector<std::pair<int32_t, int32_t>>> per(e->n_parents);
            for (unsigned int k = 0; k < ngap; k++) per[gaps[k * 3]].push_back({gaps[k * 3 + 1], gaps[k * 3 + 2]});
            std::vector<mgc_engine::Seg> segs;
            for (int32_t c = 0; c < e->n_parents; c++) {
                auto &v = per[c];
                std::sort(v.begin(), v.end());
                int32_t pos = 0;
                const int32_t L = e->h_parent_len[c];
                for (auto &gp : v) {
                    if (gp.first > pos) segs.push_back(mgc_engine::Seg{c, e->h_parent_aoff[c] + pos, gp.first - pos, pos, 0});
                    pos = gp.first + gp.second;
                }
                if (L > pos || v.empty()) segs.push_back(mgc_engine::Seg{c, e->h_parent_aoff[c] + pos, L - pos, pos, 1});
            }
            if (segs.empty()) segs.push_back(mgc_engine::Seg{0, e->h_parent_aoff[0], 0, 0, 1});  // nothing but gaps
            e->segs.swap(segs);
            e->segmented = true;
            set_slots(e);
            res.reset(run_pipeline_growing(e, &amb));
        }
    }
    if (e->segmented) remap_to_parents(e, res.get());
    return res.release();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char *mgc_version(void) { return "mgcod_b200 0.1 (GeneMark.hmm-2 v1.17 compatible decoding, sm_100a)"; }

const char *mgc_last_error(void) { return mgc::last_error().c_str(); }

int mgc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mgc_model_load(const char *mod_path, mgc_model **out) {
    if (!mod_path || !out) return mgc::set_error(MGC_E_ARG, "null argument");
    try {
        auto *m = new mgc_model();
        m->m = parse_model_file(mod_path);
        *out = m;
        return MGC_OK;
    } catch (const std::exception &ex) {
        std::string msg = ex.what();
        return mgc::set_error(msg.find("cannot open") != std::string::npos ? MGC_E_IO : MGC_E_PARSE, msg);
    }
}

void mgc_model_free(mgc_model *m) { delete m; }

int mgc_model_info(const mgc_model *m, int32_t info[6]) {
    if (!m || !info) return mgc::set_error(MGC_E_ARG, "null argument");
    info[0] = m->m.gcode; info[1] = m->m.cod_order; info[2] = m->m.non_order;
    info[3] = m->m.min_len; info[4] = m->m.n_bac; info[5] = m->m.n_arc;
    return MGC_OK;
}

int mgc_engine_create(int device, const mgc_model *const *models, int n_models, mgc_engine **out) {
    if (!models || !out || n_models < 1 || n_models > MGC_MAX_MODELS) return mgc::set_error(MGC_E_ARG, "bad model list");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return mgc::set_error(MGC_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(ce) +
                                              "); this engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return mgc::set_error(MGC_E_ARG, "device index out of range");
    mgc_engine *e = nullptr;
    try {
        CUDA_TRY(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw std::runtime_error("CUDA: device is not sm_100 class (built for sm_100a only)");
        e = new mgc_engine();
        e->device = device;
        e->n_sm = prop.multiProcessorCount;
        CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        for (auto &ev : e->ev) CUDA_TRY(cudaEventCreate(&ev));
        CUDA_TRY(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&e->copy_done, cudaEventDisableTiming));
        CUDA_TRY(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&e->main_done, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreate(&e->d2h_begin));
        CUDA_TRY(cudaEventCreate(&e->d2h_end));
        e->n_models = n_models;
        e->pm.resize(n_models);
        for (int i = 0; i < n_models; i++) {
            if (!models[i]) throw std::invalid_argument("null model");
            e->models.push_back(models[i]->m);
        }
        upload_models(e);
        *out = e;
        return MGC_OK;
    } catch (const std::invalid_argument &ex) {
        delete e;
        return mgc::set_error(MGC_E_ARG, ex.what());
    } catch (const std::exception &ex) {
        delete e;
        return mgc::set_error(MGC_E_CUDA, ex.what());
    }
}

void mgc_engine_destroy(mgc_engine *e) {
    if (!e) return;
    if (e->prepared.valid) { e->prepared.valid = false; try { e->prepared.ready.get(); } catch (...) {} }   // it reads members of *e
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (auto &p : e->pm) {
        for (DevBuf *b : {&p.lc_full, &p.ln_full, &p.lc_ramp, &p.ln_ramp, &p.sc, &p.quad, &p.rquad, &p.wtab, &p.wtab2, &p.pair_amb, &p.pair_ub, &p.ls, &p.ls_tile, &p.items, &p.items_sorted, &p.hist,
                          &p.item_off, &p.unit_off, &p.sums, &p.orfs, &p.starts, &p.surv, &p.sorted, &p.bcount,
                          &p.cum, &p.prev, &p.gptr, &p.av_right, &p.av_left, &p.av_ls, &p.av_idx, &p.av_cum, &p.av_pmax, &p.av_parg, &p.st, &p.st_rl,
                          &p.chain, &p.gene_count, &p.gene_off, &p.contig_total, &p.genes, &p.piece_b0, &p.piece_b1, &p.work,
                          &p.piece_total, &p.slot_gene_off, &p.cover, &p.minleft})
            b->release();
    }
    for (DevBuf *b : {&e->d_models, &e->ascii, &e->aoff, &e->cwoff, &e->clen, &e->grp2contig, &e->lo, &e->hi, &e->gc,
                      &e->tile_sums, &e->flags, &e->t_orfs, &e->t_starts, &e->t_surv, &e->t_sorted, &e->t_bcount, &e->t_cover, &e->t_scan_buckets, &e->t_gene_count, &e->t_slot_gene_off, &e->t_gene_off,
                      &e->t_genes, &e->t_chainbufs, &e->counters, &e->surv_counters, &e->work_counters, &e->model_cents,
                      &e->gene_totals, &e->t_ls, &e->t_ls_tile, &e->t_items, &e->t_items_sorted, &e->t_hist, &e->t_item_off,
                      &e->t_unit_off, &e->t_sums, &e->ascii_next, &e->slotbase, &e->slot_contig, &e->nm, &e->ncount, &e->codon, &e->cflags, &e->plen, &e->gaps, &e->gap_count, &e->unit_totals, &e->next_unit, &e->win_stride, &e->win_off, &e->win_labels, &e->win_cents, &e->win_merged, &e->win_run, &e->win_status, &e->win_hist})
        b->release();
    for (auto &ev : e->ev) cudaEventDestroy(ev);
    e->pool->release(e->table_stage);
    e->pool->release(e->win_host);   // back to the pool, which frees its blocks with the last result / engine holding it
    e->win_host = PinnedPool::Block();
    if (e->last_ready) e->last_ready->wait();
    e->last_ready.reset();
    if (e->d2h_stream) { cudaStreamSynchronize(e->d2h_stream); cudaStreamDestroy(e->d2h_stream); }
    for (cudaEvent_t ev : {e->main_done, e->d2h_begin, e->d2h_end}) if (ev) cudaEventDestroy(ev);
    if (e->copy_done) cudaEventDestroy(e->copy_done);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int guarded(const std::function<void()> &fn) {
    try {
        fn();
        return MGC_OK;
    } catch (const mgc::AlphabetError &ex) {
        return mgc::set_error(MGC_E_ALPHABET, ex.what());
    } catch (const std::invalid_argument &ex) {
        return mgc::set_error(MGC_E_ARG, ex.what());
    } catch (const std::bad_alloc &) {
        return mgc::set_error(MGC_E_NOMEM, "host allocation failed");
    } catch (const std::exception &ex) {
        std::string m = ex.what();
        cudaGetLastError();
        return mgc::set_error(m.find("scratch exhausted") != std::string::npos || m.find("out of device memory") != std::string::npos
                                  ? MGC_E_NOMEM
                                  : MGC_E_CUDA,
                              m);
    }
}

int mgc_upload_batch(mgc_engine *e, const char *seq, const int64_t *offsets, int32_t n_contigs) {
    if (!e || !seq || !offsets) return mgc::set_error(MGC_E_ARG, "null argument");
    return guarded([&] { do_upload(e, seq, offsets, n_contigs); });
}

int mgc_prefetch_batch(mgc_engine *e, const char *seq, const int64_t *offsets, int32_t n_contigs) {
    if (!e || !seq || !offsets || n_contigs <= 0) return mgc::set_error(MGC_E_ARG, "bad argument");
    return guarded([&] {
        CUDA_TRY(cudaSetDevice(e->device));
        CUDA_TRY(cudaStreamSynchronize(e->copy_stream));
        const int64_t a0 = offsets[0], a1 = offsets[n_contigs];
        if (a1 < a0) throw std::invalid_argument("offsets must ascend");
        e->ascii_next.ensure((size_t)(a1 - a0) + 2 * PACK_SLACK);
        CUDA_TRY(cudaMemcpyAsync(e->ascii_next.as<char>() + PACK_SLACK, seq + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, e->copy_stream));
        CUDA_TRY(cudaEventRecord(e->copy_done, e->copy_stream));
        // a preparation still running reads the previous offsets: let it finish before they are replaced
        if (e->prepared.valid) { e->prepared.valid = false; try { e->prepared.ready.get(); } catch (...) {} }
        e->pending.valid = true;
        e->pending.seq = seq;
        e->pending.offsets.assign(offsets, offsets + n_contigs + 1);
        // the batch's host tables (slots, word offsets, chain piece slots: ~1 ms for 43 k contigs) are built by a background
        // thread now, not between two decodes with the device idle
        mgc_engine::Prepared *pp = &e->prepared;
        const int64_t *offs = e->pending.offsets.data();      // stays untouched until the batch is adopted or replaced
        const int32_t nn = n_contigs;
        pp->ready = std::async(std::launch::async, [pp, offs, nn] {
            segs_from_offsets(offs, nn, pp->segs, pp->parent_len, pp->parent_aoff);
            compute_slot_tables(pp->segs, pp->tables);
        });
        pp->valid = true;
    });
}

int mgc_decode_resident(mgc_engine *e, mgc_result **out) {
    if (!e || !out) return mgc::set_error(MGC_E_ARG, "null argument");
    return guarded([&] { *out = do_decode(e); });
}

int mgc_decode_batch(mgc_engine *e, const char *seq, const int64_t *offsets, int32_t n_contigs, mgc_result **out) {
    if (!e || !seq || !offsets || !out) return mgc::set_error(MGC_E_ARG, "null argument");
    return guarded([&] {
        do_upload(e, seq, offsets, n_contigs);
        *out = do_decode(e);
    });
}

int mgc_merge_labels(mgc_engine *e, const int32_t *labels, const int64_t *win_offsets, int32_t n_contigs, int32_t consecutive,
                     uint32_t noise_labels, int32_t *merged, int32_t *status) {
    if (!e || !labels || !win_offsets || !merged || !status || n_contigs < 0 || consecutive < 1)
        return mgc::set_error(MGC_E_ARG, "bad argument");
    return guarded([&] {
        CUDA_TRY(cudaSetDevice(e->device));
        const int64_t nw = win_offsets[n_contigs];
        cudaStream_t st = e->stream;
        // scratch lives in the engine: no allocation per call once it has grown to the largest input seen
        e->win_labels.ensure((size_t)std::max<int64_t>(nw, 1) * 4);
        e->win_merged.ensure((size_t)std::max<int64_t>(nw, 1) * 4);
        e->win_run.ensure((size_t)std::max<int64_t>(nw, 1) * 4);
        e->win_off.ensure((size_t)(n_contigs + 1) * 8);
        e->win_status.ensure((size_t)std::max(n_contigs, 1) * 4);
        CUDA_TRY(cudaMemcpyAsync(e->win_labels.p, labels, (size_t)nw * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->win_off.p, win_offsets, (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, st));
        if (n_contigs)
            launch_merge_labels(e->win_labels.as<int32_t>(), e->win_off.as<int64_t>(), n_contigs, consecutive, noise_labels, nullptr,
                                e->win_merged.as<int32_t>(), e->win_run.as<int32_t>(), e->win_status.as<int32_t>(), st);
        if (nw) CUDA_TRY(cudaMemcpyAsync(merged, e->win_merged.p, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        if (n_contigs) CUDA_TRY(cudaMemcpyAsync(status, e->win_status.p, (size_t)n_contigs * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
    });
}

int64_t mgc_debug_candidates(mgc_engine *e, int model, mgc_candidate *out, int64_t cap) {
    if (!e || model < 0 || model >= e->n_models || !e->have_genes) return mgc::set_error(MGC_E_ARG, "bad argument");
    int64_t n = -1;
    int rc = guarded([&] {
        CUDA_TRY(cudaSetDevice(e->device));
        unsigned long long ns = 0;
        CUDA_TRY(cudaMemcpy(&ns, e->counters.as<unsigned long long>() + (size_t)model * 8 + 4, 8, cudaMemcpyDeviceToHost));
        ns &= 0xffffffffull;   // candidates in `sorted` (the bucket scan's total)
        std::vector<Cand> c(ns);
        std::vector<double> cum(ns);
        if (ns) {
            CUDA_TRY(cudaMemcpy(c.data(), e->pm[model].surv.p, ns * sizeof(Cand), cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(cum.data(), e->pm[model].cum.p, ns * 8, cudaMemcpyDeviceToHost));
        }
        n = (int64_t)ns;
        // chain scores are stored relative to the candidate's piece; add the best score of everything before the piece
        const int32_t nsl = e->n_slots;
        std::vector<int32_t> b0(nsl), b1(nsl);
        std::vector<double> pt(nsl), before(nsl, 0.0);
        if (nsl) {
            CUDA_TRY(cudaMemcpy(b0.data(), e->pm[model].piece_b0.p, (size_t)nsl * 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(b1.data(), e->pm[model].piece_b1.p, (size_t)nsl * 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(pt.data(), e->pm[model].piece_total.p, (size_t)nsl * 8, cudaMemcpyDeviceToHost));
        }
        std::vector<int32_t> sbase(e->n_contigs + 1, 0);
        for (int32_t ci = 0; ci < e->n_contigs; ci++) {
            const int64_t nbk = ((e->h_clen[ci] > 0 ? e->h_clen[ci] - 1 : 0) >> 7) + 1;
            sbase[ci + 1] = sbase[ci] + (int32_t)((nbk + PIECE_BUCKETS - 1) / PIECE_BUCKETS);
            double acc = 0.0;
            for (int32_t s = sbase[ci]; s < sbase[ci + 1]; s++) { before[s] = acc; if (b1[s] >= 0) acc += pt[s]; }
        }
        for (int64_t i = 0; i < n && i < cap && out; i++) {
            mgc_candidate &o = out[i];
            o.contig = c[i].contig; o.left = c[i].left; o.right = c[i].right; o.strand = c[i].strand; o.gc = c[i].gc;
            o.domain = c[i].flags & 1; o.partial_start = (c[i].flags >> 1) & 1; o.partial_stop = (c[i].flags >> 2) & 1;
            o.total = c[i].total; o.cum = cum[i];
            const int64_t b = e->h_cwoff[o.contig] / 4 + (o.right >> 7);
            for (int32_t s = sbase[o.contig]; s < sbase[o.contig + 1]; s++)
                if (b1[s] >= 0 && b >= b0[s] && b <= b1[s]) { o.cum += before[s]; break; }
        }
    });
    return rc == MGC_OK ? n : rc;
}

int mgc_engine_device_genes(mgc_engine *e, int model, const void **dev_genes, int64_t *count) {
    if (!e || !dev_genes || !count || model < 0 || model >= e->n_models) return mgc::set_error(MGC_E_ARG, "bad argument");
    if (!e->have_genes || (size_t)model >= e->dev_gene_count.size()) return mgc::set_error(MGC_E_ARG, "decode a batch first");
    return guarded([&] {
        CUDA_TRY(cudaSetDevice(e->device));
        CUDA_TRY(cudaStreamSynchronize(e->stream));   // the gene tables are complete (the caller reads them on another stream)
        *dev_genes = e->pm[model].genes.p;
        *count = e->dev_gene_count[model];
    });
}

int mgc_engine_stats(const mgc_engine *ce, mgc_stats *out) {
    if (!ce || !out) return mgc::set_error(MGC_E_ARG, "null argument");
    mgc_engine *e = const_cast<mgc_engine *>(ce);
    if (!e->d2h_timed && e->last_ready) {   // the copy-out of the last result is timed once it has finished
        e->last_ready->wait();
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e->d2h_begin, e->d2h_end) == cudaSuccess) {
            e->stats.d2h = ms;
            e->stats.total += ms;
        }
        e->d2h_timed = true;
    }
    *out = e->stats;
    return MGC_OK;
}

int64_t mgc_result_num_genes(const mgc_result *r, int model) {
    if (!r || model < 0 || model >= r->n_models) return -1;
    return r->n_genes[model];
}

const mgc_gene *mgc_result_genes(const mgc_result *r, int model) {
    if (!r || model < 0 || model >= r->n_models) return nullptr;
    r->wait();
    return static_cast<const mgc_gene *>(r->gene_blk[model].p);
}

const double *mgc_result_contig_totals(const mgc_result *r, int model) {
    if (!r || model < 0 || model >= r->n_models) return nullptr;
    r->wait();
    return static_cast<const double *>(r->total_blk[model].p);
}

int mgc_result_model_cents(const mgc_result *r, int64_t *cents) {
    if (!r || !cents) return mgc::set_error(MGC_E_ARG, "null argument");
    for (int i = 0; i < r->n_models; i++) cents[i] = r->cents[i];
    return MGC_OK;
}

int32_t mgc_result_min_gene_length(const mgc_result *r, int model) {
    if (!r || model < 0 || model >= r->n_models) return mgc::set_error(MGC_E_ARG, "model index out of range");
    return r->min_len[model];
}

int mgc_result_set_min_gene_length(mgc_result *r, int model, int32_t min_len) {
    if (!r || model < 0 || model >= r->n_models || min_len < 0) return mgc::set_error(MGC_E_ARG, "bad argument");
    r->min_len[model] = min_len;
    return MGC_OK;
}

void mgc_result_free(mgc_result *r) { delete r; }

int mgc_result_create(int n_models, int32_t n_contigs, const mgc_gene *const *genes, const int64_t *counts,
                      const double *const *totals, mgc_result **out) {
    if (n_models < 1 || n_models > MGC_MAX_MODELS || n_contigs < 0 || !genes || !counts || !totals || !out)
        return mgc::set_error(MGC_E_ARG, "bad argument");
    auto *r = new mgc_result();
    r->n_models = n_models;
    r->n_contigs = n_contigs;
    for (int m = 0; m < n_models; m++) {
        PinnedPool::Block g, t;   // plain heap blocks; no pool attached, freed below through host_owned
        g.bytes = (size_t)std::max<int64_t>(counts[m], 1) * sizeof(mgc_gene);
        g.p = std::malloc(g.bytes);
        t.bytes = (size_t)std::max<int32_t>(n_contigs, 1) * 8;
        t.p = std::malloc(t.bytes);
        if (!g.p || !t.p) {
            std::free(g.p);
            std::free(t.p);
            r->host_owned = true;
            mgc_result_free(r);
            return mgc::set_error(MGC_E_NOMEM, "out of host memory for the gene tables");
        }
        if (counts[m]) std::memcpy(g.p, genes[m], (size_t)counts[m] * sizeof(mgc_gene));
        if (n_contigs) std::memcpy(t.p, totals[m], (size_t)n_contigs * 8);
        r->gene_blk.push_back(g);
        r->total_blk.push_back(t);
        r->n_genes.push_back(counts[m]);
        r->min_len.push_back(90);
        long long cents = 0;
        for (int64_t k = 0; k < counts[m]; k++) cents += llrint(genes[m][k].logodd * 100.0);
        r->cents.push_back(cents);
    }
    r->host_owned = true;
    *out = r;
    return MGC_OK;
}

int mgc_window_counts(const int64_t *contig_len, const int32_t *stride, int32_t n_contigs, int32_t window, int32_t consecutive,
                      int short_contigs, int64_t *win_offsets) {
    if (!contig_len || !stride || !win_offsets || window <= 0) return mgc::set_error(MGC_E_ARG, "bad argument");
    int64_t acc = 0;
    for (int32_t c = 0; c < n_contigs; c++) {
        win_offsets[c] = acc;
        const int64_t L = contig_len[c], st = stride[c];
        if (st <= 0) return mgc::set_error(MGC_E_ARG, "stride must be positive");
        int64_t nw;
        if (L <= 0) nw = 0;
        else if (L <= 5000 && !short_contigs) nw = 1;                         // bin/pipeline.py:196-225
        else if (L <= (int64_t)window + (int64_t)(consecutive - 1) * st)      // bin/pipeline.py:227-275
            nw = L > window ? (L - window + st - 1) / st + 1 : 1;
        else nw = (L + st - 1) / st;                                          // bin/pipeline.py:194,329
        acc += nw;
    }
    win_offsets[n_contigs] = acc;
    return MGC_OK;
}

int64_t mgc_recoded_stops(mgc_engine *e, const mgc_gene *genes, int64_t n_genes, int32_t *counts, int64_t *pos_off, int32_t *positions,
                          int64_t pos_cap) {
    if (!e || (!genes && n_genes > 0) || n_genes < 0 || !counts || !pos_off) return mgc::set_error(MGC_E_ARG, "bad argument");
    if (!e->have_genes) return mgc::set_error(MGC_E_ARG, "decode a batch before asking for recoded stops");
    int64_t total = 0;
    int rc = guarded([&] {
        CUDA_TRY(cudaSetDevice(e->device));
        cudaStream_t st = e->stream;
        // genes are in the caller's record coordinates; locate each in its slot (a record split at gaps has several)
        std::vector<int32_t> first_seg(e->n_parents + 1, 0);
        for (size_t s = 0; s < e->segs.size(); s++) first_seg[e->segs[s].parent + 1] = (int32_t)s + 1;
        for (int32_t c = 0; c < e->n_parents; c++) first_seg[c + 1] = std::max(first_seg[c + 1], first_seg[c]);
        std::vector<GeneRef> refs((size_t)n_genes);
        for (int64_t i = 0; i < n_genes; i++) {
            const mgc_gene &g = genes[i];
            if (g.contig < 0 || g.contig >= e->n_parents || g.left < 1 || g.right < g.left || g.right > e->h_parent_len[g.contig])
                throw std::invalid_argument("gene " + std::to_string(i) + " lies outside its contig");
            int32_t lo = first_seg[g.contig], hi = first_seg[g.contig + 1] - 1, slot = -1;
            while (lo <= hi) {  // last segment starting at or before the gene
                const int32_t mid = (lo + hi) / 2;
                if (e->segs[mid].poff <= g.left - 1) { slot = mid; lo = mid + 1; } else hi = mid - 1;
            }
            if (slot < 0 || g.right - e->segs[slot].poff > e->segs[slot].len)
                throw std::invalid_argument("gene " + std::to_string(i) + " crosses a gap");
            const int32_t l0 = g.left - 1 - e->segs[slot].poff, r0 = g.right - 1 - e->segs[slot].poff, L = e->segs[slot].len;
            refs[i] = GeneRef{slot, g.strand < 0 ? 1 : 0, g.strand < 0 ? L - 1 - r0 : l0, r0 - l0 + 1};
        }
        DevBuf d_refs, d_counts, d_off, d_pos;
        d_refs.ensure(std::max<size_t>((size_t)n_genes * sizeof(GeneRef), 16));
        d_counts.ensure(std::max<size_t>((size_t)n_genes * 12, 16));
        Planes pl;
        pl.lo = e->lo.as<uint32_t>(); pl.hi = e->hi.as<uint32_t>(); pl.nm = e->nm.as<uint32_t>(); pl.npre = e->ncount.as<uint32_t>();
        pl.dflags = e->flags.as<int>(); pl.cflags = e->cflags.as<uint8_t>(); pl.codon = e->codon.as<uint32_t>(); pl.W = e->W;
        { const WalkTiles wt = walk_tiles_for(e->W, walk_tile_words_setting()); pl.n_tiles = wt.n; pl.tile_words = wt.words; }
        if (n_genes) CUDA_TRY(cudaMemcpyAsync(d_refs.p, refs.data(), (size_t)n_genes * sizeof(GeneRef), cudaMemcpyHostToDevice, st));
        launch_recoded_stops(pl, e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), d_refs.as<GeneRef>(), n_genes, d_counts.as<int32_t>(), nullptr,
                             nullptr, false, e->n_sm, st);
        if (n_genes) CUDA_TRY(cudaMemcpyAsync(counts, d_counts.p, (size_t)n_genes * 12, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        pos_off[0] = 0;
        for (int64_t i = 0; i < n_genes * 3; i++) pos_off[i + 1] = pos_off[i] + counts[i];
        total = pos_off[n_genes * 3];
        if (positions && total > 0 && pos_cap >= total) {
            d_off.ensure((size_t)(n_genes * 3 + 1) * 8);
            d_pos.ensure((size_t)total * 4);
            CUDA_TRY(cudaMemcpyAsync(d_off.p, pos_off, (size_t)(n_genes * 3 + 1) * 8, cudaMemcpyHostToDevice, st));
            launch_recoded_stops(pl, e->cwoff.as<int64_t>(), e->clen.as<int32_t>(), d_refs.as<GeneRef>(), n_genes, nullptr, d_off.as<int64_t>(),
                                 d_pos.as<int32_t>(), true, e->n_sm, st);
            CUDA_TRY(cudaMemcpyAsync(positions, d_pos.p, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        CUDA_TRY(cudaGetLastError());
        d_refs.release(); d_counts.release(); d_off.release(); d_pos.release();
    });
    return rc == MGC_OK ? total : rc;
}

int mgc_window_labels(mgc_engine *e, const int32_t *stride, int32_t n_contigs, int32_t window, int32_t consecutive, int short_contigs,
                      const int64_t *win_offsets, int32_t *labels, int64_t *cents) {
    if (!e || !stride || !win_offsets || !labels || !cents || window <= 0 || consecutive < 1) return mgc::set_error(MGC_E_ARG, "bad argument");
    if (!e->have_genes) return mgc::set_error(MGC_E_ARG, "decode a batch before asking for window labels");
    if (n_contigs != e->n_parents)
        return mgc::set_error(MGC_E_ARG, "window labels: the caller's contig count differs from the batch resident on the device");
    return guarded([&] {
        CUDA_TRY(cudaSetDevice(e->device));
        const int32_t n = e->n_parents;
        const int M = e->n_models;
        const int64_t nw = win_offsets[n];
        cudaStream_t st = e->stream;
        e->win_stride.ensure((size_t)n * 4);
        e->win_off.ensure((size_t)(n + 1) * 8);
        e->win_labels.ensure((size_t)std::max<int64_t>(nw, 1) * 4);
        e->win_cents.ensure((size_t)std::max<int64_t>(nw, 1) * M * 8);
        CUDA_TRY(cudaMemcpyAsync(e->win_stride.p, stride, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->win_off.p, win_offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        launch_windows(e->t_genes.as<const mgc_gene *>(), e->t_gene_off.as<const uint32_t *>(), M, e->plen.as<int32_t>(),
                       e->win_stride.as<int32_t>(), e->win_off.as<int64_t>(), n, window, consecutive, short_contigs,
                       e->win_labels.as<int32_t>(), e->win_cents.as<long long>(), nullptr, st);
        if (nw) {
            CUDA_TRY(cudaMemcpyAsync(labels, e->win_labels.p, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(cents, e->win_cents.p, (size_t)nw * M * 8, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
    });
}

int mgc_isoform_labels(mgc_engine *e, const int32_t *stride, int32_t n_contigs, int32_t window, int32_t consecutive, int short_contigs,
                       const int64_t *win_offsets, int32_t *labels, int64_t *cents, int32_t *merged, int32_t *status) {
    if (!e || !stride || !win_offsets || !labels || !merged || !status || window <= 0 || consecutive < 1)
        return mgc::set_error(MGC_E_ARG, "bad argument");
    if (!e->have_genes) return mgc::set_error(MGC_E_ARG, "decode a batch before asking for window labels");
    if (n_contigs != e->n_parents)
        return mgc::set_error(MGC_E_ARG, "window labels: the caller's contig count differs from the batch resident on the device");
    return guarded([&] {
        CUDA_TRY(cudaSetDevice(e->device));
        const int32_t n = e->n_parents;
        const int M = e->n_models;
        const int64_t nw = win_offsets[n], nw1 = std::max<int64_t>(nw, 1);
        cudaStream_t st = e->stream;
        e->win_stride.ensure((size_t)n * 4);
        e->win_off.ensure((size_t)(n + 1) * 8);
        e->win_labels.ensure((size_t)nw1 * 4);
        e->win_cents.ensure((size_t)nw1 * M * 8);
        e->win_merged.ensure((size_t)nw1 * 4);
        e->win_run.ensure((size_t)nw1 * 4);
        e->win_status.ensure((size_t)std::max(n, 1) * 4);
        e->win_hist.ensure(MGC_MAX_MODELS * 8 + 16);
        // one page-locked block receives labels | merged | status | cents: a single pass over PCIe, then plain copies
        const size_t o_lab = 0, o_mer = (size_t)nw1 * 4, o_sta = o_mer + (size_t)nw1 * 4;
        const size_t o_cen = (o_sta + (size_t)std::max(n, 1) * 4 + 15) & ~(size_t)15;
        const size_t need = o_cen + (cents ? (size_t)nw1 * M * 8 : 0);
        if (e->win_host.bytes < need) {
            e->pool->release(e->table_stage);
    e->pool->release(e->win_host);
            e->win_host = PinnedPool::Block();
            e->win_host = e->pool->acquire(need);
        }
        char *hp = static_cast<char *>(e->win_host.p);
        CUDA_TRY(cudaMemsetAsync(e->win_hist.p, 0, MGC_MAX_MODELS * 8 + 16, st));
        CUDA_TRY(cudaMemcpyAsync(e->win_stride.p, stride, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->win_off.p, win_offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        unsigned long long *hist = e->win_hist.as<unsigned long long>();
        uint32_t *mask = reinterpret_cast<uint32_t *>(hist + MGC_MAX_MODELS);
        // window sums + labels + label histogram -> noise labels -> smoothing; the labels never leave the device in between
        launch_windows(e->t_genes.as<const mgc_gene *>(), e->t_gene_off.as<const uint32_t *>(), M, e->plen.as<int32_t>(),
                       e->win_stride.as<int32_t>(), e->win_off.as<int64_t>(), n, window, consecutive, short_contigs,
                       e->win_labels.as<int32_t>(), e->win_cents.as<long long>(), hist, st);
        launch_noise_mask(hist, M, mask, st);
        if (n)
            launch_merge_labels(e->win_labels.as<int32_t>(), e->win_off.as<int64_t>(), n, consecutive, 0u, mask, e->win_merged.as<int32_t>(),
                                e->win_run.as<int32_t>(), e->win_status.as<int32_t>(), st);
        if (nw) {
            CUDA_TRY(cudaMemcpyAsync(hp + o_lab, e->win_labels.p, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(hp + o_mer, e->win_merged.p, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
            if (cents) CUDA_TRY(cudaMemcpyAsync(hp + o_cen, e->win_cents.p, (size_t)nw * M * 8, cudaMemcpyDeviceToHost, st));
        }
        if (n) CUDA_TRY(cudaMemcpyAsync(hp + o_sta, e->win_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
        if (nw) {
            std::memcpy(labels, hp + o_lab, (size_t)nw * 4);
            std::memcpy(merged, hp + o_mer, (size_t)nw * 4);
            if (cents) std::memcpy(cents, hp + o_cen, (size_t)nw * M * 8);
        }
        if (n) std::memcpy(status, hp + o_sta, (size_t)n * 4);
    });
}

}  // extern "C"
