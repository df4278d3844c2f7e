// synth_workload.cpp — the synthetic metagenome of BASELINE.json config 5: contig lengths ~ L^-2 on
// [min_len, max_len], gene-like content with a per-contig GC regime and true genetic code, stop-codon reassignment after
// the reference's utils/simulate_gcodes.py (:96-101, :112-149).
//
// A workload GENERATOR, host only, deliberately outside the product library (mgcod_b200/libmgcod_b200.so): bench.py's
// reference arm, the golden-file scripts and the tests produce their inputs with it without mapping any product code.
// Built by workload/Makefile into workload/libmgc_workload.so; Python face: workload/__init__.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace {

// ---- xoshiro256** ----
struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) {
        for (auto &v : s) v = splitmix(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
};

const char NT[4] = {'A', 'C', 'G', 'T'};

// codon distribution at GC level gc for genetic code 11 (same construction as mgcod_b200/synth.py codon_usage)
void codon_usage(double gc, double C[64]) {
    const double at = (1.0 - gc) / 2.0, g_c = gc / 2.0;
    const double base[4] = {at, g_c, g_c, at};
    const double f1[4] = {1.15, 0.80, 1.45, 0.70}, f2[4] = {1.20, 0.95, 0.75, 1.10}, f3[4] = {0.95, 1.05, 0.95, 1.05};
    double p1[4], p2[4], p3[4], s1 = 0, s2 = 0, s3 = 0;
    for (int x = 0; x < 4; x++) {
        p1[x] = base[x] * f1[x];
        p2[x] = base[x] * f2[x];
        double w = (x == 1 || x == 2) ? gc : 1.0 - gc;
        p3[x] = std::pow(w, 1.0 + 2.0 * 0.35) * f3[x];
        s1 += p1[x]; s2 += p2[x]; s3 += p3[x];
    }
    double tot = 0;
    for (int i = 0; i < 64; i++) {
        int a = i >> 4, b = (i >> 2) & 3, c = i & 3;
        double v = (p1[a] / s1) * (p2[b] / s2) * (p3[c] / s3);
        v *= 1.0 + 0.45 * std::sin(1.7 * i + 0.3) * std::cos(0.9 * i);
        if (i == 48 || i == 50 || i == 56) v = 0.0;  // TAA TAG TGA never inside a gene before recoding
        C[i] = v;
        tot += v;
    }
    for (int i = 0; i < 64; i++) C[i] /= tot;
}

int draw(const double *cdf, int n, double u) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

void revcomp_inplace(char *s, int64_t n) {
    for (int64_t i = 0, j = n - 1; i <= j; i++, j--) {
        auto comp = [](char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; };
        char a = comp(s[i]), b = comp(s[j]);
        s[i] = b;
        s[j] = a;
    }
}

void synth_one(uint64_t seed, int64_t index, char *out, int64_t L, int32_t *truth) {
    uint64_t sd = seed * 0x9e3779b97f4a7c15ULL + (uint64_t)index * 0xd1342543de82ef95ULL + 1;
    Rng rng(sd);
    static const int gcs[4] = {30, 40, 50, 60};
    static const int codes[4] = {11, 4, 15, 101};
    const double gc = gcs[rng.next() & 3] / 100.0;
    const int code = codes[rng.next() & 3];
    if (truth) *truth = code;
    double C[64], cdf[64], acc = 0;
    codon_usage(gc, C);
    for (int i = 0; i < 64; i++) { acc += C[i]; cdf[i] = acc; }
    cdf[63] = 1.0;
    const double ntp[4] = {(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2};
    const double ntcdf[4] = {ntp[0], ntp[0] + ntp[1], ntp[0] + ntp[1] + ntp[2], 1.0};
    // stop codons valid under the true code and the (sense codon -> reassigned stop) substitution, p = 0.8
    // (utils/simulate_gcodes.py:96-101, 112-149 in the reference)
    int stops[3], nstop = 0, sense_from = -1, sense_to = -1;
    switch (code) {
        case 4: stops[0] = 48; stops[1] = 50; nstop = 2; sense_from = 58 /*TGG*/; sense_to = 56 /*TGA*/; break;
        case 15: stops[0] = 48; stops[1] = 56; nstop = 2; sense_from = 18 /*CAG*/; sense_to = 50 /*TAG*/; break;
        case 101: stops[0] = 50; stops[1] = 56; nstop = 2; sense_from = 16 /*CAA*/; sense_to = 48 /*TAA*/; break;
        default: stops[0] = 48; stops[1] = 50; stops[2] = 56; nstop = 3;
    }
    int64_t pos = 0;
    std::vector<char> gene;
    while (pos < L) {
        // intergenic stretch, geometric with mean 150
        int64_t ig = (int64_t)(-150.0 * std::log(1.0 - rng.uni())) + 1;
        for (int64_t k = 0; k < ig && pos < L; k++) out[pos++] = NT[draw(ntcdf, 4, rng.uni())];
        if (pos >= L) break;
        // gene: 3 * (30 + Gamma(2, 135)) codons' worth of nt, i.e. mean ~900 nt
        double gam = -135.0 * (std::log(1.0 - rng.uni()) + std::log(1.0 - rng.uni()));
        int64_t ncod = 30 + (int64_t)gam;
        ncod = std::max<int64_t>(ncod, 32);
        gene.clear();
        double us = rng.uni();
        const char *sc = us < 0.85 ? "ATG" : (us < 0.95 ? "GTG" : "TTG");
        gene.insert(gene.end(), sc, sc + 3);
        for (int64_t k = 0; k < ncod; k++) {
            int cd = draw(cdf, 64, rng.uni());
            if (cd == sense_from && rng.uni() < 0.8) cd = sense_to;
            gene.push_back(NT[cd >> 4]); gene.push_back(NT[(cd >> 2) & 3]); gene.push_back(NT[cd & 3]);
        }
        int st;
        if (nstop == 3) { double u = rng.uni(); st = u < 0.5 ? 48 : (u < 0.7 ? 50 : 56); }
        else st = rng.uni() < 0.35 ? stops[0] : stops[1];
        gene.push_back(NT[st >> 4]); gene.push_back(NT[(st >> 2) & 3]); gene.push_back(NT[st & 3]);
        if (rng.next() & 1) revcomp_inplace(gene.data(), (int64_t)gene.size());
        for (size_t k = 0; k < gene.size() && pos < L; k++) out[pos++] = gene[k];
    }
}

}  // namespace

extern "C" {

// Draws contig lengths ~ L^-2 on [min_len, max_len] until total >= total_bp; returns the count (-1: bad argument,
// -2: `lens` too small).
int64_t wl_synth_lengths(uint64_t seed, int64_t total_bp, int32_t min_len, int32_t max_len, int64_t *lens, int64_t cap) {
    if (!lens || min_len <= 0 || max_len < min_len) return -1;
    Rng rng(seed);
    int64_t n = 0, acc = 0;
    const double a = (double)min_len, b = (double)max_len;
    while (acc < total_bp) {
        if (n >= cap) return -2;
        // inverse CDF of p(L) ~ L^-2 on [a, b]
        double u = rng.uni();
        double L = 1.0 / (1.0 / a - u * (1.0 / a - 1.0 / b));
        int64_t l = (int64_t)L;
        if (l < min_len) l = min_len;
        if (l > max_len) l = max_len;
        lens[n++] = l;
        acc += l;
    }
    return n;
}

int wl_synth_contigs(uint64_t seed, const int64_t *lens, int64_t n, char *seq, int32_t *truth_code, int threads) {
    if (!lens || !seq || n < 0) return -1;
    std::vector<int64_t> off(n + 1, 0);
    for (int64_t i = 0; i < n; i++) off[i + 1] = off[i] + lens[i];
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            for (int64_t i = t; i < n; i += threads) synth_one(seed, i, seq + off[i], lens[i], truth_code ? truth_code + i : nullptr);
        });
    for (auto &th : pool) th.join();
    return 0;
}

}  // extern "C"
