// mgc_host.cpp — host-only parts of the C-ABI: error text, GFF/LST writers, synthetic workload.
//
// Writers reproduce gmhmmp2's Output::PrintGenes_GFF / PrintGenes_LST byte for byte except for the
// date lines (format strings: SURVEY.md §8b, pinned by tests/golden).  The reference reads the GFF
// back with pandas (bin/data.py:66-96) and caches it on disk (bin/data.py:39).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mgcod_b200.h"
#include "mgc_util.hpp"

namespace mgc {

std::string &last_error() {
    static thread_local std::string msg;
    return msg;
}

int set_error(int code, const std::string &msg) {
    last_error() = msg;
    return code;
}

}  // namespace mgc

namespace {

std::string now_string() {
    std::time_t t = std::time(nullptr);
    char buf[64];
    std::strftime(buf, sizeof buf, "%a %b %e %H:%M:%S %Y", std::localtime(&t));
    // ctime pads the day with a space ("Oct  9"); gmhmmp2 prints the same form
    return buf;
}

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

namespace {

// Append-only text buffer flushed to a FILE in large blocks: the gene lines are formatted by hand (decimal integers and
// "%.2f" with an exact fallback), which is ~6x faster than one fprintf per field for the ~100 bytes per gene.
struct TextOut {
    FILE *f;
    std::vector<char> buf;
    size_t n = 0;
    explicit TextOut(FILE *fp) : f(fp), buf(1 << 22) {}
    void room(size_t need) {
        if (n + need > buf.size()) flush();
    }
    void flush() {
        if (n) std::fwrite(buf.data(), 1, n, f);
        n = 0;
    }
    void str(const char *s, size_t len) {
        room(len);
        if (len > buf.size()) { std::fwrite(s, 1, len, f); return; }
        std::memcpy(buf.data() + n, s, len);
        n += len;
    }
    void str(const char *s) { str(s, std::strlen(s)); }
    void ch(char c) { room(1); buf[n++] = c; }
    void u64(unsigned long long v) {
        char t[24];
        int k = 0;
        do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
        room(k);
        while (k) buf[n++] = t[--k];
    }
    void i64(long long v) {
        if (v < 0) { ch('-'); u64((unsigned long long)(-v)); } else u64((unsigned long long)v);
    }
    // printf("%.2f"): exact unless v*100 lies within 1e-6 of a rounding boundary, where snprintf decides
    void fixed2(double v) {
        const double a = std::fabs(v) * 100.0;
        const double fl = std::floor(a), fr = a - fl;
        if (!(a < 9e15) || std::fabs(fr - 0.5) < 1e-6) {
            char t[64];
            const int k = std::snprintf(t, sizeof t, "%.2f", v);
            str(t, (size_t)k);
            return;
        }
        const unsigned long long r = (unsigned long long)fl + (fr > 0.5 ? 1ull : 0ull);
        if (std::signbit(v)) ch('-');   // printf prints "-0.00" for small negative values
        u64(r / 100);
        ch('.');
        ch((char)('0' + (r / 10) % 10));
        ch((char)('0' + r % 10));
    }
};

}  // namespace

int mgc_write_gff(const mgc_result *r, int model, const char *const *contig_ids, const int64_t *contig_len, int32_t n_contigs,
                  const char *seq_file_name, const char *mod_file_name, int gcode, const char *out_path, const char *format) {
    if (!r || !contig_ids || !contig_len || !out_path || !format) return mgc::set_error(MGC_E_ARG, "null argument");
    const int64_t ng = mgc_result_num_genes(r, model);
    if (ng < 0) return mgc::set_error(MGC_E_ARG, "model index out of range");
    const bool lst = std::strcmp(format, "lst") == 0;
    if (!lst && std::strcmp(format, "gff") != 0) return mgc::set_error(MGC_E_ARG, "format must be gff or lst");
    const mgc_gene *genes = mgc_result_genes(r, model);
    const double *totals = mgc_result_contig_totals(r, model);
    const long long min_len = mgc_result_min_gene_length(r, model);
    FILE *f = std::fopen(out_path, "w");
    if (!f) return mgc::set_error(MGC_E_IO, std::string("cannot open output file: ") + out_path);
    const std::string date = now_string();
    if (lst) std::fprintf(f, "# GeneMark.hmm-2 LST format\n");
    else std::fprintf(f, "##gff-version 2\n");
    std::fprintf(f, "# GeneMark.hmm-2 prokaryotic version: 1.17\n");
    std::fprintf(f, "# File with sequence: %s\n", seq_file_name ? seq_file_name : "");
    std::fprintf(f, "# File with MetaGeneMark parameters: %s\n", mod_file_name ? mod_file_name : "");
    std::fprintf(f, "# translation table: %d\n", gcode);
    std::fprintf(f, "# output date start: %s\n", date.c_str());
    int64_t k = 0;
    long gene_id = 0;
    TextOut o(f);
    char tmp[256];
    for (int32_t c = 0; c < n_contigs; c++) {
        const char *id = contig_ids[c];
        const size_t idlen = std::strlen(id);
        const long long L = (long long)contig_len[c];
        o.ch('\n');
        if (lst) {
            o.str("# sequence-region "); o.ch(L > 0 ? '1' : '0'); o.ch(' '); o.i64(L); o.str("\nSequenceID: "); o.str(id, idlen); o.ch('\n');
        } else {
            o.str("##sequence-region "); o.str(id, idlen); o.ch(' '); o.ch(L > 0 ? '1' : '0'); o.ch(' '); o.i64(L); o.ch('\n');
        }
        int64_t first = k;
        long long len_sum = 0;
        while (k < ng && genes[k].contig == c) {
            const mgc_gene &g = genes[k];
            const int len = g.right - g.left + 1;
            len_sum += len;
            gene_id++;
            const char *type = g.domain ? "archaea" : "bacteria";
            if (lst) {
                const int n = std::snprintf(tmp, sizeof tmp, "%6ld   %c   %c%d   %c%d %7d %s\n", gene_id, g.strand > 0 ? '+' : '-',
                                            g.partial_left ? '<' : ' ', g.left, g.partial_right ? '>' : ' ', g.right, len, type);
                o.str(tmp, (size_t)n);
            } else {
                o.str(id, idlen); o.str("\tGeneMark.hmm2\tCDS\t"); o.i64(g.left); o.ch('\t'); o.i64(g.right); o.ch('\t'); o.fixed2(g.logodd);
                o.ch('\t'); o.ch(g.strand > 0 ? '+' : '-'); o.str("\t0\tgene_id "); o.i64(gene_id); o.str("; gene_type "); o.str(type); o.str("; ");
                if (g.partial_left || g.partial_right) { o.str("partial "); o.ch(g.partial_left ? '1' : '0'); o.ch(g.partial_right ? '1' : '0'); o.str("; "); }
                o.str("gc "); o.i64(g.gc); o.str("; length "); o.i64(len); o.str(";\n");
            }
            k++;
        }
        const int64_t n = k - first;
        if (n > 0) {
            const int m = std::snprintf(tmp, sizeof tmp, "\ttotal_logodd\t%g\taverage_length\t%lld\taverage_density\t%.2f\n", totals[c],
                                        (long long)((double)len_sum / (double)n + 0.5), (double)n / ((double)L / 1000.0));
            o.str("# "); o.str(id, idlen); o.str(tmp, (size_t)m);
        } else if (L >= min_len) {  // a record the binary decoded without finding a gene; shorter records are skipped outright
            o.str("# "); o.str(id, idlen); o.str("\ttotal_logodd\t0\taverage_length\t0\taverage_density\t0.00\n");
        }
    }
    o.flush();
    if (lst) std::fprintf(f, "\n# output date end: %s\n# Done with file: %s\n", date.c_str(), seq_file_name ? seq_file_name : "");
    if (std::fclose(f) != 0) return mgc::set_error(MGC_E_IO, "write failed");
    return MGC_OK;
}

int64_t mgc_synth_lengths(uint64_t seed, int64_t total_bp, int32_t min_len, int32_t max_len, int64_t *lens, int64_t cap) {
    if (!lens || min_len <= 0 || max_len < min_len) return mgc::set_error(MGC_E_ARG, "bad argument");
    Rng rng(seed);
    int64_t n = 0, acc = 0;
    const double a = (double)min_len, b = (double)max_len;
    while (acc < total_bp) {
        if (n >= cap) return mgc::set_error(MGC_E_NOMEM, "length buffer too small");
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

int mgc_synth_contigs(uint64_t seed, const int64_t *lens, int64_t n, char *seq, int32_t *truth_code, int threads) {
    if (!lens || !seq || n < 0) return mgc::set_error(MGC_E_ARG, "bad argument");
    std::vector<int64_t> off(n + 1, 0);
    for (int64_t i = 0; i < n; i++) off[i + 1] = off[i] + lens[i];
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            for (int64_t i = t; i < n; i += threads) synth_one(seed, i, seq + off[i], lens[i], truth_code ? truth_code + i : nullptr);
        });
    for (auto &th : pool) th.join();
    return MGC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// FASTA ingest (replaces Biopython's SeqIO.parse at bin/data.py:108-124 for large inputs): the file is mapped, header
// lines are located with memchr, sequence lines are copied without their line ends into ONE buffer, which can be
// page-locked so that mgc_decode_batch / mgc_prefetch_batch copy from it asynchronously.
// ------------------------------------------------------------------------------------------------
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

struct mgc_fasta {
    std::vector<std::string> ids, deflines;
    std::vector<int64_t> offsets;
    char *seq = nullptr;
    bool pinned = false;
    ~mgc_fasta() {
        if (seq) {
            if (pinned) cudaFreeHost(seq);
            else std::free(seq);
        }
    }
};

int mgc_fasta_read(const char *path, int page_locked, mgc_fasta **out) {
    if (!path || !out) return mgc::set_error(MGC_E_ARG, "null argument");
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return mgc::set_error(MGC_E_IO, std::string("error, cannot open file: ") + path);
    struct stat sb;
    if (::fstat(fd, &sb) != 0) { ::close(fd); return mgc::set_error(MGC_E_IO, std::string("cannot stat ") + path); }
    const size_t size = (size_t)sb.st_size;
    const char *map = nullptr;
    if (size) {
        map = static_cast<const char *>(::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0));
        if (map == MAP_FAILED) { ::close(fd); return mgc::set_error(MGC_E_IO, std::string("cannot map ") + path); }
        ::madvise(const_cast<char *>(map), size, MADV_SEQUENTIAL);
    }
    std::unique_ptr<mgc_fasta> f(new mgc_fasta());
    if (page_locked && cudaHostAlloc(reinterpret_cast<void **>(&f->seq), size + 64, cudaHostAllocDefault) == cudaSuccess) f->pinned = true;
    else {
        cudaGetLastError();  // no device / no memory: ordinary pages still work, copies are then staged by the driver
        f->seq = static_cast<char *>(std::malloc(size + 64));
    }
    if (!f->seq) {
        if (map) ::munmap(const_cast<char *>(map), size);
        ::close(fd);
        return mgc::set_error(MGC_E_NOMEM, "out of host memory");
    }
    int64_t n = 0;
    const char *p = map, *end = map + size;
    bool in_record = false;
    while (p < end) {
        const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(end - p)));
        const char *le = nl ? nl : end;
        const char *a = p, *b = le;
        while (b > a && (b[-1] == '\r' || b[-1] == ' ' || b[-1] == '\t')) b--;
        if (a < b && *a == '>') {
            f->offsets.push_back(n);   // start of this record; the end of the last one is appended after the loop
            std::string def(a + 1, b);
            size_t k = 0;
            while (k < def.size() && def[k] != ' ' && def[k] != '\t') k++;
            f->ids.emplace_back(def.substr(0, k));
            f->deflines.emplace_back(std::move(def));
            in_record = true;
        } else if (in_record) {
            while (a < b && (*a == ' ' || *a == '\t')) a++;
            std::memcpy(f->seq + n, a, (size_t)(b - a));
            n += b - a;
        }
        p = nl ? nl + 1 : end;
    }
    f->offsets.push_back(n);
    if (f->ids.empty()) f->offsets.assign(1, 0);
    if (map) ::munmap(const_cast<char *>(map), size);
    ::close(fd);
    *out = f.release();
    return MGC_OK;
}

int32_t mgc_fasta_num_records(const mgc_fasta *f) { return f ? (int32_t)f->ids.size() : -1; }
const char *mgc_fasta_id(const mgc_fasta *f, int32_t i) { return f && i >= 0 && i < (int32_t)f->ids.size() ? f->ids[i].c_str() : nullptr; }
const char *mgc_fasta_defline(const mgc_fasta *f, int32_t i) {
    return f && i >= 0 && i < (int32_t)f->deflines.size() ? f->deflines[i].c_str() : nullptr;
}
const char *mgc_fasta_sequence(const mgc_fasta *f) { return f ? f->seq : nullptr; }
const int64_t *mgc_fasta_offsets(const mgc_fasta *f) { return f ? f->offsets.data() : nullptr; }
void mgc_fasta_free(mgc_fasta *f) { delete f; }
