// mgc_host.cpp — host-only parts of the C-ABI: error text, GFF/LST writers, synthetic workload.
//
// Writers reproduce gmhmmp2's Output::PrintGenes_GFF / PrintGenes_LST byte for byte except for the
// date lines (format strings: SURVEY.md §8b, pinned by tests/golden).  The reference reads the GFF
// back with pandas (bin/data.py:66-96) and caches it on disk (bin/data.py:39).
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
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

}  // namespace

extern "C" {

namespace {

// Append-only text buffer flushed to a FILE in large blocks: the gene lines are formatted by hand (decimal integers and
// "%.2f" with an exact fallback), which is ~6x faster than one fprintf per field for the ~100 bytes per gene.
// Text buffers are kept for the next file of the process (up to kTextCacheMax bytes): a fresh 8 MB block costs its page
// faults again — more than formatting into it (python -m mgcod_b200.multi writes four files per genome).
struct TextBlocks {
    struct Block { char *p; size_t bytes; };
    std::mutex mu;
    std::vector<Block> free_blocks;
    size_t cached = 0;
    static constexpr size_t kTextCacheMax = (size_t)512 << 20;
    char *acquire(size_t need, size_t *got) {
        {
            std::lock_guard<std::mutex> lk(mu);
            size_t best = free_blocks.size();
            for (size_t i = 0; i < free_blocks.size(); i++)
                if (free_blocks[i].bytes >= need && (best == free_blocks.size() || free_blocks[i].bytes < free_blocks[best].bytes)) best = i;
            if (best != free_blocks.size()) {
                const Block b = free_blocks[best];
                free_blocks.erase(free_blocks.begin() + best);
                cached -= b.bytes;
                *got = b.bytes;
                return b.p;
            }
        }
        *got = need + need / 8;
        return static_cast<char *>(std::malloc(*got));   // never zero-filled: the formatting thread touches its own pages first
    }
    void release(char *p, size_t bytes) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (cached + bytes <= kTextCacheMax) {
                free_blocks.push_back(Block{p, bytes});
                cached += bytes;
                return;
            }
        }
        std::free(p);
    }
};
TextBlocks &text_blocks() {
    static TextBlocks *b = new TextBlocks();
    return *b;
}

struct TextOut {
    FILE *f;
    char *buf = nullptr;
    size_t cap = 0, n = 0;
    explicit TextOut(FILE *fp, size_t reserve = 1 << 22) : f(fp), cap(reserve) {}
    TextOut(TextOut &&o) noexcept : f(o.f), buf(o.buf), cap(o.cap), n(o.n) { o.buf = nullptr; }
    TextOut(const TextOut &) = delete;
    ~TextOut() { text_blocks().release(buf, cap); }
    void room(size_t need) {
        if (!buf) buf = text_blocks().acquire(std::max(cap, need), &cap);
        if (n + need <= cap) return;
        if (f) flush();
        if (n + need > cap) {     // memory-only mode (f == nullptr): grow
            size_t got = 0;
            char *nb = text_blocks().acquire(std::max(cap * 2, n + need), &got);
            std::memcpy(nb, buf, n);
            text_blocks().release(buf, cap);
            buf = nb;
            cap = got;
        }
    }
    void flush() {
        if (n && f) std::fwrite(buf, 1, n, f);
        if (f) n = 0;
    }
    void str(const char *s, size_t len) {
        room(len);
        std::memcpy(buf + n, s, len);
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
    // --- unchecked appends for a line whose maximum length was reserved with line(): the gene lines are ~25 small pieces each,
    // and a capacity test per piece costs more than the formatting ---
    char *line(size_t max_len) { room(max_len); return buf + n; }
    void done(char *end) { n = (size_t)(end - buf); }
    static char *put(char *p, const char *s, size_t len) { std::memcpy(p, s, len); return p + len; }
    static char *put_u64(char *p, unsigned long long v) {
        static const char digits[201] =
            "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
        char t[24];
        int k = 24;
        while (v >= 100) { const unsigned d = (unsigned)(v % 100) * 2; v /= 100; t[--k] = digits[d + 1]; t[--k] = digits[d]; }
        if (v >= 10) { const unsigned d = (unsigned)v * 2; t[--k] = digits[d + 1]; t[--k] = digits[d]; }
        else t[--k] = (char)('0' + v);
        std::memcpy(p, t + k, (size_t)(24 - k));
        return p + (24 - k);
    }
    static char *put_i64(char *p, long long v) {
        if (v < 0) { *p++ = '-'; return put_u64(p, (unsigned long long)(-v)); }
        return put_u64(p, (unsigned long long)v);
    }
    static char *put_fixed2(char *p, double v) {   // as fixed2() below
        const double a = std::fabs(v) * 100.0;
        const double fl = std::floor(a), fr = a - fl;
        if (!(a < 9e15) || std::fabs(fr - 0.5) < 1e-6) return p + std::snprintf(p, 64, "%.2f", v);
        const unsigned long long r = (unsigned long long)fl + (fr > 0.5 ? 1ull : 0ull);
        if (std::signbit(v)) *p++ = '-';
        p = put_u64(p, r / 100);
        *p++ = '.';
        *p++ = (char)('0' + (r / 10) % 10);
        *p++ = (char)('0' + r % 10);
        return p;
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
    const std::string date = now_string();
    std::string head = lst ? "# GeneMark.hmm-2 LST format\n" : "##gff-version 2\n";
    head += "# GeneMark.hmm-2 prokaryotic version: 1.17\n";
    head += std::string("# File with sequence: ") + (seq_file_name ? seq_file_name : "") + "\n";
    head += std::string("# File with MetaGeneMark parameters: ") + (mod_file_name ? mod_file_name : "") + "\n";
    head += "# translation table: " + std::to_string(gcode) + "\n";
    head += "# output date start: " + date + "\n";
    // The records are formatted by several threads, each a contiguous range of contigs with about the same number of genes,
    // into memory; the pieces are written in order.  gene_id counts across records, so a piece starts at 1 + the number of
    // genes in front of it (the table is ordered by contig).
    std::vector<int64_t> first_gene((size_t)n_contigs + 1, ng);     // first gene of every contig
    {
        int64_t k = 0;
        for (int32_t c = 0; c < n_contigs; c++) {
            first_gene[c] = k;
            while (k < ng && genes[k].contig == c) k++;
        }
        first_gene[n_contigs] = k;
    }
    int pieces = (int)std::min<int64_t>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency() / 4), 8u), ng / 20000 + 1);
    if (const char *env = std::getenv("MGCOD_B200_WRITER_THREADS")) pieces = std::max(1, std::atoi(env));
    std::vector<int32_t> cut((size_t)pieces + 1, n_contigs);
    cut[0] = 0;
    for (int t = 1; t < pieces; t++) {
        const int64_t want = ng * t / pieces;
        cut[t] = (int32_t)(std::lower_bound(first_gene.begin(), first_gene.end(), want) - first_gene.begin());
        cut[t] = std::min<int32_t>(std::max(cut[t], cut[t - 1]), n_contigs);
    }
    std::vector<TextOut> outs;
    outs.reserve(pieces);
    for (int t = 0; t < pieces; t++) outs.emplace_back(nullptr, (size_t)((first_gene[cut[t + 1]] - first_gene[cut[t]]) * 112 + (cut[t + 1] - cut[t]) * 160 + 4096));
    auto format_range = [&](int t) {
        TextOut &o = outs[t];
        char tmp[256];
        int64_t k = first_gene[cut[t]];
        long gene_id = (long)k;
        for (int32_t c = cut[t]; c < cut[t + 1]; c++) {
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
                    char *q = o.line(idlen + 256);
                    q = TextOut::put(q, id, idlen); q = TextOut::put(q, "\tGeneMark.hmm2\tCDS\t", sizeof("\tGeneMark.hmm2\tCDS\t") - 1); q = TextOut::put_i64(q, g.left); *q++ = '\t';
                    q = TextOut::put_i64(q, g.right); *q++ = '\t'; q = TextOut::put_fixed2(q, g.logodd);
                    *q++ = '\t'; *q++ = g.strand > 0 ? '+' : '-'; q = TextOut::put(q, "\t0\tgene_id ", sizeof("\t0\tgene_id ") - 1); q = TextOut::put_i64(q, gene_id);
                    q = g.domain ? TextOut::put(q, "; gene_type archaea; ", sizeof("; gene_type archaea; ") - 1) : TextOut::put(q, "; gene_type bacteria; ", sizeof("; gene_type bacteria; ") - 1);
                    if (g.partial_left || g.partial_right) { q = TextOut::put(q, "partial ", sizeof("partial ") - 1); *q++ = g.partial_left ? '1' : '0'; *q++ = g.partial_right ? '1' : '0'; q = TextOut::put(q, "; ", sizeof("; ") - 1); }
                    q = TextOut::put(q, "gc ", sizeof("gc ") - 1); q = TextOut::put_i64(q, g.gc); q = TextOut::put(q, "; length ", sizeof("; length ") - 1); q = TextOut::put_i64(q, len); q = TextOut::put(q, ";\n", sizeof(";\n") - 1);
                    o.done(q);
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
    };
    const bool trace = std::getenv("MGCOD_B200_TRACE") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t0 = now();
    if (pieces == 1) format_range(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < pieces; t++) pool.emplace_back(format_range, t);
        for (auto &th : pool) th.join();
    }
    const double t1 = now();
    // plain write(2) of header, pieces, footer (buffered writes to ONE file are serialised by the kernel: writing the pieces
    // side by side with pwrite was slower; the four files of a run are written side by side by their callers)
    std::string foot;
    if (lst) foot = "\n# output date end: " + date + "\n# Done with file: " + (seq_file_name ? seq_file_name : "") + "\n";
    const int fd = ::open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) return mgc::set_error(MGC_E_IO, std::string("cannot open output file: ") + out_path);
    auto put_all = [fd](const char *p, size_t n) {
        while (n) {
            const ssize_t k = ::write(fd, p, n);
            if (k <= 0) return false;
            p += k; n -= (size_t)k;
        }
        return true;
    };
    bool ok = put_all(head.data(), head.size());
    for (int t = 0; t < pieces && ok; t++) ok = put_all(outs[t].buf, outs[t].n);
    ok = ok && put_all(foot.data(), foot.size());
    if (trace) std::fprintf(stderr, "write_gff: %d pieces, format %.2f ms, write %.2f ms\n", pieces, t1 - t0, now() - t1);
    if (::close(fd) != 0 || !ok) return mgc::set_error(MGC_E_IO, "write failed");
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

// Page-locked letter buffers are kept for the next file of the process (python -m mgcod_b200.multi reads thousands):
// cudaHostAlloc costs ~0.3 ms per MB, more than reading the file.  At most kBufCacheMax bytes stay cached.
namespace {
struct LetterBuffers {
    struct Block { char *p; size_t bytes; };
    std::mutex mu;
    std::vector<Block> free_blocks;
    size_t cached = 0;
    static constexpr size_t kBufCacheMax = (size_t)6 << 30;
    char *acquire(size_t need, size_t *got) {
        {
            std::lock_guard<std::mutex> lk(mu);
            size_t best = free_blocks.size();
            for (size_t i = 0; i < free_blocks.size(); i++)
                if (free_blocks[i].bytes >= need && (best == free_blocks.size() || free_blocks[i].bytes < free_blocks[best].bytes)) best = i;
            if (best != free_blocks.size()) {
                Block b = free_blocks[best];
                free_blocks.erase(free_blocks.begin() + best);
                cached -= b.bytes;
                *got = b.bytes;
                return b.p;
            }
        }
        char *p = nullptr;
        const size_t bytes = need + need / 8 + 4096;
        if (cudaHostAlloc(reinterpret_cast<void **>(&p), bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        *got = bytes;
        return p;
    }
    void release(char *p, size_t bytes) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (cached + bytes <= kBufCacheMax) {
                free_blocks.push_back(Block{p, bytes});
                cached += bytes;
                return;
            }
        }
        cudaFreeHost(p);
    }
};
LetterBuffers &letter_buffers() {
    static LetterBuffers *b = new LetterBuffers();   // never destroyed: the CUDA runtime may be gone at exit
    return *b;
}
}  // namespace

struct mgc_fasta {
    std::vector<std::string> ids, deflines;
    std::string ids_joined, deflines_joined;   // '\n'-separated, for one-call retrieval
    std::vector<const char *> id_ptrs;         // mgc_fasta_id_array, built on first use
    std::mutex mu;
    std::vector<int64_t> offsets;
    char *seq = nullptr;
    size_t seq_bytes = 0;
    bool pinned = false;
    ~mgc_fasta() {
        if (seq) {
            if (pinned) letter_buffers().release(seq, seq_bytes);
            else std::free(seq);
        }
    }
};

// Multi-threaded reader.  Three passes, the last two spread over the host cores by bytes:
//   1. record starts: '>' at the beginning of a line (memchr over disjoint byte ranges);
//   2. letters per record (bytes above 0x20 in the body) -> offsets by a prefix sum;
//   3. the bodies copied line by line into their final place, whitespace removed.
// Text semantics follow Biopython's FASTA parser (bin/data.py:117 uses SeqIO.parse): the id is the first
// whitespace-delimited token of the defline — "> id x" has the id "id" —, line ends and blanks anywhere in the sequence
// are dropped, text in front of the first '>' is ignored.
namespace {

constexpr size_t kPinnedMax = (size_t)8 << 30;   // larger letter buffers stay in ordinary pages (the copy is then staged)

inline bool is_blank(unsigned char c) { return c <= 0x20; }

int64_t count_letters(const char *a, const char *b) {
    int64_t n = 0;
    for (const char *p = a; p < b; p++) n += (unsigned char)*p > 0x20;
    return n;
}

// copies [a, b) without blanks, line by line (memchr + memcpy; a line holding an interior blank is compacted bytewise)
char *copy_letters(const char *a, const char *b, char *dst) {
    while (a < b) {
        const char *nl = static_cast<const char *>(std::memchr(a, '\n', (size_t)(b - a)));
        const char *le = nl ? nl : b;
        const char *x = a, *y = le;
        while (y > x && is_blank((unsigned char)y[-1])) y--;
        while (x < y && is_blank((unsigned char)*x)) x++;
        const size_t len = (size_t)(y - x);
        if (len) {
            if ((size_t)count_letters(x, y) != len) {   // a blank inside the line
                for (const char *p = x; p < y; p++) {
                    *dst = *p;
                    dst += (unsigned char)*p > 0x20;
                }
            } else {
                std::memcpy(dst, x, len);
                dst += len;
            }
        }
        a = nl ? nl + 1 : b;
    }
    return dst;
}

template <class F> void parallel_ranges(int threads, F fn) {
    if (threads <= 1) { fn(0); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(fn, t);
    for (auto &th : pool) th.join();
}

}  // namespace

int mgc_fasta_read(const char *path, int page_locked, mgc_fasta **out) {
    if (!path || !out) return mgc::set_error(MGC_E_ARG, "null argument");
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return mgc::set_error(MGC_E_IO, std::string("error, cannot open file: ") + path);
    struct stat sb;
    if (::fstat(fd, &sb) != 0) { ::close(fd); return mgc::set_error(MGC_E_IO, std::string("cannot stat ") + path); }
    const size_t size = (size_t)sb.st_size;
    const char *map = nullptr;
    if (size) {
        map = static_cast<const char *>(::mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0));
        if (map == MAP_FAILED) { ::close(fd); return mgc::set_error(MGC_E_IO, std::string("cannot map ") + path); }
    }
    auto unmap = [&] {
        if (map) ::munmap(const_cast<char *>(map), size);
        ::close(fd);
    };
    (void)unmap;
    const bool trace = std::getenv("MGCOD_B200_TRACE") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t0 = now();
    std::unique_ptr<mgc_fasta> f(new mgc_fasta());
    int threads = (int)std::min<size_t>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u), size / (4u << 20) + 1);
    if (const char *env = std::getenv("MGCOD_B200_FASTA_THREADS")) threads = std::max(1, std::atoi(env));
    // pass 1: record starts
    std::vector<std::vector<size_t>> found(threads);
    parallel_ranges(threads, [&](int t) {
        const size_t lo = size * (size_t)t / threads, hi = size * (size_t)(t + 1) / threads;
        const char *p = map + lo, *end = map + hi;
        while (p < end) {
            p = static_cast<const char *>(std::memchr(p, '>', (size_t)(end - p)));
            if (!p) break;
            if (p == map || p[-1] == '\n') found[t].push_back((size_t)(p - map));
            p++;
        }
    });
    const double t1 = now();
    std::vector<size_t> head;
    for (auto &v : found) head.insert(head.end(), v.begin(), v.end());
    const size_t n_rec = head.size();
    std::vector<size_t> body(n_rec);           // first byte behind the defline
    f->ids.resize(n_rec);
    f->deflines.resize(n_rec);
    for (size_t i = 0; i < n_rec; i++) {
        const char *a = map + head[i] + 1, *lim = map + (i + 1 < n_rec ? head[i + 1] : size);
        const char *nl = static_cast<const char *>(std::memchr(a, '\n', (size_t)(lim - a)));
        const char *b = nl ? nl : lim;
        body[i] = (size_t)((nl ? nl + 1 : lim) - map);
        while (b > a && is_blank((unsigned char)b[-1])) b--;
        f->deflines[i].assign(a, b);
        const char *x = a;
        while (x < b && (*x == ' ' || *x == '\t')) x++;
        const char *y = x;
        while (y < b && *y != ' ' && *y != '\t') y++;
        f->ids[i].assign(x, y);
    }
    // work split of passes 2 and 3: record ranges of about equal bytes
    std::vector<size_t> cut(threads + 1, n_rec);
    cut[0] = 0;
    for (int t = 1; t < threads; t++) {
        const size_t want = size * (size_t)t / threads;
        cut[t] = (size_t)(std::lower_bound(head.begin(), head.end(), want) - head.begin());
    }
    // pass 2: letters per record
    const double t2 = now();
    f->offsets.assign(n_rec + 1, 0);
    parallel_ranges(threads, [&](int t) {
        for (size_t i = cut[t]; i < cut[t + 1]; i++)
            f->offsets[i + 1] = count_letters(map + body[i], map + (i + 1 < n_rec ? head[i + 1] : size));
    });
    for (size_t i = 0; i < n_rec; i++) f->offsets[i + 1] += f->offsets[i];
    const size_t total = (size_t)f->offsets[n_rec];
    const double t3 = now();
    if (page_locked && total + 64 <= kPinnedMax) f->seq = letter_buffers().acquire(total + 64, &f->seq_bytes);
    if (f->seq) f->pinned = true;
    else f->seq = static_cast<char *>(std::malloc(total + 64));   // no device / no memory / too large: ordinary pages still work,
                                                                   // copies are then staged by the driver
    if (!f->seq) {
        unmap();
        return mgc::set_error(MGC_E_NOMEM, "out of host memory");
    }
    // pass 3: bodies into place
    const double t4 = now();
    bool ok = true;
    parallel_ranges(threads, [&](int t) {
        for (size_t i = cut[t]; i < cut[t + 1]; i++) {
            char *dst = f->seq + f->offsets[i];
            char *e = copy_letters(map + body[i], map + (i + 1 < n_rec ? head[i + 1] : size), dst);
            if (e != f->seq + f->offsets[i + 1]) ok = false;
        }
    });
    const double t5 = now();
    for (size_t i = 0; i < n_rec; i++) {
        f->ids_joined += f->ids[i];
        f->ids_joined += '\n';
        f->deflines_joined += f->deflines[i];
        f->deflines_joined += '\n';
    }
    std::thread([map, size, fd] {     // returning the mapping takes milliseconds for a large file: not on the caller's time
        if (map) ::munmap(const_cast<char *>(map), size);
        ::close(fd);
    }).detach();
    map = nullptr;
    if (trace)
        std::fprintf(stderr, "fasta_read: %d threads, map+starts %.2f ms, deflines %.2f, count %.2f, alloc %.2f, copy %.2f, unmap %.2f\n", threads,
                     t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, now() - t5);
    if (!ok) return mgc::set_error(MGC_E_IO, "FASTA reader: letter counts of the two passes differ (file changed while reading?)");
    *out = f.release();
    return MGC_OK;
}

int32_t mgc_fasta_num_records(const mgc_fasta *f) { return f ? (int32_t)f->ids.size() : -1; }
const char *mgc_fasta_id(const mgc_fasta *f, int32_t i) { return f && i >= 0 && i < (int32_t)f->ids.size() ? f->ids[i].c_str() : nullptr; }
const char *mgc_fasta_defline(const mgc_fasta *f, int32_t i) {
    return f && i >= 0 && i < (int32_t)f->deflines.size() ? f->deflines[i].c_str() : nullptr;
}
const char *mgc_fasta_ids_joined(const mgc_fasta *f, int deflines, int64_t *len) {
    if (!f) return nullptr;
    const std::string &s = deflines ? f->deflines_joined : f->ids_joined;
    if (len) *len = (int64_t)s.size();
    return s.data();
}
const char *const *mgc_fasta_id_array(const mgc_fasta *cf) {
    mgc_fasta *f = const_cast<mgc_fasta *>(cf);
    if (!f) return nullptr;
    std::lock_guard<std::mutex> lk(f->mu);
    if (f->id_ptrs.size() != f->ids.size()) {
        f->id_ptrs.resize(f->ids.size());
        for (size_t i = 0; i < f->ids.size(); i++) f->id_ptrs[i] = f->ids[i].c_str();
    }
    return f->id_ptrs.data();
}
const char *mgc_fasta_sequence(const mgc_fasta *f) { return f ? f->seq : nullptr; }
const int64_t *mgc_fasta_offsets(const mgc_fasta *f) { return f ? f->offsets.data() : nullptr; }
void mgc_fasta_free(mgc_fasta *f) { delete f; }
