// mgc_model.cpp — .mod parser and log-table builder (host side of the engine).
// See mgc_model.hpp for what it replaces in the reference.
#include "mgc_model.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace mgc {
namespace {

struct RawSection {
    int domain = 0, gc = 0;
    std::map<std::string, double> num;
    std::vector<double> cod;  // [4^(kc+1)][3] natural index
    std::vector<double> non;  // [4^(kn+1)]
    bool has_cod = false, has_non = false;
    std::string label;        // as written ("__B_GC_35"): the binary visits sections in the lexicographic order of these
    long gc_dist_values = 0;  // lines seen under $GENE_GC_DIST_MAT
};

int letter(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

long natural_index(const std::string &kmer) {
    long v = 0;
    for (char c : kmer) {
        int x = letter(c);
        if (x < 0) return -1;
        v = v * 4 + x;
    }
    return v;
}

double need(const RawSection &s, const char *key) {
    auto it = s.num.find(key);
    if (it == s.num.end()) throw std::runtime_error(std::string("error, parameter not found: $") + key);
    return it->second;
}

double opt(const RawSection &s, const char *key, double dflt) {
    auto it = s.num.find(key);
    return it == s.num.end() ? dflt : it->second;
}

// joint table of (n+1)-mers -> joint table of n-mers by summing out the LAST letter (prefix marginal);
// `cols` interleaved columns per k-mer.
std::vector<double> prefix_marginal(const std::vector<double> &t, int cols) {
    std::vector<double> r(t.size() / 4, 0.0);
    long n = (long)(t.size() / cols / 4);
    for (long i = 0; i < n; i++)
        for (int x = 0; x < 4; x++)
            for (int c = 0; c < cols; c++) r[i * cols + c] += t[(i * 4 + x) * cols + c];
    return r;
}

}  // namespace

Model parse_model_file(const std::string &path) {
    // the whole file at once, then tokens by hand: a real mgm_*.mod is ~100 000 lines of "KMER p1 p2 p3", and one
    // istringstream per line cost more than everything else the engine does for a small genome
    std::string text;
    {
        std::ifstream in(path, std::ios::binary);
        if (!in) throw std::runtime_error("error, cannot open file: " + path);
        in.seekg(0, std::ios::end);
        const std::streamoff n = in.tellg();
        in.seekg(0, std::ios::beg);
        text.resize(n > 0 ? (size_t)n : 0);
        if (n > 0) in.read(&text[0], n);
    }
    std::vector<RawSection> raw;
    raw.reserve(256);          // cur points into it: no reallocation while parsing (sections beyond are appended by index)
    size_t cur_i = (size_t)-1;
    enum { SCALARS, NON, COD, SKIP, GC_DIST } mode = SCALARS;
    int want = 0;              // k-mer length of the matrix being read
    const char *p = text.data(), *end = p + text.size();
    auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\r'; };
    while (p < end) {
        const char *le = static_cast<const char *>(std::memchr(p, '\n', (size_t)(end - p)));
        if (!le) le = end;
        const char *a = p;
        p = le < end ? le + 1 : end;
        while (a < le && blank(*a)) a++;
        if (a == le) continue;
        const char *b = a;
        while (b < le && !blank(*b)) b++;          // first token [a, b)
        RawSection *cur = cur_i == (size_t)-1 ? nullptr : &raw[cur_i];
        if (b - a >= 2 && a[0] == '_' && a[1] == '_') {
            const std::string tok(a, b);
            cur_i = (size_t)-1;
            mode = SCALARS;
            int dom = -1;
            if (tok.rfind("__B_GC_", 0) == 0) dom = 0;
            else if (tok.rfind("__A_GC_", 0) == 0) dom = 1;
            if (dom >= 0) {
                int gc = std::atoi(tok.c_str() + 7);
                if (gc < 0 || gc > 100) throw std::runtime_error("error, GC bin out of range: " + tok);
                raw.emplace_back();
                cur_i = raw.size() - 1;
                raw[cur_i].domain = dom;
                raw[cur_i].gc = gc;
                raw[cur_i].label = tok;
            }
            continue;  // other sections (__NATIVE, __START, ...) are not on the -M path
        }
        if (!cur) continue;
        if (*a == '$') {
            const std::string key(a + 1, b);
            mode = SCALARS;
            if (key == "NON_MAT") {
                int kn = (int)need(*cur, "NON_ORDER");
                if (kn < 0 || kn > kMaxOrder) throw std::runtime_error("error, $NON_ORDER outside 0..5");
                cur->non.assign(pow4(kn + 1), 0.0);
                cur->has_non = true;
                mode = NON;
                want = kn + 1;
            } else if (key == "COD_MAT") {
                int kc = (int)need(*cur, "COD_ORDER");
                if (kc < 0 || kc > kMaxOrder) throw std::runtime_error("error, $COD_ORDER outside 0..5");
                cur->cod.assign(pow4(kc + 1) * 3, 0.0);
                cur->has_cod = true;
                mode = COD;
                want = kc + 1;
            } else if (key == "GENE_GC_DIST_MAT") {
                mode = GC_DIST;
            } else if (key.size() > 4 && key.compare(key.size() - 4, 4, "_MAT") == 0) {
                mode = SKIP;
            } else {
                const char *v = b;
                while (v < le && blank(*v)) v++;
                if (v < le) {
                    const std::string val(v, le);      // strtod needs a terminated string
                    char *stop = nullptr;
                    const double x = std::strtod(val.c_str(), &stop);
                    if (stop != val.c_str()) cur->num[key] = x;
                }
            }
            continue;
        }
        if (mode == GC_DIST) { cur->gc_dist_values++; continue; }
        if (mode == NON || mode == COD) {
            long id = 0;
            bool ok = (int)(b - a) == want;
            for (const char *q = a; q < b && ok; q++) {
                const int x = letter(*q);
                ok = x >= 0;
                id = id * 4 + x;
            }
            if (!ok) throw std::runtime_error("error, label has unexpected length");
            // the values: up to three numbers behind the k-mer; the line ends at a newline, which stops strtod
            const char *v = b;
            const int cols = mode == NON ? 1 : 3;
            for (int c = 0; c < cols; c++) {
                while (v < le && blank(*v)) v++;
                char *stop = nullptr;
                const double x = v < le ? std::strtod(v, &stop) : 0.0;     // text ends with the file's own bytes; a number never runs past '\n'
                if (v >= le || stop == v || stop > le) throw std::runtime_error(mode == NON ? "error, missing value in $NON_MAT" : "error, missing value in $COD_MAT");
                if (mode == NON) cur->non[id] = x;
                else cur->cod[id * 3 + c] = x;
                v = stop;
            }
        }
    }
    if (raw.empty()) throw std::runtime_error("error, no __B_GC_/__A_GC_ section in " + path);

    Model m;
    m.path = path;
    // slots ordered by (domain, gc)
    std::vector<const RawSection *> slots;
    for (int d = 0; d < 2; d++)
        for (int g = 0; g <= 100; g++)
            for (const auto &s : raw)
                if (s.domain == d && s.gc == g) slots.push_back(&s);  // a repeated label: last one wins below
    for (int d = 0; d < 2; d++)
        for (int g = 0; g <= 100; g++) m.map[d][g] = -1;
    {
        std::vector<const RawSection *> uniq;
        for (const RawSection *s : slots) {
            if (!uniq.empty() && uniq.back()->domain == s->domain && uniq.back()->gc == s->gc) uniq.back() = s;
            else uniq.push_back(s);
        }
        slots.swap(uniq);
    }
    m.nsec = (int)slots.size();
    const RawSection &first = *slots[0];
    m.gcode = (int)need(first, "GCODE");
    m.cod_order = (int)need(first, "COD_ORDER");
    m.non_order = (int)need(first, "NON_ORDER");
    m.min_len = (int)need(first, "GENE_MIN_LENGTH");
    switch (m.gcode) {
        case 11: case 4: case 25: case 15: case 101: case 1: break;
        default: throw std::runtime_error("error, unsupported genetic code found");
    }
    if (m.cod_order > kMaxOrder || m.cod_order < 0) throw std::runtime_error("error, $COD_ORDER outside 0..5");
    if (m.non_order > m.cod_order || m.non_order < 0) throw std::runtime_error("error, $NON_ORDER must be in 0..$COD_ORDER");
    const int kc = m.cod_order, kn = m.non_order;
    const long nfull = pow4(kc + 1), nnon = pow4(kn + 1);
    const int kr = kc > 0 ? kc : 1, knr = kn > 0 ? kn : 1;
    const long nramp = pow4(kr), nnramp = pow4(knr);
    m.lc_full.assign((size_t)m.nsec * 3 * nfull, 0.0);
    m.ln_full.assign((size_t)m.nsec * nnon, 0.0);
    m.lc_ramp.assign((size_t)m.nsec * kr * 3 * nramp, 0.0);
    m.ln_ramp.assign((size_t)m.nsec * knr * nnramp, 0.0);
    m.sc.resize(m.nsec);

    for (int si = 0; si < m.nsec; si++) {
        const RawSection &s = *slots[si];
        if ((int)need(s, "COD_ORDER") != kc || (int)need(s, "NON_ORDER") != kn || (int)need(s, "GCODE") != m.gcode)
            throw std::runtime_error("error, sections of one file must share $GCODE, $COD_ORDER and $NON_ORDER");
        if (!s.has_cod) throw std::runtime_error("error, parameter not found: $COD_MAT");
        if (!s.has_non) throw std::runtime_error("error, parameter not found: $NON_MAT");
        (s.domain == 0 ? m.n_bac : m.n_arc)++;

        // joint tables for every window length
        std::vector<std::vector<double>> cj(kc + 2), nj(kn + 2);
        cj[kc + 1] = s.cod;
        for (int n = kc; n >= 1; n--) cj[n] = prefix_marginal(cj[n + 1], 3);
        nj[kn + 1] = s.non;
        for (int n = kn; n >= 1; n--) nj[n] = prefix_marginal(nj[n + 1], 1);

        auto cond_cod = [&](int order, long kmer, int phase) {  // kmer has order+1 letters
            double den;
            if (order == 0) {
                den = 0;
                for (int x = 0; x < 4; x++) den += cj[1][x * 3 + phase];
            } else den = cj[order][(kmer >> 2) * 3 + phase];
            return std::log(cj[order + 1][kmer * 3 + phase] / den);
        };
        auto cond_non = [&](int order, long kmer) {
            double den;
            if (order == 0) {
                den = 0;
                for (int x = 0; x < 4; x++) den += nj[1][x];
            } else den = nj[order][kmer >> 2];
            return std::log(nj[order + 1][kmer] / den);
        };
        for (int ph = 0; ph < 3; ph++)
            for (long k = 0; k < nfull; k++)
                m.lc_full[((size_t)si * 3 + ph) * nfull + plane_index((uint32_t)k, kc + 1)] = cond_cod(kc, k, ph);
        for (long k = 0; k < nnon; k++) m.ln_full[(size_t)si * nnon + plane_index((uint32_t)k, kn + 1)] = cond_non(kn, k);
        for (int j = 0; j < kc; j++)
            for (int ph = 0; ph < 3; ph++)
                for (long k = 0; k < pow4(j + 1); k++)
                    m.lc_ramp[(((size_t)si * kr + j) * 3 + ph) * nramp + plane_index((uint32_t)k, j + 1)] = cond_cod(j, k, ph);
        for (int j = 0; j < kn; j++)
            for (long k = 0; k < pow4(j + 1); k++)
                m.ln_ramp[((size_t)si * knr + j) * nnramp + plane_index((uint32_t)k, j + 1)] = cond_non(j, k);

        // probability of a codon under the non-coding model (DESIGN.md §"start/stop log-odds")
        double tot1 = 0, p1[4];
        for (int x = 0; x < 4; x++) tot1 += nj[1][x];
        for (int x = 0; x < 4; x++) p1[x] = nj[1][x] / tot1;
        auto non_codon = [&](int c) {
            int x0 = c >> 4, x1 = (c >> 2) & 3, x2 = c & 3;
            int rc = (3 - x2) * 16 + (3 - x1) * 4 + (3 - x0);
            if (kn == 2) return nj[3][c] / tot1;
            if (kn > 2) return 0.5 * (nj[3][c] + nj[3][rc]) / tot1;
            if (kn == 1) return 0.125 * (p1[x0] + p1[3 - x0]) * (p1[x1] + p1[3 - x1]) * (p1[x2] + p1[3 - x2]);
            return p1[x0] * p1[x1] * p1[x2];
        };
        SectionScalars &sc = m.sc[si];
        sc.dc = need(s, "COD_DURATION_DECAY");
        sc.dn = need(s, "NON_DURATION_DECAY");
        sc.dur_const = std::log(1.5) - 3.0 * std::log(sc.dc) + std::log(std::exp(1.0 / sc.dn) - 1.0);
        sc.dur_slope = 1.0 / sc.dn - 1.0 / sc.dc;
        const int start_codon[3] = {14, 46, 62};  // ATG GTG TTG
        const char *start_key[3] = {"ATG", "GTG", "TTG"};
        const int stop_codon[3] = {48, 50, 56};   // TAA TAG TGA
        const char *stop_key[3] = {"TAA", "TAG", "TGA"};
        for (int t = 0; t < 3; t++) {
            sc.start_lo[t] = std::log(opt(s, start_key[t], 0.0) / non_codon(start_codon[t]));
            sc.stop_lo[t] = std::log(opt(s, stop_key[t], 0.0) / non_codon(stop_codon[t]));
        }
        sc.ln_pn = std::log(need(s, "COD_P_N") / need(s, "NON_P_N"));
        sc.ln_cpn = std::log(need(s, "COD_P_N"));
    }
    // Optional keys (each probed against the binary, DESIGN.md §2 rule 9).  $TO_MGM, $TO_ATYPICAL_FIRST_BACTERIA and
    // $TO_ATYPICAL_SECOND_ARCHAEA are FILE-level: the binary walks its sections in the lexicographic order of their labels
    // ("__A_GC_35" < "__B_GC_100" < "__B_GC_35" < "__B_GC_9") and every section that carries a key overwrites the value.
    // $GENE_GC_DIST only has its matrix checked (101 entries), it does not change the -M path's output; $TO_NATIVE, $BUILD
    // and unknown keys are ignored; a non-zero $RBS / $PROMOTER / $SC_RBS / $SC_PROMOTER switches a site model on, whose
    // scores this build does not reproduce: refused rather than scored silently wrong.
    {
        std::vector<const RawSection *> order(slots.begin(), slots.end());
        std::sort(order.begin(), order.end(), [](const RawSection *a, const RawSection *b) { return a->label < b->label; });
        double to_mgm = 0.5, to_atyp[2] = {0.5, 0.5};
        const char *names[3] = {"TO_MGM", "TO_ATYPICAL_FIRST_BACTERIA", "TO_ATYPICAL_SECOND_ARCHAEA"};
        for (const RawSection *s : order) {
            for (const char *k : {"RBS", "PROMOTER", "SC_RBS", "SC_PROMOTER"})
                if (opt(*s, k, 0.0) != 0.0)
                    throw std::runtime_error(std::string("error, $") + k + " site models are not supported on the -M path of this build");
            if (opt(*s, "GENE_GC_DIST", 0.0) != 0.0) {
                if (s->gc_dist_values == 0) throw std::runtime_error("error, parameter not found: $GENE_GC_DIST_MAT");
                if (s->gc_dist_values != 101)
                    throw std::runtime_error("error, diff between expected and observed number of values for key $GENE_GC_DIST_MAT 101 " +
                                             std::to_string(s->gc_dist_values));
            }
            auto it = s->num.find(names[0]);
            if (it != s->num.end()) to_mgm = it->second;
            for (int d = 0; d < 2; d++) {
                it = s->num.find(names[1 + d]);
                if (it != s->num.end()) to_atyp[d] = it->second;
            }
        }
        const double vals[3] = {to_mgm, to_atyp[0], to_atyp[1]};
        for (int i = 0; i < 3; i++)
            if (vals[i] < 0) {
                char buf[160];
                std::snprintf(buf, sizeof buf, "error in value of the parameter with key: $%s %g", names[i], vals[i]);
                throw std::runtime_error(buf);
            }
        for (int si = 0; si < m.nsec; si++) m.sc[si].tr = std::log(to_mgm * to_atyp[slots[si]->domain]);
    }
    // a section serves the GC values up to its label: missing bin -> next higher provided; above all -> highest
    for (int d = 0; d < 2; d++) {
        int next = -1, highest = -1;
        for (int si = 0; si < m.nsec; si++)
            if (slots[si]->domain == d) highest = si;  // slots ascend in gc within a domain
        for (int g = 100; g >= 0; g--) {
            for (int si = 0; si < m.nsec; si++)
                if (slots[si]->domain == d && slots[si]->gc == g) next = si;
            m.map[d][g] = next >= 0 ? next : highest;
        }
    }
    return m;
}

}  // namespace mgc
