// mgc_model.hpp — parsed MetaGeneMark parameter file and the flattened tables the kernels read.
//
// Replaces gmhmmp2's ParameterParser / Pset / Model::CalculateLogOdd for the `-M file.mod` path
// (call site in the reference: bin/data.py:40-53).  Grammar and arithmetic: SURVEY.md §A.2-A.4,
// DESIGN.md §"Model tables".
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mgc {

constexpr int kMaxOrder = 5;  // tables are 3*4^(k+1) doubles per section; order 5 = 96 KB

struct SectionScalars {
    double dc, dn;          // $COD_DURATION_DECAY, $NON_DURATION_DECAY
    double dur_const;       // ln 1.5 - 3 ln Dc + ln(exp(1/Dn) - 1)
    double dur_slope;       // 1/Dn - 1/Dc
    double start_lo[3];     // ln(P(ATG|GTG|TTG) / Pnon(codon))
    double stop_lo[3];      // ln(P(TAA|TAG|TGA) / Pnon(codon))
    double ln_pn;           // ln($COD_P_N / $NON_P_N)
    double ln_cpn;          // ln($COD_P_N)
    double tr;              // ln(to_mgm * to_atypical of the section's domain): ln 0.25 unless the file sets $TO_MGM /
                            // $TO_ATYPICAL_FIRST_BACTERIA / $TO_ATYPICAL_SECOND_ARCHAEA (file-level values, DESIGN.md §2 rule 9)
};

// One parsed file.  Section slots are numbered 0..nsec-1 in the order (domain, gc).
struct Model {
    int gcode = 0, cod_order = -1, non_order = -1, min_len = 0;
    int n_bac = 0, n_arc = 0;
    int nsec = 0;
    int map[2][101];                    // gc -> slot serving it (next higher provided bin, else highest); -1 = none
    // Flattened, "plane-major" k-mer index: for a window of n letters (oldest letter = bit 0),
    //   index = (hi_bits << n) | lo_bits   with letter code = 2*hi + lo, A=0 C=1 G=2 T=3.
    std::vector<double> lc_full;        // [nsec][3][4^(kc+1)]   ln P_cod(last | first kc), phase = codon position of the first letter
    std::vector<double> ln_full;        // [nsec][4^(kn+1)]      ln P_non(last | first kn)
    std::vector<double> lc_ramp;        // [nsec][kc][3][4^kc]   order j uses the first 4^(j+1) entries of row j
    std::vector<double> ln_ramp;        // [nsec][max(kn,1)][4^max(kn,1)]
    std::vector<SectionScalars> sc;     // [nsec]
    std::string path;
};

// Throws std::runtime_error with gmhmmp2-style messages ("error, parameter not found: $KEY", ...).
Model parse_model_file(const std::string &path);

inline long pow4(int k) { return 1L << (2 * k); }

// natural (base-4, oldest letter most significant) index -> plane-major index for an n-letter window
inline uint32_t plane_index(uint32_t natural, int n) {
    uint32_t hi = 0, lo = 0;
    for (int r = 0; r < n; r++) {               // r = 0 is the oldest letter
        uint32_t code = (natural >> (2 * (n - 1 - r))) & 3u;
        hi |= (code >> 1) << r;
        lo |= (code & 1u) << r;
    }
    return (hi << n) | lo;
}

}  // namespace mgc
