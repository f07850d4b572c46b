// hamiltonian.hpp — host-side Hamiltonian container of libsgs (the product, not the oracle).
// Mirrors the reference's Hamiltonian (cpp/state.hpp:17-44): terms in file order
// (`original_paulis`), bucketed by last site (`paulis`), and the flattened paulis_list order
// (cpp/state.cpp:88-96).  Terms are stored packed (bit 2j = z_j, bit 2j+1 = x_j).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sgs {

constexpr int kMaxWords = 4;   // 128 qubit positions

struct PackedRow {
    uint64_t w[kMaxWords] = {0, 0, 0, 0};
};

constexpr int kCompactSites = 32;     // a compact row holds a term that spans at most 32 sites

struct Hamiltonian {
    int n = 0, k = 0, l = 0;            // l: period (periodic files), else 0
    int W = 1;                          // words per absolute row = ceil(2 * n / 64) (any n)
    bool wide = false;                  // n + 1 > 128: absolute rows do not fit a PackedRow; every term must be compact
    std::vector<PackedRow> terms;       // file order, absolute rows (all zero in a wide Hamiltonian)
    std::vector<uint64_t> crow;         // file order, compact rows: bit 2j = z, 2j+1 = x of site begin + j (0 if the term spans > 32 sites)
    std::vector<uint8_t> compact;       // 1 if crow[t] holds the term
    std::vector<double> weights;
    std::vector<int> begin, end;        // term windows [min site, max site + 1)
    std::vector<int> order;             // paulis_list order → file index
    std::vector<int> bucket_off;        // n + 1 offsets into `order` by last site

    int T() const { return (int)weights.size(); }
    void add_term(const PackedRow &r, double w, int b, int e);               // n + 1 <= 128 only
    bool add_term_sites(const std::vector<int> &sites, const std::vector<int> &codes, double w);   // any n; codes: 0 I, 1 Z, 2 X, 3 Y
    void row_words(int t, uint64_t *out, int Wout) const;                    // absolute row of term t as Wout words
    void finalize();                    // builds order / bucket_off / W
};

struct ParseError {
    int code;
    std::string msg;
};

// Hamiltonian::from_file / from_file_periodic (cpp/state.cpp:58-81, py/state.py:34-72)
bool parse_hamiltonian_text(const std::string &text, bool periodic, int cmax, Hamiltonian &H, ParseError &err);
bool load_hamiltonian_file(const std::string &path, bool periodic, int cmax, Hamiltonian &H, ParseError &err);

// letters over [begin,end) with sign, e.g. "+XIZY" (cpp/stab.cpp:347-364)
std::string row_to_string(const uint64_t *row, int sign, int begin, int end);

}  // namespace sgs
