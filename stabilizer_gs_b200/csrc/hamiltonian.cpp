// hamiltonian.cpp — Hamiltonian text parser of libsgs.
// Format (README.md:33-45 of the reference): first line "n k" (or "l k" for periodic files),
// then one term per line: weight, Pauli letters, one site per letter.
#include "hamiltonian.hpp"

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <sstream>

#include "../../include/sgs.h"

namespace sgs {

void Hamiltonian::add_term(const PackedRow &r, double w, int b, int e)
{
    terms.push_back(r);
    weights.push_back(w);
    begin.push_back(b);
    end.push_back(e);
    // compact copy (relative to the window start) when the window is short enough
    uint64_t c = 0;
    const bool fits = e - b <= kCompactSites;
    if (fits)
        for (int j = b; j < e; j++) c |= ((r.w[(2 * j) >> 6] >> ((2 * j) & 63)) & 3) << (2 * (j - b));
    crow.push_back(c);
    compact.push_back(fits ? 1 : 0);
}

// a term given letter by letter; window = [min site, max site + 1) (cpp/stab.cpp:120-122).  In a wide Hamiltonian
// (n + 1 > 128) only the compact row exists, so the window may span at most 32 sites.
bool Hamiltonian::add_term_sites(const std::vector<int> &sites, const std::vector<int> &codes, double w)
{
    const int b = *std::min_element(sites.begin(), sites.end());
    const int e = *std::max_element(sites.begin(), sites.end()) + 1;
    if (!wide) {
        PackedRow r;
        for (size_t i = 0; i < sites.size(); i++) {
            const int bit = 2 * sites[i];
            r.w[bit >> 6] = (r.w[bit >> 6] & ~(3ULL << (bit & 63))) | ((uint64_t)codes[i] << (bit & 63));
        }
        add_term(r, w, b, e);
        return true;
    }
    if (e - b > kCompactSites) return false;
    uint64_t c = 0;
    for (size_t i = 0; i < sites.size(); i++) {
        const int bit = 2 * (sites[i] - b);
        c = (c & ~(3ULL << bit)) | ((uint64_t)codes[i] << bit);
    }
    terms.push_back(PackedRow());
    weights.push_back(w);
    begin.push_back(b);
    end.push_back(e);
    crow.push_back(c);
    compact.push_back(1);
    return true;
}

void Hamiltonian::row_words(int t, uint64_t *out, int Wout) const
{
    for (int i = 0; i < Wout; i++) out[i] = 0;
    if (!wide) { for (int i = 0; i < Wout && i < kMaxWords; i++) out[i] = terms[t].w[i]; return; }
    for (int j = 0; j < end[t] - begin[t]; j++) {
        const uint64_t v = (crow[t] >> (2 * j)) & 3;
        const int bit = 2 * (begin[t] + j);
        if (v && (bit >> 6) < Wout) out[bit >> 6] |= v << (bit & 63);
    }
}

void Hamiltonian::finalize()
{
    // add_Pauli buckets by last site (cpp/state.cpp:83-86); paulis_list concatenates the buckets (:88-96)
    order.clear();
    bucket_off.assign(n + 1, 0);
    std::vector<int> cnt(n + 1, 0);
    for (int t = 0; t < T(); t++) cnt[end[t] - 1]++;
    for (int q = 0, o = 0; q <= n; q++) { bucket_off[q] = o; if (q < n) o += cnt[q]; }
    order.assign(T(), 0);
    std::vector<int> cur(bucket_off.begin(), bucket_off.end());
    for (int t = 0; t < T(); t++) order[cur[end[t] - 1]++] = t;            // file order inside a bucket
    W = (2 * n + 63) / 64;
    if (W < 1) W = 1;
}

static bool is_blank(const std::string &s)
{
    return std::all_of(s.begin(), s.end(), [](unsigned char c) { return std::isspace(c); });
}

bool parse_hamiltonian_text(const std::string &text, bool periodic, int cmax, Hamiltonian &H, ParseError &err)
{
    std::istringstream in(text);
    std::string line;
    if (!std::getline(in, line)) { err = {SGS_ERR_PARSE, "empty Hamiltonian file"}; return false; }
    {
        std::istringstream hs(line);
        int a = 0, b = 0;
        if (!(hs >> a >> b)) { err = {SGS_ERR_PARSE, "first line must be 'n k'"}; return false; }
        H = Hamiltonian();
        H.k = b;
        if (periodic) { H.l = a; H.n = a * (1 + cmax) + 2 * b; }   // py/state.py:55-56
        else { H.n = a; }
    }
    if (H.n <= 0 || H.k <= 0) { err = {SGS_ERR_PARSE, "n and k must be positive"}; return false; }
    // n + 1 > 128: a "wide" Hamiltonian — only the k-local / periodic DP can search it (its objects live on windows
    // of <= 2k sites whatever n is); the sparse search refuses it.  Every term must then span at most 32 sites.
    H.wide = H.n + 1 > SGS_MAX_QUBITS;
    if (H.n > (1 << 20)) { err = {SGS_ERR_LIMIT, "n exceeds 2^20"}; return false; }
    int lineno = 1;
    while (std::getline(in, line)) {
        lineno++;
        if (is_blank(line)) break;   // the Python reference stops at the first blank line (py/state.py:41-43)
        std::istringstream ls(line);
        double w;
        std::string letters;
        if (!(ls >> w >> letters)) { err = {SGS_ERR_PARSE, "line " + std::to_string(lineno) + ": expected 'weight letters sites...'"}; return false; }
        std::vector<int> sites;
        int s;
        while (ls >> s) sites.push_back(s);
        if (sites.size() != letters.size() || sites.empty()) {
            err = {SGS_ERR_PARSE, "line " + std::to_string(lineno) + ": one site per letter required"};
            return false;
        }
        int qmin = *std::min_element(sites.begin(), sites.end());
        int qmax = *std::max_element(sites.begin(), sites.end());
        int lo = periodic ? -H.n : 0, hi = periodic ? H.n : 0;
        for (int sh = lo; sh <= hi; sh++) {
            int off = periodic ? sh * H.l : 0;
            if (qmax + off >= H.n || qmin + off < 0) {
                if (periodic) continue;          // py/state.py:70: keep only the shifts that fit
                err = {SGS_ERR_PARSE, "line " + std::to_string(lineno) + ": site out of range"};
                return false;
            }
            std::vector<int> codes(letters.size()), at(letters.size());
            for (size_t i = 0; i < letters.size(); i++) {
                switch (letters[i]) {
                case 'I': codes[i] = 0; break;
                case 'Z': codes[i] = 1; break;
                case 'X': codes[i] = 2; break;
                case 'Y': codes[i] = 3; break;
                default:                          // cpp/stab.cpp:138-140 throws invalid_argument
                    err = {SGS_ERR_PARSE, "line " + std::to_string(lineno) + ": invalid character '" + letters[i] + "'"};
                    return false;
                }
                at[i] = sites[i] + off;
            }
            if (!H.add_term_sites(at, codes, w)) {
                err = {SGS_ERR_LIMIT, "line " + std::to_string(lineno) + ": with n > 127 a term may span at most 32 sites"};
                return false;
            }
        }
    }
    H.finalize();
    return true;
}

bool load_hamiltonian_file(const std::string &path, bool periodic, int cmax, Hamiltonian &H, ParseError &err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = {SGS_ERR_IO, "cannot open '" + path + "'"}; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    return parse_hamiltonian_text(ss.str(), periodic, cmax, H, err);
}

std::string row_to_string(const uint64_t *row, int sign, int begin, int end)
{
    std::string s(1, sign >= 0 ? '+' : '-');
    for (int j = begin; j < end; j++) {
        int bit = 2 * j;
        int v = (int)((row[bit >> 6] >> (bit & 63)) & 3);   // bit0 = z, bit1 = x
        s += "IZXY"[v];
    }
    return s;
}

}  // namespace sgs
