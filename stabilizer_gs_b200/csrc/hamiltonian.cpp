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
}

void Hamiltonian::finalize()
{
    // add_Pauli buckets by last site (cpp/state.cpp:83-86); paulis_list concatenates the buckets (:88-96)
    order.clear();
    bucket_off.assign(n + 1, 0);
    for (int q = 0; q < n; q++) {
        bucket_off[q] = (int)order.size();
        for (int t = 0; t < T(); t++)
            if (end[t] - 1 == q) order.push_back(t);
    }
    bucket_off[n] = (int)order.size();
    W = (2 * (n + 1) + 63) / 64;
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
    if (H.n + 1 > SGS_MAX_QUBITS) { err = {SGS_ERR_LIMIT, "n exceeds SGS_MAX_QUBITS - 1"}; return false; }
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
            PackedRow r;
            for (size_t i = 0; i < letters.size(); i++) {
                uint64_t v;
                switch (letters[i]) {
                case 'I': v = 0; break;
                case 'Z': v = 1; break;
                case 'X': v = 2; break;
                case 'Y': v = 3; break;
                default:                          // cpp/stab.cpp:138-140 throws invalid_argument
                    err = {SGS_ERR_PARSE, "line " + std::to_string(lineno) + ": invalid character '" + letters[i] + "'"};
                    return false;
                }
                int bit = 2 * (sites[i] + off);
                r.w[bit >> 6] = (r.w[bit >> 6] & ~(3ULL << (bit & 63))) | (v << (bit & 63));
            }
            H.add_term(r, w, qmin + off, qmax + off + 1);
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
