// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h). The text side of load_alignment (reference
// src/haslr_assemble/src/Longread.cpp:250-289) the way the reference does it: getline, split on tabs into strings,
// istringstream conversions (str2type, Common.hpp:126-133), the first optional column that starts with "cg:Z:".
// All rows are returned (the load filters are part of oracle_compact_lr). The CIGAR string is turned into run-length
// operations the way expand_cigar (Common.cpp) reads it: a decimal count followed by one operation character.
#include <cctype>
#include <sstream>
#include <string>
#include <vector>

#include "oracle.h"

namespace {
// str2type<uint32_t> (Common.hpp:126-133) is `istringstream >> n`. Restated without the stream: numeric extraction needs
// the C++ locale facets, which are not initialised when this library is dlopen'ed into a C host process (python) — the
// stream then fails and every column reads 0. What libstdc++'s extraction does (checked against a stand-alone program,
// tests/test_paf_host.py::test_oracle_number_reading lists the cases): skip white space, an optional sign, decimal
// digits up to the first other character; no digits -> 0 here (uninitialised in the reference); a magnitude above
// 2^32 - 1 -> 4294967295; '-' negates modulo 2^32.
template <typename T> T str2type(const std::string& str) {
    size_t i = 0;
    while (i < str.size() && isspace((unsigned char)str[i])) ++i;
    bool neg = false;
    if (i < str.size() && (str[i] == '+' || str[i] == '-')) { neg = str[i] == '-'; ++i; }
    unsigned long long v = 0;
    bool over = false;
    for (; i < str.size() && isdigit((unsigned char)str[i]); ++i) {
        v = v * 10 + (unsigned long long)(str[i] - '0');
        if (v > 0xFFFFFFFFull) over = true, v = 0xFFFFFFFFull;
    }
    if (over) return (T)0xFFFFFFFFull;
    return (T)(neg ? (uint32_t)(0u - (uint32_t)v) : (uint32_t)v);
}
void str_split(const std::string& s, char delim, std::vector<std::string>& out) {
    out.clear();
    std::istringstream in(s);
    std::string tok;
    while (std::getline(in, tok, delim)) out.push_back(tok);
    if (!s.empty() && s.back() == delim) out.push_back("");
}
}  // namespace

extern "C" int64_t oracle_parse_paf(const char* text, uint64_t n_bytes, uint64_t row_cap, uint64_t op_cap,
                                    uint32_t* q_id, uint32_t* q_len, uint32_t* q_start, uint32_t* q_end, uint8_t* is_rev,
                                    uint32_t* t_id, uint32_t* t_len, uint32_t* t_start, uint32_t* t_end, uint32_t* n_match,
                                    uint32_t* n_block, uint8_t* mapq, uint32_t* cg_off, uint32_t* cg_ops, uint64_t* n_ops_out) {
    std::istringstream fin(std::string(text, text + n_bytes));
    std::string line;
    std::vector<std::string> f;
    uint64_t rows = 0, ops = 0;
    cg_off[0] = 0;
    while (std::getline(fin, line)) {
        if (line.empty()) continue;
        str_split(line, '\t', f);
        if (f.size() < 12) return -2;                      // the reference would index past the vector
        if (rows >= row_cap) return -1;
        q_id[rows] = str2type<uint32_t>(f[0]); q_len[rows] = str2type<uint32_t>(f[1]);
        q_start[rows] = str2type<uint32_t>(f[2]); q_end[rows] = str2type<uint32_t>(f[3]);
        is_rev[rows] = (uint8_t)(f[4].size() && f[4][0] == '-' ? 1 : 0);
        t_id[rows] = str2type<uint32_t>(f[5]); t_len[rows] = str2type<uint32_t>(f[6]);
        t_start[rows] = str2type<uint32_t>(f[7]); t_end[rows] = str2type<uint32_t>(f[8]);
        n_match[rows] = str2type<uint32_t>(f[9]); n_block[rows] = str2type<uint32_t>(f[10]);
        mapq[rows] = (uint8_t)str2type<uint32_t>(f[11]);
        std::string cg;
        for (size_t i = 12; i < f.size(); i++)
            if (f[i].substr(0, 5) == "cg:Z:") { cg = f[i].substr(5); break; }
        for (size_t c = 0; c < cg.size();) {
            uint32_t n = 0;
            while (c < cg.size() && isdigit((unsigned char)cg[c])) n = n * 10 + (uint32_t)(cg[c++] - '0');
            if (c >= cg.size()) break;
            const char op = cg[c++];
            if (ops >= op_cap) return -1;
            cg_ops[ops++] = (n << 2) | (op == 'M' ? 0u : op == 'I' ? 1u : 2u);
        }
        cg_off[++rows] = (uint32_t)ops;
    }
    *n_ops_out = ops;
    return (int64_t)rows;
}
