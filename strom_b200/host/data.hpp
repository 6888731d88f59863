// data.hpp -- alignment -> compressed site patterns (mirrors reference src/data.hpp), plus what the
// B200 engine adds: 1-byte tip codes and contiguous pattern slices for sharding across GPUs.
//   * state codes: A,C,G,T -> 0..3; gap / missing (NCL's negative codes) -> 4   (src/data.hpp:235-238)
//   * compressPatterns: unique site columns in lexicographic order + counts     (src/data.hpp:102-161)
// NCL is not available here, so the NEXUS reader below is a small stand-in for the subset the
// reference's fixtures use (taxa/characters or data block, non-interleaved matrix, gap/missing/equate,
// bracket comments).  IUPAC ambiguity symbols get distinct codes >= 5; the engine, like BeagleLib,
// treats every code >= 4 as missing.
#pragma once
#include <algorithm>
#include <cctype>
#include <fstream>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include "xstrom.hpp"
#ifdef STROM_B200_ENGINE
#include "../../include/strom_b200.h"
#endif

namespace strom {

class Data {
  public:
    typedef std::vector<double> pattern_counts_t;
    typedef std::vector<std::string> taxon_names_t;
    typedef std::vector<int> pattern_t;
    typedef std::vector<pattern_t> data_matrix_t;
    typedef std::shared_ptr<Data> SharedPtr;

    void getDataFromFile(const std::string& filename);
    // same entry point for data that is already in memory: matrix[taxon][site] of state codes
    void setDataMatrix(const taxon_names_t& names, const data_matrix_t& matrix, bool compress = true);

    const pattern_counts_t& getPatternCounts() const { return _pattern_counts; }
    const taxon_names_t& getTaxonNames() const { return _taxon_names; }
    const data_matrix_t& getDataMatrix() const { return _data_matrix; }

    void clear() {
        _pattern_counts.clear();
        _taxon_names.clear();
        _data_matrix.clear();
    }
    unsigned getNumPatterns() const { return static_cast<unsigned>(_pattern_counts.size()); }
    unsigned getNumTaxa() const { return static_cast<unsigned>(_taxon_names.size()); }
    unsigned getSeqLen() const {
        return static_cast<unsigned>(std::accumulate(_pattern_counts.begin(), _pattern_counts.end(), 0.0));
    }

    // ---- engine-side packing -------------------------------------------------------------
    // Pattern slice of rank `rank` out of `nranks`: [floor(rank*P/nranks), floor((rank+1)*P/nranks))
    static void sliceBounds(unsigned npatterns, unsigned rank, unsigned nranks, unsigned& lo, unsigned& hi) {
        lo = static_cast<unsigned>(static_cast<unsigned long long>(npatterns) * rank / nranks);
        hi = static_cast<unsigned>(static_cast<unsigned long long>(npatterns) * (rank + 1) / nranks);
    }
    // 1-byte codes (0..3, 4 = missing) of taxon `t` for patterns [lo, hi)
    std::vector<unsigned char> packTipCodes(unsigned t, unsigned lo, unsigned hi) const {
        std::vector<unsigned char> out(hi - lo);
        const pattern_t& row = _data_matrix[t];
        for (unsigned p = lo; p < hi; ++p) out[p - lo] = (row[p] >= 0 && row[p] < 4) ? static_cast<unsigned char>(row[p]) : 4;
        return out;
    }

    // ---- pattern compression on the GPU (engine builds) ------------------
    // Alignments of at least `min_sites` sites are compressed by sb200CompressPatterns on CUDA device `device`
    // (same patterns, same order, same counts); shorter ones, and any alignment with a state code >= 32, take
    // the host path.  Default: 32768 sites, device 0.
    void setDeviceCompression(unsigned min_sites, int device = 0) {
        _device_min_sites = min_sites;
        _device = device;
    }
    bool compressedOnDevice() const { return _compressed_on_device; }

  private:
    void compressPatterns();
    bool compressPatternsOnDevice();
    unsigned _device_min_sites = 32768;
    int _device = 0;
    bool _compressed_on_device = false;
    static std::string stripComments(const std::string& s);
    static std::string lower(std::string s) {
        for (char& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
        return s;
    }

    pattern_counts_t _pattern_counts;
    taxon_names_t _taxon_names;
    data_matrix_t _data_matrix;
};

inline std::string Data::stripComments(const std::string& s) {
    std::string out;
    int depth = 0;
    for (char ch : s) {
        if (ch == '[') ++depth;
        else if (ch == ']') { if (depth > 0) --depth; }
        else if (depth == 0) out += ch;
    }
    return out;
}

inline void Data::compressPatterns() {
    if (_data_matrix.empty()) throw XStrom("Attempted to compress an empty data matrix");
    const unsigned ntaxa = static_cast<unsigned>(_data_matrix.size());
    const unsigned seqlen = static_cast<unsigned>(_data_matrix[0].size());
    for (const pattern_t& row : _data_matrix)
        if (row.size() != seqlen) throw XStrom("data matrix rows differ in length");
    _compressed_on_device = compressPatternsOnDevice();
    if (_compressed_on_device) return;
    // sort site indices by their column, lexicographically over taxa: the order std::map<pattern_t,...> yields
    std::vector<unsigned> order(seqlen);
    std::iota(order.begin(), order.end(), 0u);
    auto less = [&](unsigned a, unsigned b) {
        for (unsigned t = 0; t < ntaxa; ++t) {
            const int x = _data_matrix[t][a], y = _data_matrix[t][b];
            if (x != y) return x < y;
        }
        return false;
    };
    std::stable_sort(order.begin(), order.end(), less);
    data_matrix_t packed(ntaxa);
    pattern_counts_t counts;
    for (unsigned i = 0; i < seqlen; ++i) {
        if (i > 0 && !less(order[i - 1], order[i])) {
            counts.back() += 1.0;
            continue;
        }
        counts.push_back(1.0);
        for (unsigned t = 0; t < ntaxa; ++t) packed[t].push_back(_data_matrix[t][order[i]]);
    }
    _data_matrix.swap(packed);
    _pattern_counts.swap(counts);
    const unsigned total = static_cast<unsigned>(std::accumulate(_pattern_counts.begin(), _pattern_counts.end(), 0.0));
    if (total != seqlen)
        throw XStrom("Total number of sites before compaction (" + std::to_string(seqlen) + ") not equal to total number of sites after (" +
                     std::to_string(total) + ")");
}

// The engine's radix sort over site columns; false = not applicable here (host path takes over).
inline bool Data::compressPatternsOnDevice() {
#ifdef STROM_B200_ENGINE
    const unsigned ntaxa = static_cast<unsigned>(_data_matrix.size());
    const unsigned seqlen = static_cast<unsigned>(_data_matrix[0].size());
    if (seqlen < _device_min_sites) return false;
    std::vector<unsigned char> codes(static_cast<size_t>(ntaxa) * seqlen);
    for (unsigned t = 0; t < ntaxa; ++t)
        for (unsigned i = 0; i < seqlen; ++i) {
            const int c = _data_matrix[t][i];
            if (c < 0 || c >= 32) return false;
            codes[static_cast<size_t>(t) * seqlen + i] = static_cast<unsigned char>(c);
        }
    std::vector<unsigned char> packed(codes.size());
    pattern_counts_t counts(seqlen);
    int npatterns = 0;
    const int code = sb200CompressPatterns(_device, static_cast<int>(ntaxa), static_cast<int>(seqlen), codes.data(), packed.data(),
                                           counts.data(), &npatterns);
    if (code != 0) throw XStrom("pattern compression on the GPU failed (engine code " + std::to_string(code) + ")");
    counts.resize(static_cast<size_t>(npatterns));
    for (unsigned t = 0; t < ntaxa; ++t)
        _data_matrix[t].assign(packed.begin() + static_cast<size_t>(t) * npatterns, packed.begin() + static_cast<size_t>(t + 1) * npatterns);
    _pattern_counts.swap(counts);
    return true;
#else
    return false;
#endif
}

inline void Data::setDataMatrix(const taxon_names_t& names, const data_matrix_t& matrix, bool compress) {
    clear();
    _taxon_names = names;
    _data_matrix = matrix;
    if (_taxon_names.size() != _data_matrix.size()) throw XStrom("taxon names and data matrix rows differ in number");
    if (compress) compressPatterns();
    else _pattern_counts.assign(_data_matrix.empty() ? 0 : _data_matrix[0].size(), 1.0);
}

inline void Data::getDataFromFile(const std::string& filename) {
    std::ifstream in(filename.c_str());
    if (!in) throw XStrom("could not open data file " + filename);
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = stripComments(ss.str());
    const std::string low = lower(text);

    char gap = '-', missing = '?';
    std::map<char, char> equate;
    size_t fpos = low.find("format");
    if (fpos != std::string::npos) {
        const size_t fend = low.find(';', fpos);
        const std::string fmt = text.substr(fpos, fend - fpos), flow = low.substr(fpos, fend - fpos);
        auto valueOf = [&](const std::string& key) -> std::string {
            size_t k = flow.find(key);
            if (k == std::string::npos) return "";
            k = flow.find('=', k);
            if (k == std::string::npos) return "";
            ++k;
            while (k < fmt.size() && std::isspace(static_cast<unsigned char>(fmt[k]))) ++k;
            if (k < fmt.size() && fmt[k] == '"') {
                const size_t e = fmt.find('"', k + 1);
                return fmt.substr(k + 1, e - k - 1);
            }
            size_t e = k;
            while (e < fmt.size() && !std::isspace(static_cast<unsigned char>(fmt[e]))) ++e;
            return fmt.substr(k, e - k);
        };
        std::string v = valueOf("gap");
        if (!v.empty()) gap = v[0];
        v = valueOf("missing");
        if (!v.empty()) missing = v[0];
        std::istringstream eq(valueOf("equate"));
        std::string pair;
        while (eq >> pair)
            if (pair.size() >= 3 && pair[1] == '=') equate[static_cast<char>(std::toupper(static_cast<unsigned char>(pair[0])))] = pair[2];
    }
    size_t mpos = low.find("matrix");
    if (mpos == std::string::npos) throw XStrom("no matrix found in " + filename);
    const size_t mend = text.find(';', mpos);
    std::istringstream body(text.substr(mpos + 6, mend - mpos - 6));

    static const std::string ambig = "RYMKSWHBVDNX";
    auto code = [&](char ch) -> int {
        char c = static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
        std::map<char, char>::const_iterator it = equate.find(c);
        if (it != equate.end()) c = static_cast<char>(std::toupper(static_cast<unsigned char>(it->second)));
        switch (c) {
            case 'A': return 0;
            case 'C': return 1;
            case 'G': return 2;
            case 'T': return 3;
            default: break;
        }
        if (c == gap || c == missing) return 4;              // NCL's negative codes -> 4
        const size_t a = ambig.find(c);
        if (a != std::string::npos) return 5 + static_cast<int>(a);
        throw XStrom(std::string("unknown DNA symbol '") + ch + "' in " + filename);
    };

    taxon_names_t names;
    data_matrix_t matrix;
    std::map<std::string, size_t> rowOf;
    std::string line;
    while (std::getline(body, line)) {
        std::istringstream ls(line);
        std::string name, chunk;
        if (!(ls >> name)) continue;
        size_t row;
        std::map<std::string, size_t>::const_iterator it = rowOf.find(name);
        if (it == rowOf.end()) {
            row = names.size();
            rowOf[name] = row;
            names.push_back(name);
            matrix.push_back(pattern_t());
        } else {
            row = it->second;
        }
        while (ls >> chunk)
            for (char ch : chunk) matrix[row].push_back(code(ch));
    }
    setDataMatrix(names, matrix, true);
}

}  // namespace strom
