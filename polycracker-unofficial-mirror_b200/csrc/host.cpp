// Host-side helpers of the C ABI that need no GPU: FASTA reader, k-mer text codecs and the text artefact
// writers of the reference's file formats (.kcount: polycracker.py:199/208, .kcount.fa: polycracker.py:231).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/polycracker_b200.h"

namespace {
thread_local std::string h_err;
}
extern "C" const char* pc_last_error(void);

static const char* BASES = "ACGT";

extern "C" {

int pc_kmer_decode(const uint64_t* keys, int64_t n, int k, char* out) {
    if (!keys || !out || k < 1 || k > 32 || n < 0) return PC_EINVAL;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t x = keys[i];
        char* o = out + i * k;
        for (int j = 0; j < k; ++j) o[j] = BASES[(x >> (2 * (k - 1 - j))) & 3u];
    }
    return PC_OK;
}

int pc_kmer_encode(const char* text, int64_t n, int k, uint64_t* keys) {
    if (!text || !keys || k < 1 || k > 32 || n < 0) return PC_EINVAL;
    for (int64_t i = 0; i < n; ++i) {
        const char* t = text + i * k;
        uint64_t x = 0; bool ok = true;
        for (int j = 0; j < k; ++j) {
            unsigned c = (unsigned char)t[j] & 0xDFu;
            unsigned code;
            switch (c) { case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break; default: code = 0; ok = false; }
            x = (x << 2) | code;
        }
        keys[i] = ok ? x : ~0ull;
    }
    return PC_OK;
}

// ---- FASTA ---------------------------------------------------------------------------------------------
static int fasta_pass(const char* path, int64_t* n_seqs, int64_t* n_bases, int64_t* name_bytes, uint8_t* seq,
                      int64_t* offsets, char* names) {
    FILE* f = fopen(path, "rb");
    if (!f) return PC_EIO;
    std::vector<char> buf(1 << 22);
    int64_t ns = 0, nb = 0, nn = 0;
    bool in_header = false, header_token_done = false, at_line_start = true;
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < got; ++i) {
            char ch = buf[i];
            if (in_header) {
                if (ch == '\n') { in_header = false; at_line_start = true; if (names) names[nn] = '\n'; ++nn; }
                else if (!header_token_done) {
                    if (ch == ' ' || ch == '\t' || ch == '\r') header_token_done = true;
                    else { if (names) names[nn] = ch; ++nn; }
                }
                continue;
            }
            if (ch == '\n') { at_line_start = true; continue; }
            if (ch == '\r') continue;
            if (at_line_start && ch == '>') {
                in_header = true; header_token_done = false;
                if (offsets) offsets[ns] = nb;
                ++ns;
                continue;
            }
            at_line_start = false;
            if (ns == 0) continue;                 // junk before the first header
            if (seq) seq[nb] = (uint8_t)ch;
            ++nb;
        }
    }
    fclose(f);
    if (offsets) offsets[ns] = nb;
    if (n_seqs) *n_seqs = ns;
    if (n_bases) *n_bases = nb;
    if (name_bytes) *name_bytes = nn;
    return PC_OK;
}

int pc_fasta_scan(const char* path, int64_t* n_seqs, int64_t* n_bases, int64_t* name_bytes) {
    if (!path) return PC_EINVAL;
    return fasta_pass(path, n_seqs, n_bases, name_bytes, nullptr, nullptr, nullptr);
}

int pc_fasta_read(const char* path, uint8_t* seq, int64_t* offsets, char* names) {
    if (!path || !seq || !offsets || !names) return PC_EINVAL;
    return fasta_pass(path, nullptr, nullptr, nullptr, seq, offsets, names);
}

// ---- writers -------------------------------------------------------------------------------------------
int pc_write_kcount(const char* path, const uint64_t* keys, const uint32_t* counts, int64_t n, int k, int append) {
    if (!path || (n > 0 && (!keys || !counts)) || k < 1 || k > 32) return PC_EINVAL;
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) return PC_EIO;
    std::vector<char> buf; buf.reserve(1 << 22);
    char line[64];
    for (int64_t i = 0; i < n; ++i) {
        uint64_t x = keys[i];
        for (int j = 0; j < k; ++j) line[j] = BASES[(x >> (2 * (k - 1 - j))) & 3u];
        int len = k;
        line[len++] = '\t';
        len += snprintf(line + len, sizeof line - len, "%u\n", counts[i]);
        buf.insert(buf.end(), line, line + len);
        if (buf.size() > (1u << 22) - 128) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    int bad = ferror(f);
    fclose(f);
    return bad ? PC_EIO : PC_OK;
}

int pc_write_kmer_fasta(const char* path, const uint64_t* keys, int64_t n, int k, int rc_labels) {
    if (!path || (n > 0 && !keys) || k < 1 || k > 32) return PC_EINVAL;
    FILE* f = fopen(path, "wb");
    if (!f) return PC_EIO;
    std::vector<char> buf; buf.reserve(1 << 22);
    char km[33];
    for (int64_t i = 0; i < n; ++i) {
        uint64_t x = keys[i];
        if (rc_labels) {                       // BBTools-style orientation: print the reverse complement (= max)
            for (int j = 0; j < k; ++j) km[j] = BASES[3u - ((x >> (2 * j)) & 3u)];
        } else {
            for (int j = 0; j < k; ++j) km[j] = BASES[(x >> (2 * (k - 1 - j))) & 3u];
        }
        buf.push_back('>'); buf.insert(buf.end(), km, km + k); buf.push_back('\n');
        buf.insert(buf.end(), km, km + k); buf.push_back('\n');
        if (buf.size() > (1u << 22) - 128) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    int bad = ferror(f);
    fclose(f);
    return bad ? PC_EIO : PC_OK;
}


// Text artefacts of the mapping stages, written straight from the CSR the GPU produced (row = chunk/record,
// column = repeat k-mer, value = number of exact occurrences):
//   mode 0  headerless SAM, one record per occurrence (bbmap.sh outm=..., polycracker.py:252).  Only the fields the
//           reference reads are meaningful: QNAME = k-mer, RNAME = chunk (polycracker.py:271-272); POS is 0.
//   mode 1  blasted.bed: "chunk\t0\tlen\tkmer" per occurrence (polycracker.py:272)
//   mode 2  blasted_merged.bed: one line per chunk with >= 1 hit, names comma-joined with multiplicity
//           (`sort().merge(c=4,o='collapse')`, polycracker.py:280)
int pc_write_hit_lines(const char* path, int mode, const int64_t* indptr, const int32_t* indices, const float* data,
                       int64_t n_rows, const uint64_t* keys, int k, const char* names_blob, const int64_t* name_off,
                       const int64_t* row_len, int append) {
    if (!path || !indptr || !names_blob || !name_off || k < 1 || k > 32 || mode < 0 || mode > 2) return PC_EINVAL;
    if (indptr[n_rows] > 0 && (!indices || !data || !keys)) return PC_EINVAL;
    if (mode != 0 && !row_len) return PC_EINVAL;
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) return PC_EIO;
    std::string buf; buf.reserve(1 << 22);
    char km[33]; km[k] = 0;
    char num[32];
    for (int64_t r = 0; r < n_rows; ++r) {
        int64_t b = indptr[r], e = indptr[r + 1];
        if (b == e) continue;
        const char* name = names_blob + name_off[r];
        size_t nlen = (size_t)(name_off[r + 1] - name_off[r]);
        int numlen = row_len ? snprintf(num, sizeof num, "%lld", (long long)row_len[r]) : 0;
        if (mode == 2) { buf.append(name, nlen); buf.append("\t0\t"); buf.append(num, numlen); buf.push_back('\t'); }
        bool first = true;
        for (int64_t i = b; i < e; ++i) {
            uint64_t x = keys[indices[i]];
            for (int j = 0; j < k; ++j) km[j] = BASES[(x >> (2 * (k - 1 - j))) & 3u];
            int64_t cnt = (int64_t)data[i];
            for (int64_t c = 0; c < cnt; ++c) {
                if (mode == 0) {
                    buf.append(km, k); buf.append("\t0\t"); buf.append(name, nlen); buf.append("\t0\t255\t");
                    int l = snprintf(num, sizeof num, "%dM", k); buf.append(num, l);
                    buf.append("\t*\t0\t0\t"); buf.append(km, k); buf.append("\t*\n");
                } else if (mode == 1) {
                    buf.append(name, nlen); buf.append("\t0\t"); buf.append(num, numlen); buf.push_back('\t');
                    buf.append(km, k); buf.push_back('\n');
                } else {
                    if (!first) buf.push_back(',');
                    buf.append(km, k); first = false;
                }
                if (buf.size() > (1u << 22) - 4096) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
            }
        }
        if (mode == 2) buf.push_back('\n');
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    int bad = ferror(f);
    fclose(f);
    return bad ? PC_EIO : PC_OK;
}

}  // extern "C"
