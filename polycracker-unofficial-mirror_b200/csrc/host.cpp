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

}  // extern "C"
