// Deterministic synthetic polyploid genomes for bench.py and the tests.
//
// Counter-based: every base is a pure function of (seed, chromosome, position) or of a per-chromosome
// splitmix64 stream, so any rank can generate any chromosome independently and the CPU checker and the GPU
// path always see identical bytes.  Shape of one genome:
//   * S subgenomes x n_chrom chromosomes, named Sub<A|B|D..>_chr<n> (so name.split('_')[0] is the truth
//     label the reference's final_stats uses, polycracker.py:1708);
//   * homoeologous chromosomes share an iid-uniform ancestral background and differ by `divergence`
//     substitutions per base;
//   * a repeat library of shared + subgenome-specific families (log-normal copy numbers) pasted as mutated
//     copies, either strand, until ~repeat_fraction of the bases are covered;
//   * N runs (n_fraction of bases) and lower-case soft-masked blocks (softmask_fraction).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../include/polycracker_b200.h"

namespace {

inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline uint64_t h2(uint64_t a, uint64_t b) { return sm64(a ^ sm64(b)); }
inline double u01(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { s += 0x9E3779B97F4A7C15ull; return sm64(s); }
    double uni() { return u01(next()); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
};

const uint64_t TAG_ANC = 0xA11CE5EEDull, TAG_DIV = 0xD1CE0FF5ull, TAG_FAM = 0xFA3117ull, TAG_REP = 0x4E9EA7ull,
               TAG_LEN = 0x1E461ull;
const char LETTERS[] = "ABDEFGHIJKLM";
const char BASE[] = "ACGT";

int64_t chrom_len(const pc_synth_params* p, int64_t c) {
    int64_t total = (int64_t)p->n_subgenomes * p->n_chrom;
    int64_t base = std::max<int64_t>(1, p->genome_size / total);
    int64_t trim = (int64_t)(h2(p->seed ^ TAG_LEN, (uint64_t)c) % (uint64_t)std::max<int64_t>(1, base / 64));
    return std::max<int64_t>(1, base - trim);
}

struct Family { int len; double weight; int sub; std::vector<uint8_t> cons; };

void build_families(const pc_synth_params* p, std::vector<Family>& fams) {
    int F = p->n_shared_families + p->n_specific_families;
    fams.resize(F);
    for (int f = 0; f < F; ++f) {
        uint64_t fs = h2(p->seed ^ TAG_FAM, (uint64_t)f);
        int span = std::max(1, p->max_family_len - p->min_family_len + 1);
        fams[f].len = p->min_family_len + (int)(sm64(fs ^ 1) % (uint64_t)span);
        double u1 = std::max(1e-12, u01(sm64(fs ^ 2))), u2 = u01(sm64(fs ^ 3));
        double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        fams[f].weight = exp(1.5 * z);
        fams[f].sub = f < p->n_shared_families ? -1 : (f - p->n_shared_families) % p->n_subgenomes;
        fams[f].cons.resize(fams[f].len);
        for (int j = 0; j < fams[f].len; ++j) fams[f].cons[j] = (uint8_t)(h2(fs ^ 4, (uint64_t)j) & 3u);
    }
}

}  // namespace

extern "C" {

int64_t pc_synth_n_chrom(const pc_synth_params* p) {
    if (!p || p->n_subgenomes < 1 || p->n_subgenomes > 12 || p->n_chrom < 1) return -1;
    return (int64_t)p->n_subgenomes * p->n_chrom;
}

int pc_synth_chrom_len(const pc_synth_params* p, int64_t chrom, int64_t* len, char name[64]) {
    int64_t total = pc_synth_n_chrom(p);
    if (total < 0 || chrom < 0 || chrom >= total) return PC_EINVAL;
    if (len) *len = chrom_len(p, chrom);
    if (name) snprintf(name, 64, "Sub%c_chr%d", LETTERS[chrom / p->n_chrom], (int)(chrom % p->n_chrom) + 1);
    return PC_OK;
}

int pc_synth_chrom(const pc_synth_params* p, int64_t chrom, uint8_t* out) {
    int64_t total = pc_synth_n_chrom(p);
    if (total < 0 || chrom < 0 || chrom >= total || !out) return PC_EINVAL;
    const int s = (int)(chrom / p->n_chrom), n = (int)(chrom % p->n_chrom);
    const int64_t L = chrom_len(p, chrom);
    const uint64_t anc_seed = h2(p->seed ^ TAG_ANC, (uint64_t)n);
    const uint64_t div_seed = h2(p->seed ^ TAG_DIV, (uint64_t)chrom);
    const uint64_t div_thr = (uint64_t)(std::min(1.0, std::max(0.0, p->divergence)) * 9007199254740992.0);

    // 1. ancestral background + homoeolog divergence (codes 0..3 for now)
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < L; ++i) {
        uint64_t a = h2(anc_seed, (uint64_t)i);
        uint32_t b = (uint32_t)(a & 3u);
        uint64_t d = h2(div_seed, (uint64_t)i);
        if ((d >> 11) < div_thr) b = (b + 1u + (uint32_t)((d & 0x7FFu) % 3u)) & 3u;
        out[i] = (uint8_t)b;
    }

    // 2. repeats
    std::vector<Family> fams;
    build_families(p, fams);
    std::vector<int> elig; std::vector<double> cum;
    double tw = 0;
    for (int f = 0; f < (int)fams.size(); ++f)
        if (fams[f].sub < 0 || fams[f].sub == s) { elig.push_back(f); tw += fams[f].weight; cum.push_back(tw); }
    Rng rng(h2(p->seed ^ TAG_REP, (uint64_t)chrom));
    double frac = std::min(0.999, std::max(0.0, p->repeat_fraction));
    int64_t target = (int64_t)(-log(1.0 - frac) * (double)L);
    int64_t pasted = 0;
    while (!elig.empty() && pasted < target) {
        double x = rng.uni() * tw;
        int idx = (int)(std::lower_bound(cum.begin(), cum.end(), x) - cum.begin());
        if (idx >= (int)elig.size()) idx = (int)elig.size() - 1;
        const Family& fam = fams[elig[idx]];
        int64_t len = std::min<int64_t>(fam.len, L);
        int64_t pos = (int64_t)rng.below((uint64_t)(L - len + 1));
        double mut = rng.uni() * p->max_copy_mutation;
        uint64_t mut_thr = (uint64_t)(mut * 9007199254740992.0);
        bool rev = rng.next() & 1u;
        uint64_t cs = rng.next();
        for (int64_t j = 0; j < len; ++j) {
            uint32_t b = rev ? 3u - fam.cons[fam.len - 1 - j] : fam.cons[j];
            uint64_t d = h2(cs, (uint64_t)j);
            if ((d >> 11) < mut_thr) b = (b + 1u + (uint32_t)((d & 0x7FFu) % 3u)) & 3u;
            out[pos + j] = (uint8_t)b;
        }
        pasted += len;
    }

    // 3. codes -> letters
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < L; ++i) out[i] = (uint8_t)BASE[out[i] & 3u];

    // 4. N runs
    int64_t n_target = (int64_t)(p->n_fraction * (double)L);
    int64_t placed = 0;
    while (placed < n_target) {
        int span = std::max(1, p->max_n_run - p->min_n_run + 1);
        int64_t len = p->min_n_run + (int64_t)rng.below((uint64_t)span);
        len = std::min<int64_t>(std::min<int64_t>(len, n_target - placed), L);
        if (len <= 0) break;
        int64_t pos = (int64_t)rng.below((uint64_t)(L - len + 1));
        memset(out + pos, 'N', (size_t)len);
        placed += len;
    }

    // 5. soft-masked (lower-case) blocks
    int64_t s_target = (int64_t)(p->softmask_fraction * (double)L);
    placed = 0;
    while (placed < s_target) {
        int64_t len = 200 + (int64_t)rng.below(4801);
        len = std::min<int64_t>(std::min<int64_t>(len, s_target - placed), L);
        if (len <= 0) break;
        int64_t pos = (int64_t)rng.below((uint64_t)(L - len + 1));
        for (int64_t j = 0; j < len; ++j) out[pos + j] |= 0x20;
        placed += len;
    }
    return PC_OK;
}

}  // extern "C"
