// unc_index_build.cpp -- bwa-compatible FM-index construction (host C++).
//
// Replaces BwaIndex::create -> bwa_idx_build (reference src/bwa_index.hpp:92-101,
// submods/bwa/bwtindex.c:255-323) for the `uncalled index` command.  Writes the same five
// files, byte for byte, that `bwa index` writes:
//   <p>.pac  forward-only 2-bit packed sequence          (bntseq.c:306-322)
//   <p>.ann  <p>.amb  sequence / ambiguity tables          (bntseq.c:66-95)
//   <p>.bwt  primary, L2[1..4], Occ-interleaved BWT of fwd+revcomp   (bwtindex.c:64-130,
//            bwt_bwtupdate_core :132-163, bwt.c:381-393)
//   <p>.sa   SA sampled every 32 rows                      (bwt.c:61-84,395-405)
// The BWT is a mathematical object, so it is built here with a linear-time SA-IS suffix
// sorter written for this project rather than with bwa's is.c / bwt_gen.c.
// Ambiguous bases become lrand48()&3 after srand48(11) exactly as bwa does (bntseq.c:266,296).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/unc_b200.h"

namespace {

// ---- SA-IS (induced sorting).  s[n-1] must be a unique smallest sentinel (value 0).
template <typename T>
void sais(const T *s, int32_t *SA, int32_t n, int32_t K) {
    std::vector<bool> t((size_t) n);  // true = S-type
    t[n - 1] = true;
    for (int32_t i = n - 2; i >= 0; i--) t[i] = s[i] < s[i + 1] || (s[i] == s[i + 1] && t[i + 1]);
    auto is_lms = [&](int32_t i) { return i > 0 && t[i] && !t[i - 1]; };
    std::vector<int32_t> bkt((size_t) K + 1);
    auto buckets = [&](bool end) {
        std::fill(bkt.begin(), bkt.end(), 0);
        for (int32_t i = 0; i < n; i++) bkt[s[i]]++;
        int32_t sum = 0;
        for (int32_t c = 0; c <= K; c++) {
            sum += bkt[c];
            bkt[c] = end ? sum : sum - bkt[c];
        }
    };
    auto induce = [&]() {
        buckets(false);
        for (int32_t i = 0; i < n; i++) {
            int32_t j = SA[i] - 1;
            if (SA[i] > 0 && !t[j]) SA[bkt[s[j]]++] = j;
        }
        buckets(true);
        for (int32_t i = n - 1; i >= 0; i--) {
            int32_t j = SA[i] - 1;
            if (SA[i] > 0 && t[j]) SA[--bkt[s[j]]] = j;
        }
    };
    // stage 1: sort the LMS substrings
    buckets(true);
    for (int32_t i = 0; i < n; i++) SA[i] = -1;
    for (int32_t i = 1; i < n; i++)
        if (is_lms(i)) SA[--bkt[s[i]]] = i;
    induce();
    int32_t n1 = 0;
    for (int32_t i = 0; i < n; i++)
        if (is_lms(SA[i])) SA[n1++] = SA[i];
    for (int32_t i = n1; i < n; i++) SA[i] = -1;
    int32_t name = 0, prev = -1;
    for (int32_t i = 0; i < n1; i++) {
        int32_t pos = SA[i];
        bool diff = false;
        for (int32_t d = 0;; d++) {
            if (prev == -1 || s[pos + d] != s[prev + d] || t[pos + d] != t[prev + d]) { diff = true; break; }
            if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
        }
        if (diff) { name++; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (int32_t i = n - 1, j = n - 1; i >= n1; i--)
        if (SA[i] >= 0) SA[j--] = SA[i];
    // stage 2: solve the reduced problem
    int32_t *SA1 = SA, *s1 = SA + n - n1;
    if (name < n1) sais<int32_t>(s1, SA1, n1, name - 1);
    else for (int32_t i = 0; i < n1; i++) SA1[s1[i]] = i;
    // stage 3: induce the final order
    buckets(true);
    for (int32_t i = 1, j = 0; i < n; i++)
        if (is_lms(i)) s1[j++] = i;
    for (int32_t i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
    for (int32_t i = n1; i < n; i++) SA[i] = -1;
    for (int32_t i = n1 - 1; i >= 0; i--) {
        int32_t j = SA[i];
        SA[i] = -1;
        SA[--bkt[s[j]]] = j;
    }
    induce();
}

struct Ann { std::string name, anno; int64_t offset; int32_t len, n_ambs; };
struct Amb { int64_t offset; int32_t len; char amb; };

int nt4(int c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

bool write_file(const std::string &fn, const void *p, size_t n) {
    FILE *fp = fopen(fn.c_str(), "wb");
    if (!fp) return false;
    bool ok = n == 0 || fwrite(p, 1, n, fp) == n;
    return fclose(fp) == 0 && ok;
}

}  // namespace

extern "C" int unc_index_build(const char *fasta_path, const char *prefix_c) {
    if (!fasta_path || !prefix_c) return UNC_E_ARG;
    const std::string prefix = prefix_c;
    FILE *fp = fopen(fasta_path, "rb");
    if (!fp) return UNC_E_IO;
    // ---- parse FASTA (kseq semantics: name = first word of the header, comment = the rest)
    std::vector<Ann> anns;
    std::vector<Amb> ambs;
    std::vector<uint8_t> fwd;  // one base per byte
    srand48(11);
    {
        std::vector<char> buf;
        fseek(fp, 0, SEEK_END);
        long sz = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        buf.resize((size_t) sz);
        if (sz > 0 && fread(buf.data(), 1, (size_t) sz, fp) != (size_t) sz) { fclose(fp); return UNC_E_IO; }
        fclose(fp);
        size_t i = 0, n = buf.size();
        while (i < n) {
            while (i < n && buf[i] != '>') {  // skip to the next header
                while (i < n && buf[i] != '\n') i++;
                if (i < n) i++;
            }
            if (i >= n) break;
            i++;  // '>'
            size_t ls = i;
            while (i < n && buf[i] != '\n') i++;
            std::string header(buf.data() + ls, buf.data() + i);
            if (!header.empty() && header.back() == '\r') header.pop_back();
            if (i < n) i++;
            Ann a;
            size_t sp = header.find_first_of(" \t");
            a.name = header.substr(0, sp);
            a.anno = "(null)";
            if (sp != std::string::npos && sp + 1 < header.size()) a.anno = header.substr(sp + 1);
            a.offset = (int64_t) fwd.size();
            a.n_ambs = 0;
            int lasts = 0;
            int64_t len = 0;
            while (i < n && buf[i] != '>') {
                char ch = buf[i++];
                if (ch == '\n' || ch == '\r' || ch == ' ' || ch == '\t') continue;
                int c = nt4(ch);
                if (c >= 4) {
                    if (lasts == ch) {
                        ambs.back().len++;
                    } else {
                        Amb h = {a.offset + len, 1, ch};
                        ambs.push_back(h);
                        a.n_ambs++;
                    }
                    c = (int) (lrand48() & 3);
                }
                lasts = ch;
                fwd.push_back((uint8_t) c);
                len++;
            }
            a.len = (int32_t) len;
            anns.push_back(a);
        }
    }
    const int64_t l_pac = (int64_t) fwd.size();
    if (l_pac == 0) return UNC_E_IO;
    if (2 * l_pac + 1 >= 0x7FFFFFF0ll) return UNC_E_TOO_LARGE;

    // ---- .pac (forward only), .ann, .amb
    {
        std::vector<uint8_t> pac((size_t) (l_pac >> 2) + ((l_pac & 3) == 0 ? 0 : 1), 0);
        for (int64_t l = 0; l < l_pac; l++) pac[(size_t) (l >> 2)] |= (uint8_t) (fwd[(size_t) l] << ((~l & 3) << 1));
        if ((l_pac % 4) == 0) pac.push_back(0);
        pac.push_back((uint8_t) (l_pac % 4));
        if (!write_file(prefix + ".pac", pac.data(), pac.size())) return UNC_E_IO;
        FILE *fa = fopen((prefix + ".ann").c_str(), "w");
        if (!fa) return UNC_E_IO;
        fprintf(fa, "%lld %d %u\n", (long long) l_pac, (int) anns.size(), 11u);
        for (const Ann &a : anns) {
            fprintf(fa, "%d %s", 0, a.name.c_str());
            if (!a.anno.empty()) fprintf(fa, " %s\n", a.anno.c_str());
            else fprintf(fa, "\n");
            fprintf(fa, "%lld %d %d\n", (long long) a.offset, a.len, a.n_ambs);
        }
        fclose(fa);
        fa = fopen((prefix + ".amb").c_str(), "w");
        if (!fa) return UNC_E_IO;
        fprintf(fa, "%lld %d %u\n", (long long) l_pac, (int) anns.size(), (unsigned) ambs.size());
        for (const Amb &h : ambs) fprintf(fa, "%lld %d %c\n", (long long) h.offset, h.len, h.amb);
        fclose(fa);
    }

    // ---- text = forward + reverse complement, suffix array, BWT
    const int64_t n = 2 * l_pac;  // bwt->seq_len
    std::vector<uint8_t> text((size_t) n + 1);
    for (int64_t i = 0; i < l_pac; i++) text[(size_t) i] = (uint8_t) (fwd[(size_t) i] + 1);
    for (int64_t i = 0; i < l_pac; i++) text[(size_t) (l_pac + i)] = (uint8_t) (3 - fwd[(size_t) (l_pac - 1 - i)] + 1);
    text[(size_t) n] = 0;  // sentinel
    std::vector<int32_t> SA((size_t) n + 1);
    sais<uint8_t>(text.data(), SA.data(), (int32_t) (n + 1), 4);

    uint64_t L2[5] = {0, 0, 0, 0, 0};
    for (int64_t i = 0; i < n; i++) L2[text[(size_t) i]]++;  // text value c+1 -> L2[c+1]
    for (int i = 2; i <= 4; i++) L2[i] += L2[i - 1];
    // BWT over rows 0..n (row 0 = sentinel suffix); '$' (row `primary`) is dropped
    uint64_t primary = 0;
    std::vector<uint8_t> bw((size_t) n);
    {
        size_t k = 0;
        for (int64_t i = 0; i <= n; i++) {
            if (SA[(size_t) i] == 0) primary = (uint64_t) i;
            else bw[k++] = (uint8_t) (text[(size_t) SA[(size_t) i] - 1] - 1);
        }
    }
    // Occ interleave: 4 x u64 cumulative counts before every 128 symbols, plus the final totals
    const uint64_t n_occ = (uint64_t) ((n + 127) / 128 + 1);
    const uint64_t bwt_words = (uint64_t) ((n + 15) >> 4) + n_occ * 8;
    std::vector<uint32_t> out((size_t) bwt_words, 0u);
    {
        uint64_t c[4] = {0, 0, 0, 0};
        size_t k = 0;
        for (int64_t i = 0; i < n; i++) {
            if ((i & 127) == 0) { memcpy(&out[k], c, 32); k += 8; }
            if ((i & 15) == 0) k++;
            out[k - 1] |= (uint32_t) bw[(size_t) i] << ((15 - (i & 15)) << 1);
            c[bw[(size_t) i]]++;
        }
        memcpy(&out[k], c, 32);
        if (k + 8 != bwt_words) return UNC_E_IO;
    }
    {
        FILE *fb = fopen((prefix + ".bwt").c_str(), "wb");
        if (!fb) return UNC_E_IO;
        fwrite(&primary, 8, 1, fb);
        fwrite(&L2[1], 8, 4, fb);
        fwrite(out.data(), 4, out.size(), fb);
        if (fclose(fb) != 0) return UNC_E_IO;
    }
    // ---- sampled SA: rows 32, 64, ... (row 0 is written as -1 on load and not stored)
    {
        const uint64_t sa_intv = 32, seq_len = (uint64_t) n;
        const uint64_t n_sa = (seq_len + sa_intv) / sa_intv;
        std::vector<uint64_t> sa((size_t) n_sa, 0);
        for (uint64_t j = 1; j < n_sa; j++) sa[(size_t) j] = (uint64_t) SA[(size_t) (j * sa_intv)];
        FILE *fs = fopen((prefix + ".sa").c_str(), "wb");
        if (!fs) return UNC_E_IO;
        fwrite(&primary, 8, 1, fs);
        fwrite(&L2[1], 8, 4, fs);
        fwrite(&sa_intv, 8, 1, fs);
        fwrite(&seq_len, 8, 1, fs);
        fwrite(sa.data() + 1, 8, (size_t) (n_sa - 1), fs);
        if (fclose(fs) != 0) return UNC_E_IO;
    }
    return UNC_OK;
}
