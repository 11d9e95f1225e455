// Host logic of the HS_separate_reads drop-in; see hs_sepreads.h. Every function states the reference lines
// whose observable behaviour it reproduces (reference = RolandFaure/Hairsplitter, src/separate_reads.cpp and
// src/cluster_graph.cpp). Where the reference's result depends on the iteration order of a libstdc++ container
// or on std::sort's treatment of equal keys, the same container / the same call is used here on the same
// sequence of operations, so the result is the same by construction.
#include "hs_sepreads.h"

#include "hs_colbin.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>

#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace hs {

// ---- .col parser (src/separate_reads.cpp:46-190) ---------------------------------------------------------------
namespace {

// whitespace-separated fields of a line, like successive `iss >> token`
struct Fields {
    const char* p;
    const char* end;
    Fields(const char* b, const char* e) : p(b), end(e) {}
    // the classic locale's whitespace (what `iss >>` skips), inline: std::isspace is a library call per character
    static bool blank(char c) { return c == ' ' || (unsigned char)(c - 9) < 5; }
    bool next(const char*& b, const char*& e) {
        while (p < end && blank(*p)) p++;
        if (p >= end) return false;
        b = p;
        while (p < end && !blank(*p)) p++;
        e = p;
        return true;
    }
    std::string next_string() {
        const char *b, *e;
        return next(b, e) ? std::string(b, e) : std::string();
    }
};

int stoi_or_throw(const std::string& s) { return std::stoi(s); }

// a token of 1..9 decimal digits (what the writer emits): parsed in place; anything else goes through the
// library call the reference uses (std::stoi / atoi semantics on signs, blanks, junk, overflow)
inline bool small_uint(const char* b, const char* e, int& v) {
    const size_t n = (size_t)(e - b);
    if (n == 0 || n > 9) return false;
    int x = 0;
    for (const char* q = b; q < e; q++) {
        const unsigned d = (unsigned)(*q - '0');
        if (d > 9) return false;
        x = x * 10 + (int)d;
    }
    v = x;
    return true;
}

// the lines of one CONTIG block [b, e) (its CONTIG line first); same line-by-line behaviour as the sequential
// loop of parse_column_file (:62-186)
void parse_col_block(const char* b, const char* e, bool numbers, int max_coverage, float rarest_strain_abundance, ColContig& c) {
    std::vector<int> read_idxs, codes;
    bool have_contig = false;
    for (const char* ls = b; ls < e;) {
        const char* le = (const char*)std::memchr(ls, '\n', (size_t)(e - ls));
        if (!le) le = e;
        Fields f(ls, le);
        const char *tb, *te;
        const char* line_b = ls;
        const char* line_e = le;
        ls = le + 1;
        if (!f.next(tb, te)) continue;
        const size_t tl = (size_t)(te - tb);
        if (tl == 6 && !std::memcmp(tb, "CONTIG", 6)) {
            have_contig = true;
            c.line.assign(line_b, line_e);
            f.next_string();  // name
            c.length = std::atoi(f.next_string().c_str());
            const std::string cov = f.next_string();
            c.coverage = cov.empty() ? 0.0 : std::strtod(cov.c_str(), nullptr);
        } else if (tl == 4 && !std::memcmp(tb, "SNPS", 4)) {
            if (!have_contig) continue;  // the reference would index snps[-1]
            const char *pb = nullptr, *pe = nullptr, *rb_ = nullptr, *re_ = nullptr, *sb_ = nullptr, *se_ = nullptr;
            const bool hp = f.next(pb, pe), hr = f.next(rb_, re_), hs = f.next(sb_, se_);
            const std::string pos = hp ? std::string(pb, pe) : std::string();
            uint8_t ref_base, second_base;
            if (numbers) {
                int v;
                ref_base = (hr && small_uint(rb_, re_, v)) ? (uint8_t)(char)v : (uint8_t)(char)stoi_or_throw(hr ? std::string(rb_, re_) : std::string());
                second_base = (hs && small_uint(sb_, se_, v)) ? (uint8_t)(char)v : (uint8_t)(char)stoi_or_throw(hs ? std::string(sb_, se_) : std::string());
            } else {
                ref_base = (uint8_t)(hr ? *rb_ : 0);
                second_base = (uint8_t)(hs ? *sb_ : 0);
            }
            const char *ib = nullptr, *ie = nullptr, *cb = nullptr, *ce = nullptr;
            const bool have_idx = f.next(ib, ie);
            const bool have_content = f.next(cb, ce);
            // content: comma-terminated integers (or characters), one per read (:106-124)
            codes.clear();
            if (have_content) {
                const char* t0 = cb;
                for (const char* q = cb; q < ce; q++) {
                    if (*q != ',') continue;
                    if (numbers) {
                        int v;
                        if (!small_uint(t0, q, v)) v = stoi_or_throw(std::string(t0, q));
                        codes.push_back((int)(uint8_t)v);
                    } else {
                        for (const char* z = t0; z < q; z++) codes.push_back((int)(uint8_t)*z);
                    }
                    t0 = q + 1;
                }
            }
            // a sixth field, when present, replaces the read indices (`iss >> readsIdx` a second time, :126)
            {
                const char *xb, *xe;
                if (f.next(xb, xe)) { ib = xb; ie = xe; }
            }
            read_idxs.clear();
            if (have_idx) {
                const char* t0 = ib;
                for (const char* q = ib; q < ie; q++) {
                    if (*q != ',') continue;
                    int v;
                    if (!small_uint(t0, q, v)) v = std::atoi(std::string(t0, q).c_str());
                    read_idxs.push_back(v);
                    t0 = q + 1;
                }
            }
            Column snp;
            int pv;
            snp.pos = small_uint(pos.data(), pos.data() + pos.size(), pv) ? pv : std::atoi(pos.c_str());
            snp.ref_base = ref_base;
            snp.second_base = second_base;
            snp.content.reserve(codes.size());
            snp.readIdxs.reserve(codes.size());
            int cov_maj = 0, cov_sec = 0, cov = 0;
            for (size_t n = 0; n < codes.size(); n++) {
                const int idx = n < read_idxs.size() ? read_idxs[n] : 0;
                if (codes[n] != ' ' && cov < max_coverage) {
                    snp.content.push_back((uint8_t)codes[n]);
                    snp.readIdxs.push_back((uint32_t)idx);
                    if ((uint8_t)codes[n] == ref_base) cov_maj++;
                    else if ((uint8_t)codes[n] == second_base) cov_sec++;
                }
                if (codes[n] != ' ' && idx >= 0) cov++;
            }
            if ((float)cov_sec >= rarest_strain_abundance * (float)(cov_maj + cov_sec)) c.snps.push_back(std::move(snp));
        } else if (tl == 4 && !std::memcmp(tb, "READ", 4)) {
            if (!have_contig) continue;
            c.read_lines.emplace_back(line_b, line_e);
            f.next_string();  // name
            f.next_string();  // start on the read
            f.next_string();  // end on the read
            const std::string start_contig = f.next_string(), end_contig = f.next_string();
            try {
                c.limits.emplace_back(std::stoi(start_contig), std::stoi(end_contig));
            } catch (const std::invalid_argument&) {
                std::cout << "error in parsing read limits" << std::endl;
                std::cout << "line : " << std::string(line_b, line_e) << std::endl;
                std::exit(1);
            }
        }
    }
}

}  // namespace

void parse_col_head(const char* b, const char* e, ColContig& c) { parse_col_block(b, e, true, 0, 0.0f, c); }

// The reference reads the file line by line into one growing structure (:62-186). Here the file is mapped once, cut at
// its CONTIG lines and the blocks are parsed in parallel: a CONTIG block only depends on the letters-or-numbers switch,
// which the first SNPS line of the file sets (:93-103).
void parse_column_file(const std::string& path, std::vector<ColContig>& contigs, int max_coverage,
                       float rarest_strain_abundance) {
    // the writer's binary sidecar, when it belongs to this very file (hs_colbin.h): same structures, no text to tokenise
    if (read_col_sidecar(path, contigs, max_coverage, rarest_strain_abundance)) {
        if (std::getenv("HS_TIMING")) fprintf(stderr, "[hs timing] (.col read from its binary sidecar)\n");
        return;
    }
    parse_column_text(path, contigs, max_coverage, rarest_strain_abundance);
}

void parse_column_text(const std::string& path, std::vector<ColContig>& contigs, int max_coverage,
                       float rarest_strain_abundance) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return;  // an unreadable file leaves no contigs, like the reference's failed ifstream
    struct stat sb;
    if (::fstat(fd, &sb) != 0 || sb.st_size == 0) {
        ::close(fd);
        return;
    }
    const size_t size = (size_t)sb.st_size;
    void* map = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) {
        std::cout << "ERROR: cannot map " << path << std::endl;
        std::exit(1);
    }
    ::madvise(map, size, MADV_SEQUENTIAL);
    const char* base = (const char*)map;
    const char* end = base + size;
    // block starts: lines whose first field is CONTIG. Found in parallel over slices of the file.
    std::vector<const char*> starts;
    {
        const int nt = std::max(1, omp_get_max_threads());
        std::vector<std::vector<const char*>> found((size_t)nt);
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nt; t++) {
            const char* lo = base + size * (size_t)t / (size_t)nt;
            const char* hi = base + size * (size_t)(t + 1) / (size_t)nt;
            if (t > 0 && lo > base) {  // first line that starts inside the slice
                const char* nl = (const char*)std::memchr(lo - 1, '\n', (size_t)(end - (lo - 1)));
                lo = nl ? nl + 1 : end;
            }
            for (const char* ls = lo; ls < hi;) {
                const char* le = (const char*)std::memchr(ls, '\n', (size_t)(end - ls));
                if (!le) le = end;
                Fields f(ls, le);
                const char *tb, *te;
                if (f.next(tb, te) && te - tb == 6 && !std::memcmp(tb, "CONTIG", 6)) found[t].push_back(ls);
                ls = le + 1;
            }
        }
        for (const auto& v : found) starts.insert(starts.end(), v.begin(), v.end());
    }
    if (starts.empty()) {
        ::munmap(map, size);
        return;
    }
    // letters or numbers: decided by the first SNPS line that follows a CONTIG line (:93-103)
    bool numbers = false;
    for (const char* ls = starts[0]; ls < end;) {
        const char* le = (const char*)std::memchr(ls, '\n', (size_t)(end - ls));
        if (!le) le = end;
        Fields f(ls, le);
        const char *tb, *te;
        if (f.next(tb, te) && te - tb == 4 && !std::memcmp(tb, "SNPS", 4)) {
            f.next_string();
            const std::string ref_s = f.next_string();
            numbers = !ref_s.empty() && !std::isalpha((unsigned char)ref_s[0]) && ref_s[0] != '-';
            break;
        }
        ls = le + 1;
    }
    const size_t first = contigs.size();
    contigs.resize(first + starts.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t b = 0; b < starts.size(); b++)
        parse_col_block(starts[b], b + 1 < starts.size() ? starts[b + 1] : end, numbers, max_coverage, rarest_strain_abundance,
                        contigs[first + b]);
    ::munmap(map, size);
}

// ---- shuffled sweep orders ----------------------------------------------------------------------------------------
Shuffler::Shuffler() {
    const char* s = std::getenv("HS_PIN_SEED");
    if (s && *s) {
        pinned = true;
        pin = (uint32_t)std::strtoul(s, nullptr, 10);
    }
}

const std::vector<int>& Shuffler::order(int n) {
    if (pinned) {
        auto it = cache.find(n);
        if (it != cache.end()) return it->second;
        std::vector<int>& o = cache[n];
        o.resize((size_t)n);
        std::iota(o.begin(), o.end(), 0);
        std::mt19937 g(pin);
        std::shuffle(o.begin(), o.end(), g);
        return o;
    }
    scratch.resize((size_t)n);
    std::iota(scratch.begin(), scratch.end(), 0);
    std::mt19937 g(seed());
    std::shuffle(scratch.begin(), scratch.end(), g);
    return scratch;
}

// ---- the window walk of main() (src/separate_reads.cpp:1523-1622, 1674-1678) ------------------------------------
void plan_windows(const ColContig& c, int W, std::vector<Window>& out) {
    out.clear();
    const std::vector<Column>& snps = c.snps;
    const long length = c.length;
    const int n_reads = (int)c.read_lines.size();
    size_t sidx = 0;  // suspectPostitionIdx
    int chunk = -1;
    std::vector<char> mask;
    while ((long)(chunk + 1) * W + 100 <= length) {
        chunk++;
        int upper = (chunk + 1) * W;
        const bool last = (long)(chunk + 1) * W + 100 > length;
        if (last) upper = (int)(length + 1);
        Window win;
        win.chunk = chunk;
        win.start = chunk * W;
        win.end = std::min(upper - 1, (int)length);
        if (sidx >= snps.size() || snps[sidx].pos > upper - 1) {
            // no SNP in the window: reads covering its middle get 0, the others -2 (:1543-1562)
            win.has_snps = false;
            win.reads_here.assign((size_t)n_reads, -2);
            const int left = chunk * W, right = std::min(upper - 1, (int)length);
            int middle = (left + right) / 2;
            if (middle < 500) middle = std::min(500, (int)(length / 2));
            if (middle > (int)length - 500) middle = std::max((int)(length / 2), (int)length - 500);
            for (size_t r = 0; r < c.limits.size(); r++)
                if (c.limits[r].first <= middle && c.limits[r].second >= middle) win.reads_here[r] = 0;
            out.push_back(std::move(win));
            continue;
        }
        win.has_snps = true;
        mask.assign((size_t)n_reads, 0);
        if (chunk == 0) {  // 20 % margin for the reads to start aligning (:1566-1570)
            while (sidx < snps.size() - 1 && snps[sidx].pos < chunk * W + 0.2 * W && snps[sidx + 1].pos < chunk * W + 0.4 * W) sidx++;
        }
        for (uint32_t r : snps[sidx].readIdxs) mask[r] = 1;
        while (sidx < snps.size() && snps[sidx].pos < upper - 1) sidx++;
        if (sidx > 0) sidx--;
        if (last) {  // and to stop aligning (:1582-1586)
            while (sidx > 0 && snps[sidx].pos > upper - 1 - 0.2 * W && snps[sidx - 1].pos > upper - 1 - 0.4 * W) sidx--;
        }
        // keep the reads also present at the last SNP; indices beyond its last read stay as they are (:1588-1596)
        uint32_t idxmask = 0;
        for (uint32_t r : snps[sidx].readIdxs) {
            while (idxmask < r) {
                mask[idxmask] = 0;
                idxmask++;
            }
            idxmask++;
        }
        sidx++;
        for (int r = 0; r < n_reads; r++)
            if (mask[r]) win.masked.push_back(r);
        int lastpos = -10;
        for (size_t s = 0; s < snps.size(); s++) {
            const int p = snps[s].pos;
            if (p >= chunk * W && p < chunk * W + W && p > lastpos + 10) {
                lastpos = p;
                win.restart_snps.push_back((int)s);
            }
        }
        out.push_back(std::move(win));
    }
}

void snp_start_labels(const Column& snp, const std::vector<char>& mask, std::vector<int>& labels) {
    const int n = (int)mask.size();
    labels.resize((size_t)n);
    for (int r = 0; r < n; r++) labels[r] = r;
    int first[256];
    std::fill(first, first + 256, -1);
    for (size_t r = 0; r < snp.content.size(); r++) {
        const uint32_t read = snp.readIdxs[r];
        if (mask[read]) {
            if (first[snp.content[r]] < 0) first[snp.content[r]] = (int)read;
            labels[read] = first[snp.content[r]];
        }
    }
}

// ---- chinese whispers (src/cluster_graph.cpp:152-310) -----------------------------------------------------------
std::vector<int> chinese_whispers(const ReadGraph& g, const std::vector<int>& initial, const std::vector<char>& mask,
                                  Shuffler& sh) {
    std::vector<int> clusters = initial;
    const int n = (int)initial.size();
    std::vector<int> count((size_t)mask.size(), 0);
    int changes = 3, iterations = 0;
    while (changes > 2 && iterations < 15) {
        changes = 0;
        const std::vector<int>& order = sh.order(n);
        for (int i : order) {
            if (!mask[i]) continue;
            int max_index = 0, max_value = 0;
            for (int e = g.off[i]; e < g.off[i + 1]; e++) {
                const int l = clusters[g.nbr[e]];
                if (l >= 0) count[l]++;
            }
            for (int e = g.off[i]; e < g.off[i + 1]; e++) {  // lowest label among the most frequent
                const int l = clusters[g.nbr[e]];
                if (l < 0) continue;
                if (count[l] > max_value || (count[l] == max_value && l < max_index)) {
                    max_value = count[l];
                    max_index = l;
                }
            }
            for (int e = g.off[i]; e < g.off[i + 1]; e++) {
                const int l = clusters[g.nbr[e]];
                if (l >= 0) count[l] = 0;
            }
            if (max_value > 0) {
                if (clusters[i] != max_index) changes++;
                clusters[i] = max_index;
            }
        }
        iterations++;
    }
    for (size_t i = 0; i < mask.size(); i++)
        if (!mask[i]) clusters[i] = -2;
    return clusters;
}

// ---- low-memory read graph (src/separate_reads.cpp:538-693) -----------------------------------------------------
void create_read_graph_low_memory(const std::vector<Column>& snps, const std::vector<char>& mask, ReadGraph& g,
                                  float error_rate) {
    const int n = (int)mask.size();
    std::vector<std::pair<int, std::vector<int>>> reads((size_t)n, std::make_pair(-1, std::vector<int>()));
    int idx_snp = 0;
    for (const Column& snp : snps) {
        for (size_t r = 0; r < snp.readIdxs.size(); r++) {
            auto& rd = reads[snp.readIdxs[r]];
            if (rd.first == -1) rd.first = idx_snp;
            if (snp.content[r] == snp.ref_base) rd.second.push_back(1);
            else if (snp.content[r] == snp.second_base) rd.second.push_back(2);
            else rd.second.push_back(0);
        }
        idx_snp++;
    }
    std::vector<char> mask_extend((size_t)n, 0);
    for (int r = 0; r < n; r++)
        if (mask[r] && reads[r].first != -1) mask_extend[r] = 1;

    std::vector<std::vector<int>> lists((size_t)n);
    std::vector<float> dist((size_t)n, 0);
    std::vector<int> sim((size_t)n, 0), diff((size_t)n, 0);
    std::vector<std::pair<int, float>> smallest;
    for (int read1 = 0; read1 < n; read1++) {
        if (!mask_extend[read1]) continue;
        dist.assign((size_t)n, 0);
        sim.assign((size_t)n, 0);
        diff.assign((size_t)n, 0);
        int max_compat = 0;
        for (int read2 = 0; read2 < n; read2++) {
            if (!mask_extend[read2] || read1 == read2) continue;
            int nb_similar = 0, nb_different = 0;
            const int first_common = std::max(reads[read1].first, reads[read2].first);
            const int last_common = (int)std::min(reads[read1].second.size() + reads[read1].first - 1,
                                                  reads[read2].second.size() + reads[read2].first - 1);
            for (int pos = first_common; pos <= last_common; pos++) {
                const int v1 = reads[read1].second[pos - reads[read1].first];
                const int v2 = reads[read2].second[pos - reads[read2].first];
                if (v1 == 2 && v2 == 2) nb_similar += 3;
                else if (v1 == 1 && v2 == 1) nb_similar++;
                else if (v1 != 0 && v2 != 0) nb_different++;
            }
            dist[read2] = 1 - std::max(0, nb_different - 1) / float(nb_different + nb_similar);  // may be NaN (0/0)
            if (nb_similar > max_compat) max_compat = nb_similar;
            sim[read2] = nb_similar;
            diff[read2] = nb_different;
        }
        for (int r = 0; r < n; r++)
            if (mask[r] && r != read1 && sim[r] + diff[r] < 0.7 * max_compat) dist[r] = 0;
        smallest.clear();
        for (int r = 0; r < n; r++) smallest.push_back(std::make_pair(r, dist[r]));
        std::sort(smallest.begin(), smallest.end(),
                  [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });
        int nb = 0;
        const float below = 1 - error_rate * 2;
        float above = 1;
        if (smallest.size() > 1) above = smallest[0].second - (smallest[0].second - smallest[1].second) * 3;
        if (above == 1) {
            int idx = 0;
            while (idx < (int)smallest.size() && smallest[idx].second == 1) idx += 1;
            if (idx < (int)smallest.size()) {
                idx = std::min(idx + 4, (int)smallest.size() - 1);
                above = smallest[idx].second;
            }
        }
        for (const auto& nbr : smallest) {
            if (nbr.second > below && (nb < 5 || nbr.second == 1 || nbr.second >= above) && mask[nbr.first]) {
                nb++;
                lists[read1].push_back(nbr.first);
                lists[nbr.first].push_back(read1);
            }
        }
    }
    g.clear(n);
    g.list_mode = true;
    for (int r = 0; r < n; r++) {
        std::sort(lists[r].begin(), lists[r].end());
        lists[r].erase(std::unique(lists[r].begin(), lists[r].end()), lists[r].end());
        g.off[r + 1] = g.off[r] + (int)lists[r].size();
        g.nbr.insert(g.nbr.end(), lists[r].begin(), lists[r].end());
    }
}

// ---- post-processing of the clusterings of one window ------------------------------------------------------------
namespace {

// merge_clusterings (src/separate_reads.cpp:840-885)
std::vector<int> merge_clusterings(const std::vector<std::vector<int>>& local, const ReadGraph& g,
                                   const std::vector<char>& mask, Shuffler& sh) {
    std::vector<double> aggregated(local[0].size(), 0);
    for (size_t i = 0; i < local.size(); i++) {
        const double w = std::pow(2.0, (double)i);
        for (size_t j = 0; j < local[i].size(); j++) aggregated[j] += local[i][j] * w;
    }
    std::unordered_map<double, int> seen;
    std::vector<int> ints;
    ints.reserve(aggregated.size());
    int index = 0;
    for (size_t i = 0; i < aggregated.size(); i++) {
        auto it = seen.find(aggregated[i]);
        if (it == seen.end()) {
            seen[aggregated[i]] = index;
            ints.push_back(index);
            index++;
        } else {
            ints.push_back(it->second);
        }
    }
    for (size_t i = 0; i < ints.size(); i++)
        if (!mask[i]) ints[i] = -2;
    return chinese_whispers(g, ints, mask, sh);
}

// merge_close_clusters (src/cluster_graph.cpp:402-501)
void merge_close_clusters(const ReadGraph& g, std::vector<int>& clusters, const std::vector<char>& mask, Shuffler& sh) {
    const int n = (int)clusters.size();
    std::set<int> tested;
    std::vector<int> initial_count((size_t)n, 0);
    for (int c : clusters)
        if (c >= 0 && c < n) initial_count[c] += 1;
    std::vector<int> votes((size_t)mask.size(), 0);
    std::vector<int> touched;
    for (int node = 0; node < n; node++) {
        if (tested.find(clusters[node]) != tested.end() || clusters[node] < 0) continue;
        const int cluster_to_test = clusters[node];
        std::vector<int> newclusters = clusters;
        int changes = 3;
        std::vector<int> count = initial_count;
        int iterations = 0;
        while (changes > 0 && iterations < 10) {
            changes = 0;
            const std::vector<int>& order = sh.order(n);
            for (int i : order) {
                if (!mask[i] || newclusters[i] != cluster_to_test) continue;
                touched.clear();
                if (g.list_mode) {
                    // the reference tests the POSITIONS 0..deg-1 for membership in the list (:443-447)
                    const int deg = g.off[i + 1] - g.off[i];
                    for (int j = 0; j < deg; j++) {
                        if (std::binary_search(g.nbr.begin() + g.off[i], g.nbr.begin() + g.off[i + 1], j) && newclusters[j] >= 0) {
                            if (votes[newclusters[j]]++ == 0) touched.push_back(newclusters[j]);
                        }
                    }
                } else {
                    for (int e = g.off[i]; e < g.off[i + 1]; e++) {
                        const int l = newclusters[g.nbr[e]];
                        if (l >= 0 && votes[l]++ == 0) touched.push_back(l);
                    }
                }
                std::sort(touched.begin(), touched.end());
                int max_index = 0, max_value = 0, second_index = 0, second_value = 0;
                for (int j : touched) {  // the scan over all labels only ever reacts to the non-zero ones
                    if (votes[j] > max_value) {
                        second_value = max_value;
                        second_index = max_index;
                        max_value = votes[j];
                        max_index = j;
                    } else if (votes[j] > second_value) {
                        second_value = votes[j];
                        second_index = j;
                    }
                }
                for (int j : touched) votes[j] = 0;
                if (max_value > 0 && max_index != cluster_to_test) {
                    count[newclusters[i]]--;
                    count[max_index]++;
                    changes++;
                    newclusters[i] = max_index;
                } else if (max_value > 0 && max_value <= 2 * second_value) {
                    count[newclusters[i]]--;
                    count[second_index]++;
                    newclusters[i] = second_index;
                    changes++;
                }
            }
            iterations++;
        }
        tested.emplace(cluster_to_test);
        if (count[cluster_to_test] == 0) {
            clusters = newclusters;
            initial_count = count;
        }
    }
}

// merge_wrongly_split_haplotypes (src/separate_reads.cpp:1007-1327)
std::vector<int> merge_wrongly_split_haplotypes(const std::vector<int>& clustered, const std::vector<Column>& snps,
                                                const ReadGraph& g, int posstart, int posend) {
    using std::unordered_map;
    std::set<int> groups;
    unordered_map<int, int> index_of_group;
    int index = 0;
    for (size_t read = 0; read < clustered.size(); read++) {
        if (clustered[read] > -1) {
            groups.emplace(clustered[read]);
            if (index_of_group.find(clustered[read]) == index_of_group.end()) {
                index_of_group[clustered[read]] = index;
                index++;
            }
        }
    }
    const size_t ng = groups.size();
    if (ng <= 1) {
        std::vector<int> one((size_t)clustered.size(), 0);
        for (size_t r = 0; r < clustered.size(); r++)
            if (clustered[r] == -2) one[r] = -2;
        return one;
    }
    std::vector<std::vector<int>> incompat(ng, std::vector<int>(ng, 0));
    std::vector<std::vector<int>> last_pos(ng, std::vector<int>(ng, -10));

    for (const Column& snp : snps) {
        if (!(snp.pos >= posstart && snp.pos < posend)) continue;
        // the same containers, filled in the same order, as the reference: the majority base of a cluster is
        // picked with `>=` while iterating an unordered_map, so ties follow its iteration order (:1063-1083)
        unordered_map<int, unsigned char> majority;
        unordered_map<int, unordered_map<unsigned char, int>> bases_in_cluster;
        unordered_map<int, int> n_in_cluster;
        for (size_t r = 0; r < snp.readIdxs.size(); r++) {
            const int read = (int)snp.readIdxs[r];
            const unsigned char base = snp.content[r];
            const int cluster = clustered[read];
            if (cluster > -1) {
                if (bases_in_cluster.find(cluster) == bases_in_cluster.end()) bases_in_cluster[cluster] = unordered_map<unsigned char, int>();
                if (bases_in_cluster[cluster].find(base) == bases_in_cluster[cluster].end()) bases_in_cluster[cluster][base] = 0;
                bases_in_cluster[cluster][base]++;
                n_in_cluster[cluster]++;
            }
        }
        std::set<unsigned char> maxbases;
        for (auto cluster : bases_in_cluster) {
            int second_max = 0, max = 0;
            char max_base = ' ';
            for (auto base : cluster.second) {
                if (base.second >= max) {
                    max_base = (char)base.first;
                    second_max = max;
                    max = base.second;
                } else if (base.second > second_max) {
                    second_max = base.second;
                }
            }
            if (second_max * 2 > max || n_in_cluster[cluster.first] * 0.5 > max) max_base = ' ';
            majority[cluster.first] = (unsigned char)max_base;
            if (max_base != ' ') maxbases.emplace((unsigned char)max_base);
        }
        if (maxbases.size() <= 1) continue;
        for (int g1 : groups) {
            for (int g2 : groups) {
                // operator[] on a cluster absent from this SNP yields 0, which counts as a base (:1096)
                if (majority[g1] != ' ' && majority[g2] != ' ' && g1 > g2) {
                    const int i1 = index_of_group[g1], i2 = index_of_group[g2];
                    if (majority[g1] != majority[g2] && snp.pos - last_pos[i1][i2] > 10) {
                        incompat[i1][i2] += 1;
                        incompat[i2][i1] += 1;
                        last_pos[i1][i2] = snp.pos;
                        last_pos[i2][i1] = snp.pos;
                    }
                }
            }
        }
    }

    // links between clusters, as a fraction of the links of the first (:1166-1213)
    std::map<std::pair<int, int>, double> links;
    unordered_map<int, int> links_of_cluster;
    for (int k = 0; k < g.n; k++) {
        for (int e = g.off[k]; e < g.off[k + 1]; e++) {
            const int c1 = clustered[g.nbr[e]], c2 = clustered[k];
            if (c1 != c2) {
                auto key = std::make_pair(c1, c2);
                if (links.find(key) == links.end()) links[key] = 0;
                links[key]++;
            }
            if (links_of_cluster.find(c1) == links_of_cluster.end()) links_of_cluster[c1] = 0;
            links_of_cluster[c1]++;
        }
    }
    for (auto link : links) links[link.first] = link.second / links_of_cluster[link.first.first];
    std::vector<std::pair<std::pair<int, int>, double>> sorted_links;
    for (auto link : links) sorted_links.emplace_back(link);
    std::sort(sorted_links.begin(), sorted_links.end(),
              [](const std::pair<std::pair<int, int>, double>& a, const std::pair<std::pair<int, int>, double>& b) {
                  return a.second > b.second;
              });

    unordered_map<int, int> new_group;
    for (int group : groups) new_group[group] = group;
    new_group[-1] = -1;
    new_group[-2] = -2;
    for (const auto& pr : sorted_links) {
        if (!(pr.second > 0.01)) continue;
        const int c1 = pr.first.first, c2 = pr.first.second;
        if (new_group[c1] == new_group[c2]) continue;
        bool incompatibility = false;
        for (int g1 : groups) {
            if (new_group[g1] != new_group[c1]) continue;
            for (int g2 : groups) {
                if (new_group[g2] == new_group[c2] && incompat[index_of_group[g1]][index_of_group[g2]] > 1) incompatibility = true;
            }
        }
        if (!incompatibility) {
            for (int g2 : groups) {
                // new_group[c2] is re-read in every turn, as in the reference (:1266-1270)
                if (new_group[g2] == new_group[c2]) new_group[g2] = new_group[c1];
            }
        }
    }
    unordered_map<int, int> new_index;
    int next = 0;
    for (int group : groups) {
        if (new_index.find(new_group[group]) == new_index.end()) {
            new_index[new_group[group]] = next;
            next++;
        }
    }
    for (int group : groups) new_group[group] = new_index[new_group[group]];
    std::vector<int> out((size_t)clustered.size(), -1);
    for (size_t read = 0; read < clustered.size(); read++) out[read] = new_group[clustered[read]];
    return out;
}

}  // namespace

void finalize_clustering(const std::vector<Column>& snps, const std::vector<std::vector<int>>& local_clusters,
                         const ReadGraph& g, bool low_memory, const std::vector<char>& mask, std::vector<int>& haplotypes,
                         int posstart, int posend, Shuffler& sh) {
    (void)low_memory;  // the graph already is the representation this flag selects
    const size_t n = mask.size();
    if (local_clusters.empty()) {
        for (size_t r = 0; r < n; r++) haplotypes[r] = mask[r] ? -1 : -2;
        return;
    }
    std::vector<int> clustered = merge_clusterings(local_clusters, g, mask, sh);
    // clusters of fewer than 5 reads become -1 (:921-936)
    std::unordered_map<int, int> sizes;
    for (size_t r = 0; r < clustered.size(); r++) {
        if (!mask[r]) clustered[r] = -2;
        else sizes[clustered[r]] += 1;
    }
    for (size_t r = 0; r < clustered.size(); r++)
        if (sizes[clustered[r]] < 5 && clustered[r] != -2) clustered[r] = -1;
    std::vector<int> merged = clustered;
    std::unordered_map<int, int> to_haplotype;
    int haplotype = 0;
    for (size_t r = 0; r < merged.size(); r++) {
        if (merged[r] > -1) {
            if (to_haplotype.find(merged[r]) == to_haplotype.end()) {
                to_haplotype[merged[r]] = haplotype;
                haplotype++;
            }
            merged[r] = to_haplotype[merged[r]];
        }
    }
    haplotypes = chinese_whispers(g, merged, mask, sh);
    std::unordered_map<int, int> to_index;
    to_index[-1] = snps.empty() ? 0 : -1;
    to_index[-2] = -2;
    int index_h = 0;
    for (int h : haplotypes) {
        if (to_index.find(h) == to_index.end()) {
            to_index[h] = index_h;
            index_h += 1;
        }
    }
    for (size_t r = 0; r < haplotypes.size(); r++) haplotypes[r] = to_index[haplotypes[r]];
    merge_close_clusters(g, haplotypes, mask, sh);
    haplotypes = merge_wrongly_split_haplotypes(haplotypes, snps, g, posstart, posend);
}

std::vector<int> merge_haplotypes_to_fit_within_limit(int max_haplotypes, const std::vector<int>& clusters,
                                                      const std::vector<char>& mask, const ReadGraph& g, Shuffler& sh) {
    std::unordered_map<int, int> count;
    for (int c : clusters)
        if (c >= 0) count[c] += 1;
    if ((int)count.size() <= max_haplotypes) return clusters;
    std::vector<std::pair<int, int>> by_size;
    for (auto c : count) by_size.push_back(std::make_pair(c.second, c.first));
    std::sort(by_size.begin(), by_size.end(), std::greater<std::pair<int, int>>());
    std::set<int> kept;
    for (int i = 0; i < max_haplotypes; i++) kept.insert(by_size[i].second);
    std::vector<int> next = clusters;
    for (size_t i = 0; i < clusters.size(); i++)
        if (clusters[i] >= 0 && kept.find(clusters[i]) == kept.end()) next[i] = -1;
    return chinese_whispers(g, next, mask, sh);
}

}  // namespace hs
