#include "hs_partition.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "../csrc/rank.cuh"  // the reference's robin_hood slot order (plain C++ when not compiled by nvcc)

namespace hs {

// The most frequent code other than `ref` in a histogram filled in first-seen order, the way the
// reference finds it: iterate a robin_hood::unordered_map<unsigned char,int> and keep the first strict
// maximum (src/Partition.cpp:56-69, src/call_variants.cpp:832-844). The map's slot order only matters
// when several codes share the maximum; then it is replayed (rank.cuh). `ref_competes`: the
// Partition/Column overload compares a `char` with an `unsigned char`, which never match above 127.
static int second_most_frequent(const std::vector<uint8_t>& seen, const int* count, int ref, bool ref_competes,
                                bool ref_in_map) {
    int best = -1, alt = -1, ties = 0;
    for (uint8_t key : seen) {
        if (key == ref && !ref_competes) continue;
        if (count[key] > best) { best = count[key]; alt = key; ties = 1; }
        else if (count[key] == best) ties++;
    }
    if (ties <= 1) return alt;
    HsRhTable t;
    hs_rh_new(t);
    for (uint8_t key : seen) hs_rh_insert(t, key);
    if (!ref_in_map) hs_rh_insert(t, (uint8_t)ref);  // content[ref_base] creates the entry
    uint8_t order[HS_RH_MAXKEYS];
    const int n = hs_rh_iterate(t, order);
    for (int i = 0; i < n; i++) {
        const int key = order[i];
        if (key == ref && !ref_competes) continue;
        if (count[key] == best) return key;
    }
    return alt;
}

Partition::Partition(const Column& snp, int pos, unsigned char ref_base) {
    pos_left = pos_right = pos;
    readIdx.assign(snp.readIdxs.begin(), snp.readIdxs.end());
    int count[256] = {0};
    std::vector<uint8_t> seen;
    for (uint8_t c : snp.content)
        if (count[c]++ == 0) seen.push_back(c);
    const bool ref_present = count[ref_base] > 0;
    const int alt = second_most_frequent(seen, count, ref_base, false, ref_present);
    state.reserve(snp.content.size());
    for (uint8_t c : snp.content) state.push_back(c == ref_base ? 1 : ((int)c == alt ? -1 : 0));
    more.assign(snp.content.size(), 1);
    less.assign(snp.content.size(), 0);
    occurrences = 1;
}

void Partition::augment(const Column& sup, int pos) {
    if (pos != -1) {
        if (pos < pos_left || pos_left == -1) pos_left = pos;
        if (pos > pos_right) pos_right = pos;
    }
    if (sup.readIdxs.empty()) return;
    // the two most frequent symbols other than ' ' (lowest value wins ties; 255 is never looked at)
    int count[256] = {0};
    for (uint8_t c : sup.content) count[c]++;
    int first = 'b', second = 'b', best = -1;
    for (int i = 0; i < 255; i++)
        if (i != ' ' && count[i] > best) { first = i; best = count[i]; }
    best = -1;
    for (int i = 0; i < 255; i++)
        if (i != first && i != ' ' && count[i] > best) { second = i; best = count[i]; }
    // orientation: does `first` go with +1 or with -1 ?
    {
        size_t n2 = 0;
        int agree = 0;
        for (size_t n1 = 0; n1 < readIdx.size(); n1++) {
            const unsigned int read = (unsigned int)readIdx[n1];
            while (n2 < sup.readIdxs.size() && sup.readIdxs[n2] < read) n2++;
            if (n2 >= sup.readIdxs.size()) break;
            if (sup.readIdxs[n2] == read) {
                const int c = sup.content[n2], s = state[n1];
                if (c == first && s == 1) agree++;
                else if (c == first && s == -1) agree--;
                else if (c == second && s == -1) agree++;
                else if (c == second && s == 1) agree--;
            }
        }
        if (agree < 0) {  // the swap goes through a `char` in the reference; both symbols are below 128 here
            std::swap(first, second);
        }
    }
    std::vector<int> idx2, more2, less2;
    std::vector<short> state2;
    const size_t cap = readIdx.size() + sup.readIdxs.size();
    idx2.reserve(cap); more2.reserve(cap); less2.reserve(cap); state2.reserve(cap);
    size_t n1 = 0;
    auto keep_old = [&](size_t i) {
        idx2.push_back(readIdx[i]); state2.push_back(state[i]); more2.push_back(more[i]); less2.push_back(less[i]);
    };
    for (size_t n2 = 0; n2 < sup.readIdxs.size(); n2++) {
        const int read = (int)sup.readIdxs[n2];
        while (n1 < readIdx.size() && readIdx[n1] < read) keep_old(n1++);
        short s = 0;
        if (sup.content[n2] == second) s = -1;
        if (sup.content[n2] == first) s = 1;
        if (n1 >= readIdx.size() || readIdx[n1] != read) {  // a read the partition did not know
            idx2.push_back(read); state2.push_back(s); more2.push_back(std::abs(s)); less2.push_back(0);
            continue;
        }
        const short old = state[n1];
        if (old == -2 || s == 0) {            // masked, or the column says nothing about this read
            keep_old(n1);
        } else if (old == 0) {
            idx2.push_back(read); state2.push_back(s); more2.push_back(1); less2.push_back(0);
        } else if (s == old) {
            idx2.push_back(read); state2.push_back(old); more2.push_back(more[n1] + 1); less2.push_back(less[n1]);
        } else {                               // s == -old: disagreement
            idx2.push_back(read);
            if (less[n1] + 1 > more[n1]) { state2.push_back((short)-old); more2.push_back(more[n1] + 1); less2.push_back(less[n1]); }
            else { state2.push_back(old); more2.push_back(more[n1]); less2.push_back(less[n1] + 1); }
        }
        n1++;
    }
    while (n1 < readIdx.size()) keep_old(n1++);
    readIdx.swap(idx2); state.swap(state2); more.swap(more2); less.swap(less2);
    occurrences += 1;
}

void Partition::merge(const Partition& o, short phased) {
    pos_left = std::min(pos_left, o.pos_left);
    pos_right = std::max(pos_right, o.pos_right);
    std::vector<int> idx2, more2, less2;
    std::vector<short> state2;
    size_t n1 = 0, n2 = 0;
    auto keep_mine = [&](size_t i) {
        idx2.push_back(readIdx[i]); state2.push_back(state[i]); more2.push_back(more[i]); less2.push_back(less[i]);
    };
    auto take_other = [&](size_t i) {
        idx2.push_back(o.readIdx[i]); state2.push_back((short)(o.state[i] * phased)); more2.push_back(o.more[i]); less2.push_back(o.less[i]);
    };
    while (n1 < readIdx.size() && n2 < o.readIdx.size()) {
        if (readIdx[n1] < o.readIdx[n2]) { keep_mine(n1++); continue; }
        if (readIdx[n1] > o.readIdx[n2]) { take_other(n2++); continue; }
        const short a = state[n1], b = o.state[n2];
        idx2.push_back(readIdx[n1]);
        if (a == 0 || b == -2) {
            state2.push_back((short)(b * phased)); more2.push_back(o.more[n2]); less2.push_back(o.less[n2]);
        } else if (b == 0 || a == -2) {
            state2.push_back(a); more2.push_back(more[n1]); less2.push_back(less[n1]);
        } else if (phased * b == a || phased * b == -a) {
            const bool agree = phased * b == a;
            const double cut = agree ? 0.9 : 0.8;
            const double conf1 = double(more[n1]) / (more[n1] + less[n1]);
            const double conf2 = double(o.more[n2]) / (o.more[n2] + o.less[n2]);
            int which = 0;  // 1: trust the other partition only, 2: trust this one only, 0: add both up
            if (conf1 < cut && conf2 > cut && o.more[n2] >= 10) which = 1;
            else if (conf2 < cut && conf1 > cut && more[n1] >= 10) which = 2;
            int m = 0, l = 0;
            if (which != 1) { m += more[n1]; l += less[n1]; }
            if (which != 2) {
                if (agree) { m += o.more[n2]; l += o.less[n2]; }
                else { m += o.less[n2]; l += o.more[n2]; }
            }
            short s = a;
            if (!agree && l > m) { s = (short)-s; std::swap(m, l); }  // the majority changed sides
            state2.push_back(s); more2.push_back(m); less2.push_back(l);
        } else {
            // states outside {1,-1,0,-2} cannot occur; the reference pushes the index only
            state2.push_back(a); more2.push_back(more[n1]); less2.push_back(less[n1]);
        }
        n1++;
        n2++;
    }
    while (n2 < o.readIdx.size()) take_other(n2++);
    while (n1 < readIdx.size()) keep_mine(n1++);
    readIdx.swap(idx2); state.swap(state2); more.swap(more2); less.swap(less2);
    occurrences += o.occurrences;
}

bool Partition::informative(bool last_read_biased, float mean_error) const {
    int suspicious[2] = {0, 0};
    int n_reads = 0;
    // the bound is unsigned in the reference (size() - adjust)
    const size_t bound = state.size() - (last_read_biased ? 1 : 0);
    for (size_t r = 0; r < bound; r++) {
        const int n = more[r] + less[r];
        float threshold = 0.5 * n + 3 * std::sqrt(n * 0.5 * (1 - 0.5));
        threshold = std::min(threshold, float(n) - 1);
        if (more[r] > threshold) {
            if (state[r] == -1) { suspicious[0]++; n_reads++; }
            else if (state[r] == 1) { suspicious[1]++; n_reads++; }
        }
    }
    const float min_reads = mean_error * n_reads / 2;
    return !(suspicious[0] < min_reads || suspicious[1] < min_reads);
}

static double log_binomial(double n, double k) { return std::lgamma(n + 1) - std::lgamma(k + 1) - std::lgamma(n - k + 1); }

float Partition::significance(int n_columns) const {
    int mutated = 0, reads = 0, columns = 0;
    for (size_t p = 0; p < state.size(); p++) {
        if (state[p] == -1 && more[p] > 1 && less[p] == 0) {
            mutated++;
            if (more[p] > columns) columns = more[p];
        }
        if (p != 0 && more[p] > 1 && less[p] == 0) reads++;  // the reference tests the INDEX against 0 and -2 (:210)
    }
    // log(float) resolves to the double overload in the reference's translation unit
    const double p_value = std::exp(std::log((double)(float(mutated) / reads)) * columns * mutated +
                                    log_binomial(reads, mutated) + log_binomial(n_columns, columns));
    return (float)std::max(0.0, p_value);
}

float Partition::confidence_score() const {
    double conf = 1;
    int n = 0;
    for (size_t c = 0; c < more.size(); c++) {
        float ci;  // getConfidence(), src/Partition.cpp:811-827
        if (state[c] == 0) ci = 0.5;
        else if (more[c] + less[c] > 0) ci = float(more[c]) / (more[c] + less[c]);
        else ci = 1;
        if (more[c] > 1) { conf *= ci; n++; }
    }
    if (conf == 1) conf = 0.99;
    return (float)(std::pow(1 / (1 - std::exp(std::log(conf) / n)), 2) * occurrences);
}

Distance distance(const Partition& part, const Column& col, char ref_base, bool want_column) {
    Distance res;
    // pass 1: histogram of the column over the reads the partition knows (masked reads excluded)
    int count[256] = {0};
    std::vector<uint8_t> seen;
    int n_bases = 0;
    {
        size_t n1 = 0;
        for (size_t n2 = 0; n2 < col.readIdxs.size(); n2++) {
            while (n1 < part.readIdx.size() && (unsigned int)part.readIdx[n1] < col.readIdxs[n2]) n1++;
            if (n1 >= part.readIdx.size()) break;
            if ((unsigned int)part.readIdx[n1] == col.readIdxs[n2] && part.state[n1] != -2) {
                if (count[col.content[n2]]++ == 0) seen.push_back(col.content[n2]);
                n_bases++;
            }
        }
    }
    if (n_bases == 0) {
        res.augmented = false;
        return res;
    }
    const unsigned char ref = (unsigned char)ref_base;
    const bool ref_competes = ref_base < 0;  // `char != unsigned char` (:838) never matches above 127
    const int alt_i = second_most_frequent(seen, count, ref, ref_competes, count[ref] > 0);
    const unsigned char alt = alt_i < 0 ? (unsigned char)' ' : (unsigned char)alt_i;
    if (want_column) {
        res.phased_column.readIdxs = col.readIdxs;
        res.phased_column.content.resize(col.content.size());
        for (size_t i = 0; i < col.content.size(); i++)
            res.phased_column.content[i] = col.content[i] == ref ? 'A' : (col.content[i] == alt ? 'a' : ' ');
    }
    // pass 2: the 2x2 table over reads with state +1 / -1
    size_t n1 = 0, n2 = 0;
    while (n1 < part.readIdx.size() && n2 < col.readIdxs.size()) {
        const unsigned int r1 = (unsigned int)part.readIdx[n1], r2 = col.readIdxs[n2];
        if (r1 < r2) { n1++; continue; }
        if (r2 < r1) { n2++; continue; }
        const int s = part.state[n1];
        const bool solid = part.less[n1] <= 1 && part.more[n1] >= 3;
        const unsigned char c = col.content[n2];
        if (c == ref) {
            if (s == 1) { res.n11++; res.solid11 += solid; }
            else if (s == -1) { res.n01++; res.solid01 += solid; }
        } else if (c == alt) {
            if (s == 1) { res.n10++; res.solid10 += solid; }
            else if (s == -1) { res.n00++; res.solid00 += solid; }
        }
        n1++;
        n2++;
    }
    res.phased = 1;
    res.second_base = alt;
    return res;
}

Distance distance(const Partition& p1, const Partition& p2, int threshold_p) {
    int scores[2] = {0, 0};
    short divergent[2] = {0, 0}, unsure[2] = {0, 0};
    int m00[2] = {0, 0}, m01[2] = {0, 0}, m10[2] = {0, 0}, m11[2] = {0, 0};
    int comparable = 0;
    size_t r1 = 0, r2 = 0;
    while (r1 < p1.readIdx.size() && r2 < p2.readIdx.size()) {
        if (p1.readIdx[r1] < p2.readIdx[r2]) { r1++; continue; }
        if (p2.readIdx[r2] < p1.readIdx[r1]) { r2++; continue; }
        if (p1.more[r1] > 1 && p2.more[r2] > 1) {
            comparable++;
            const int t1 = p1.more[r1] + p1.less[r1], t2 = p2.more[r2] + p2.less[r2];
            const float threshold1 = 0.5 * t1 + 3 * std::sqrt(t1 * 0.5 * (1 - 0.5));
            const float threshold2 = 0.5 * t2 + 3 * std::sqrt(t2 * 0.5 * (1 - 0.5));
            const bool both = p1.more[r1] > threshold1 && p2.more[r2] > threshold2;
            const bool any = p1.more[r1] > threshold1 || p2.more[r2] > threshold2;
            const int a = p1.state[r1], b = p2.state[r2];
            if ((b == 1 || b == -1) && (a == 1 || a == -1)) {
                const int same = (a == b) ? 0 : 1;  // orientation 0 = same sign agrees
                scores[same] += 1;
                scores[1 - same] -= 1;
                if (b == 1 && a == 1) { m11[0]++; m10[1]++; }
                else if (b == 1 && a == -1) { m01[0]++; m00[1]++; }
                else if (b == -1 && a == 1) { m10[0]++; m11[1]++; }
                else { m00[0]++; m01[1]++; }
                if (both) divergent[1 - same]++;
                if (any) unsure[1 - same]++;
            }
        }
        r1++;
        r2++;
    }
    Distance res;
    res.augmented = !((divergent[0] >= threshold_p && divergent[1] >= threshold_p) || (unsure[0] >= 5 && unsure[1] >= 5) ||
                      comparable == 0);
    const int best = scores[1] > scores[0] ? 1 : 0;
    res.n00 = m00[best];
    res.n01 = m01[best];
    res.n10 = m10[best];
    res.n11 = m11[best];
    res.phased = (short)(-2 * best + 1);
    return res;
}

float chi_square(const Distance& d) {
    const int n = d.n00 + d.n01 + d.n10 + d.n11;
    if (n == 0) return 0;
    const float pmax1 = float(d.n10 + d.n11) / n;
    const float pmax2 = float(d.n01 + d.n11) / n;
    if (pmax1 * (1 - pmax1) == 0 && pmax2 * (1 - pmax2) == 0) return -1;
    if (pmax1 * pmax2 * (1 - pmax1) * (1 - pmax2) == 0) return 0;
    const float res = std::pow((d.n00 - (1 - pmax1) * (1 - pmax2) * n), 2) / ((1 - pmax1) * (1 - pmax2) * n) +
                      std::pow((d.n01 - (1 - pmax1) * pmax2 * n), 2) / ((1 - pmax1) * pmax2 * n) +
                      std::pow((d.n10 - pmax1 * (1 - pmax2) * n), 2) / (pmax1 * (1 - pmax2) * n) +
                      std::pow((d.n11 - pmax1 * pmax2 * n), 2) / (pmax1 * pmax2 * n);
    return res;
}

void build_partitions(const std::vector<Column>& suspects, float mean_error, std::vector<Partition>& finals) {
    finals.clear();
    std::vector<Partition> partitions;
    int last_position = -5;
    // loop 1 (:590-638): every suspect column either reinforces the first partition it agrees with, or founds one
    for (const Column& snp : suspects) {
        if (snp.pos - last_position <= 5) continue;
        bool found = false;
        int correlating = 0;
        for (Partition& part : partitions) {
            if (std::abs(snp.pos - part.pos_right) > 50000) continue;
            const Distance dis = distance(part, snp, (char)snp.ref_base, true);
            const int comparable = dis.n00 + dis.n11 + dis.n01 + dis.n10;
            if (dis.n00 + dis.n01 > 0.1 * comparable && dis.n00 + dis.n01 < 0.9 * comparable &&
                dis.n01 + dis.n11 > 0.1 * comparable && dis.n01 + dis.n11 < 0.9 * comparable && chi_square(dis) > 15) {
                correlating += 1;
                part.correlating += 1;
            }
            const bool enough = (size_t)comparable >= snp.readIdxs.size() / 2;
            if ((dis.n01 <= std::max(0.1 * (dis.n00 + dis.n01), 1.0) && dis.n10 < std::max(0.1 * (dis.n11 + dis.n10), 1.0) && enough) ||
                (dis.n00 <= std::max(0.1 * (dis.n00 + dis.n01), 1.0) && dis.n11 < std::max(0.1 * (dis.n11 + dis.n10), 1.0) && enough)) {
                found = true;
                part.augment(dis.phased_column, snp.pos);
                break;
            }
        }
        if (!found) {
            Partition p(snp, snp.pos, snp.ref_base);
            p.correlating = correlating;
            partitions.push_back(std::move(p));
        } else {
            last_position = snp.pos;
        }
    }
    if (partitions.empty()) return;
    // loop 2 (:647-708): keep the significant, informative partitions and merge those that tell the same story
    for (Partition& cand : partitions) {
        const double p_value = cand.significance((int)suspects.size());
        if (!((p_value < 0.001 || cand.correlating > 1) && cand.informative(false, mean_error))) continue;
        bool different = true;
        for (Partition& fin : finals) {
            const Distance dis = distance(fin, cand, 2);
            if (dis.augmented && (dis.n00 + dis.n11 > 5 * (dis.n01 + dis.n10) || dis.n10 + dis.n01 > 5 * (dis.n00 + dis.n11)) &&
                dis.n10 < std::max(2, 2 * dis.n01) && dis.n01 < std::max(2, 2 * dis.n10)) {
                Partition merged = fin;
                merged.merge(cand, dis.phased);
                if (dis.n01 + dis.n10 < 0.1 * (dis.n00 + dis.n11) || merged.confidence_score() > fin.confidence_score()) {
                    fin = std::move(merged);
                    different = false;
                    break;
                }
            }
        }
        if (different) finals.push_back(cand);
    }
}

}  // namespace hs
