// TEST INFRASTRUCTURE ONLY (oracle/): thin C-ABI shim around the UNMODIFIED reference
// (RolandFaure/Hairsplitter, sources compiled where they lie under /root/reference/src).
// It lets the Python tests call the reference's own generate_msa / call_variants /
// keep_only_robust_variants / distance / computeChiSquare on flat arrays, so that both the plain-C
// restatement (hs_oracle.c) and the CUDA path (libhsgpu.so) are pinned against the real thing.
// Nothing in the product path links or loads this file.
//
// Reference entry points wrapped here:
//   generate_msa                src/call_variants.cpp:50
//   call_variants               src/call_variants.cpp:447
//   keep_only_robust_variants   src/call_variants.cpp:577
//   distance(Partition,Column)  src/call_variants.cpp:778
//   computeChiSquare            src/call_variants.cpp:1135
//   robin_hood::unordered_map   src/robin_hood.h (iteration order used for allele tie-breaks)
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <algorithm>

#include "call_variants.h"

namespace {
struct RefCV {
    std::vector<Read> allreads;
    std::vector<Overlap> allOverlaps;
    std::vector<Column> snps;  // the pileup ("msa")
    std::string newref;
    float meanDistance = 0;
    std::vector<Column> suspects;
    std::vector<Column> automatic;
    std::vector<Column> filtered;
    std::vector<Column> merged;
    std::vector<Partition> parts;
    std::unordered_map<int, std::vector<std::pair<int, int>>> readLimits;
};
}  // namespace

extern "C" {

// Builds allreads/allOverlaps for one contig (index 0) with n_reads aligned reads and runs the
// reference's generate_msa. read_seqs are in ORIGINAL read orientation (as in the FASTA); the
// CIGAR is the SAM one (applied by the reference to the reverse complement when strand == 0).
void* hsref_cv_create(const char* contig_seq, int n_reads, const char* const* read_seqs,
                      const char* const* cigars, const int* contig_start, const unsigned char* strand) {
    RefCV* h = new RefCV();
    std::string cs(contig_seq);
    Read contig(cs, cs.size());
    contig.name = "contig";
    h->allreads.push_back(contig);
    for (int i = 0; i < n_reads; i++) {
        std::string rs(read_seqs[i]);
        Read r(rs, rs.size());
        r.name = "r" + std::to_string(i);
        h->allreads.push_back(r);
        Overlap o;
        o.sequence1 = 1 + i;
        o.sequence2 = 0;
        o.position_1_1 = 0;
        o.position_1_2 = (int)rs.size();
        o.position_2_1 = contig_start[i];
        o.position_2_2 = contig_start[i];
        o.strand = strand[i] != 0;
        o.diff = 0;
        o.CIGAR = cigars[i];
        h->allreads[0].add_overlap(h->allOverlaps.size());
        h->allOverlaps.push_back(o);
    }
    robin_hood::unordered_map<int, int> insertionPositions;
    std::string tmp = "/tmp/";
    h->meanDistance = generate_msa(0, h->allOverlaps, h->allreads, h->snps, insertionPositions, 0,
                                   h->readLimits, h->newref, tmp, false);
    return h;
}

void hsref_cv_destroy(void* hv) { delete (RefCV*)hv; }

float hsref_cv_mean_distance(void* hv) { return ((RefCV*)hv)->meanDistance; }

long hsref_cv_n_columns(void* hv) { return (long)((RefCV*)hv)->snps.size(); }

long hsref_cv_n_cells(void* hv) {
    long n = 0;
    for (auto& c : ((RefCV*)hv)->snps) n += (long)c.content.size();
    return n;
}

// CSR dump of the whole pileup, columns in order, cells in the reference's order.
void hsref_cv_get_pileup(void* hv, long* col_off, unsigned int* read_idx, unsigned char* code) {
    RefCV* h = (RefCV*)hv;
    long n = 0;
    for (size_t c = 0; c < h->snps.size(); c++) {
        col_off[c] = n;
        for (size_t i = 0; i < h->snps[c].content.size(); i++) {
            read_idx[n] = h->snps[c].readIdxs[i];
            code[n] = h->snps[c].content[i];
            n++;
        }
    }
    col_off[h->snps.size()] = n;
}

void hsref_cv_get_newref(void* hv, unsigned char* out) {
    RefCV* h = (RefCV*)hv;
    memcpy(out, h->newref.data(), h->newref.size());
}

// end column (positionOfReads[n].second) of each read, from readLimits
void hsref_cv_get_read_ends(void* hv, int* start, int* end) {
    RefCV* h = (RefCV*)hv;
    auto& v = h->readLimits[0];
    for (size_t i = 0; i < v.size(); i++) {
        start[i] = v[i].first;
        end[i] = v[i].second;
    }
}

// Runs the reference's call_variants. mean_error < 0 means "use generate_msa's value".
int hsref_cv_call_variants(void* hv, float mean_error, float auto_threshold) {
    RefCV* h = (RefCV*)hv;
    std::vector<size_t> suspectPositions;
    std::string tmp = "/tmp/";
    float me = mean_error < 0 ? h->meanDistance : mean_error;
    h->automatic.clear();
    h->suspects = call_variants(h->snps, h->allreads, h->allOverlaps, 0, h->newref, suspectPositions, me,
                                auto_threshold, h->automatic, tmp, false);
    return (int)h->suspects.size();
}

float hsref_cv_depth(void* hv) { return ((RefCV*)hv)->allreads[0].depth; }

void hsref_cv_get_column_bases(void* hv, unsigned char* ref_base, unsigned char* second_base) {
    RefCV* h = (RefCV*)hv;
    for (size_t c = 0; c < h->snps.size(); c++) {
        ref_base[c] = h->snps[c].ref_base;
        second_base[c] = h->snps[c].second_base;
    }
}

static std::vector<Column>& pick(RefCV* h, int which) {
    switch (which) {
        case 0: return h->suspects;
        case 1: return h->automatic;
        case 2: return h->filtered;
        default: return h->merged;
    }
}

int hsref_cv_list_size(void* hv, int which) { return (int)pick((RefCV*)hv, which).size(); }

void hsref_cv_list_get(void* hv, int which, int* pos, unsigned char* ref_base, unsigned char* second_base) {
    auto& v = pick((RefCV*)hv, which);
    for (size_t i = 0; i < v.size(); i++) {
        pos[i] = v[i].pos;
        ref_base[i] = v[i].ref_base;
        second_base[i] = v[i].second_base;
    }
}

// Runs keep_only_robust_variants and the merge with the automatic SNPs (src/call_variants.cpp:1335-1352).
int hsref_cv_robust(void* hv, float mean_error) {
    RefCV* h = (RefCV*)hv;
    float me = mean_error < 0 ? h->meanDistance : mean_error;
    keep_only_robust_variants(h->snps, h->suspects, h->filtered, me, h->parts);
    h->merged.clear();
    size_t ia = 0, jf = 0;
    while (ia < h->automatic.size() && jf < h->filtered.size()) {
        if (h->automatic[ia].pos < h->filtered[jf].pos) {
            h->merged.push_back(h->automatic[ia]);
            ia++;
        } else if (h->automatic[ia].pos > h->filtered[jf].pos) {
            h->merged.push_back(h->filtered[jf]);
            jf++;
        } else {
            h->merged.push_back(h->automatic[ia]);
            ia++;
            jf++;
        }
    }
    return (int)h->parts.size();
}

int hsref_cv_part_size(void* hv, int p) { return (int)((RefCV*)hv)->parts[p].getReads().size(); }

void hsref_cv_part_get(void* hv, int p, int* read_idx, short* state, int* more, int* less, int* left_right) {
    Partition& P = ((RefCV*)hv)->parts[p];
    auto r = P.getReads();
    auto s = P.getPartition();
    auto m = P.getMore();
    auto l = P.getLess();
    for (size_t i = 0; i < r.size(); i++) {
        read_idx[i] = r[i];
        state[i] = s[i];
        more[i] = m[i];
        less[i] = l[i];
    }
    left_right[0] = P.get_left();
    left_right[1] = P.get_right();
}

// distance(parts[p], msa[col], ref_base). out = n00,n01,n10,n11,solid00,solid01,solid10,solid11,secondBase,augmented
void hsref_cv_distance(void* hv, int p, int col, int ref_base, int* out) {
    RefCV* h = (RefCV*)hv;
    distancePartition d = distance(h->parts[p], h->snps[col], (char)ref_base);
    out[0] = d.n00;
    out[1] = d.n01;
    out[2] = d.n10;
    out[3] = d.n11;
    out[4] = d.solid00;
    out[5] = d.solid01;
    out[6] = d.solid10;
    out[7] = d.solid11;
    out[8] = d.augmented ? (unsigned char)d.secondBase : 0;
    out[9] = d.augmented ? 1 : 0;
}

// distance() of an arbitrary caller-supplied partition against a pileup column.
void hsref_cv_distance_custom(void* hv, int n, const int* read_idx, const short* state, const int* more,
                              const int* less, int col, int ref_base, int* out) {
    RefCV* h = (RefCV*)hv;
    Partition P;
    std::vector<short> s(state, state + n);
    std::vector<int> r(read_idx, read_idx + n), m(more, more + n), l(less, less + n);
    P.new_corrected_partition(s, r, m, l);
    distancePartition d = distance(P, h->snps[col], (char)ref_base);
    out[0] = d.n00;
    out[1] = d.n01;
    out[2] = d.n10;
    out[3] = d.n11;
    out[4] = d.solid00;
    out[5] = d.solid01;
    out[6] = d.solid10;
    out[7] = d.solid11;
    out[8] = d.augmented ? (unsigned char)d.secondBase : 0;
    out[9] = d.augmented ? 1 : 0;
}

float hsref_chi_square(int n00, int n01, int n10, int n11) {
    distancePartition d;
    d.n00 = n00;
    d.n01 = n01;
    d.n10 = n10;
    d.n11 = n11;
    return computeChiSquare(d);
}

// Iteration order of robin_hood::unordered_map<unsigned char,int> after inserting `keys` in order
// with the find()/operator[] idiom of call_variants.cpp:481-494. Returns the number of entries.
int hsref_rh_order(const unsigned char* keys, int n, unsigned char* out) {
    robin_hood::unordered_map<unsigned char, int> content;
    for (int i = 0; i < n; i++) {
        if (content.find(keys[i]) == content.end()) content[keys[i]] = 0;
        content[keys[i]] += 1;
    }
    int k = 0;
    for (auto it = content.begin(); it != content.end(); it++) out[k++] = it->first;
    return k;
}

// std::sort with the comparator of call_variants.cpp:501 (libstdc++ introsort, unstable).
void hsref_sort_desc(unsigned char* keys, int* counts, int n) {
    std::vector<std::pair<unsigned char, int>> v;
    for (int i = 0; i < n; i++) v.push_back(std::make_pair(keys[i], counts[i]));
    std::sort(v.begin(), v.end(),
              [](const std::pair<unsigned char, int>& a, const std::pair<unsigned char, int>& b) {
                  return a.second > b.second;
              });
    for (int i = 0; i < n; i++) {
        keys[i] = v[i].first;
        counts[i] = v[i].second;
    }
}

}  // extern "C"
