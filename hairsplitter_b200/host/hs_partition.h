// Host side of keep_only_robust_variants (reference src/call_variants.cpp:577-768): the two sequential
// loops that BUILD the partitions (greedy aggregation of correlated suspect columns, then merging) stay
// on the host -- every step depends on the partitions the previous one changed -- while the two loops that
// USE them over every column (:721-764, 85 % of the reference's run time) run on the GPU
// (hsgpu_robust_filter). The arithmetic follows the reference expression by expression, including its
// float/double mix, because the thresholds decide which SNPs reach the .col file.
#pragma once
#include <vector>

#include "hs_types.h"

namespace hs {

// class Partition (src/Partition.h:28-84): a sparse +1/-1 vector over the reads of a contig
struct Partition {
    std::vector<int> readIdx;       // ascending neighbour indices
    std::vector<short> state;       // mostFrequentBases: 1, -1, 0 (undecided), -2 (masked)
    std::vector<int> more, less;    // moreFrequence / lessFrequence
    int occurrences = 0;            // numberOfOccurences
    int pos_left = -1, pos_right = -1;
    int correlating = 0;            // number_of_correlating_snps

    Partition() {}
    Partition(const Column& snp, int pos, unsigned char ref_base);   // src/Partition.cpp:34-80
    void augment(const Column& phased, int pos);                     // augmentPartition, :243-397
    void merge(const Partition& other, short phased);                // mergePartition, :401-537
    bool informative(bool last_read_biased, float mean_error) const; // isInformative, :141-179
    float significance(int n_columns) const;                         // isSignificant, :197-233
    float confidence_score() const;                                  // compute_conf, :716-732
};

// distancePartition (src/call_variants.h:46-60)
struct Distance {
    int n00 = 0, n01 = 0, n10 = 0, n11 = 0;
    int solid00 = 0, solid01 = 0, solid10 = 0, solid11 = 0;
    short phased = 1;
    bool augmented = true;
    unsigned char second_base = ' ';
    Column phased_column;  // partition_to_augment: the column as 'A' (ref) / 'a' (alt) / ' '
};

Distance distance(const Partition& part, const Column& col, char ref_base, bool want_column);  // :778-967
Distance distance(const Partition& a, const Partition& b, int threshold_p);                    // :977-1127
float chi_square(const Distance& d);                                                            // :1135-1163

// loops 1 and 2 of keep_only_robust_variants: suspect columns in, final partitions out
void build_partitions(const std::vector<Column>& suspects, float mean_error, std::vector<Partition>& finals);

}  // namespace hs
