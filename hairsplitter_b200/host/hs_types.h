// Host-side data model of the HS_call_variants drop-in (hairsplitter_b200/host). The reference keeps
// reads, contigs and alignments as `Read` / `Overlap` objects (src/read.h:12-77) and pileup columns
// as `Column` (src/Partition.h:8-14); these are the flat equivalents the GPU path is fed from.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace hs {

// one SAM record that survived parse_SAM's filters (src/input_output.cpp:274-536)
struct Alignment {
    int64_t read;      // index in Store::seqs (Overlap.sequence1)
    int64_t contig;    // index in Store::seqs (Overlap.sequence2)
    int pos_1_1, pos_1_2, pos_2_1, pos_2_2;  // Overlap.position_*
    bool strand;
    std::string cigar;
};

// a read or a contig (src/read.h:36-74): reads keep only their length and where their sequence line
// starts in the reads file; contigs keep their sequence
struct SeqRec {
    std::string name;
    int64_t length = 0;          // Read::size_ as given to the constructor
    int64_t file_pos = -1;       // Read::positionInFile_
    std::string sequence;        // contigs only (S line field 3)
    float depth = -1;            // Read::depth
    std::vector<int64_t> alns;   // Read::neighbors_: indices into Store::alns, in SAM order
};

struct Store {
    std::vector<SeqRec> seqs;                              // allreads
    std::unordered_map<std::string, int64_t> index;        // indices
    std::vector<int64_t> contigs;                          // backbone_reads
    std::vector<Alignment> alns;                           // allOverlaps
};

// a pileup column as the reference passes it around (src/Partition.h:8-14)
struct Column {
    int pos = 0;
    std::vector<uint32_t> readIdxs;
    std::vector<uint8_t> content;
    uint8_t ref_base = 0, second_base = 0;
};

}  // namespace hs
