// Parsers and writers of the HS_call_variants drop-in. Same observable behaviour as the reference's
// parse_reads / parse_assembly / parse_SAM / parse_reads_on_contig (src/input_output.cpp:39-569) and
// output_files (src/call_variants.cpp:1174-1213), restated over the flat Store.
#pragma once
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "hs_types.h"

namespace hs {

// a whole file mapped read-only (ingestion through one mmap instead of seekg/getline per read,
// src/input_output.cpp:546-569; the SAM text is parsed in place)
struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    bool open(const std::string& path);  // false when the file cannot be opened; an empty file maps to size 0
    void close();
    ~MappedFile() { close(); }
};

void parse_reads(const std::string& path, Store& st);
void parse_assembly(const std::string& path, Store& st);
void parse_sam(const std::string& path, Store& st, bool amplicon);
// ASCII bases -> the 2-bit words of hsgpu_pileup_input (16 bases per word, base i in bits 2(i % 16)), in the reference's
// own Sequence code: A 0, C 1, G 2, anything else 3 (src/sequence.cpp:13-23). Same result as hsgpu_pack_bases_ascii, which
// takes a branch per base; this one goes through a table, 16 bases per step (the reads of a 60 Mb metagenome at 100x are
// 6 GB of text).
void pack_bases_2bit(const char* seq, int64_t n, uint32_t* out_words);
// SAM CIGAR string -> BAM ops (len<<4 | index in "MIDNSHP=X"); "*" gives no ops. Exits like
// convert_cigar (src/tools.cpp:27-57) when a length is missing.
void cigar_ops(const std::string& cigar, std::vector<uint32_t>& ops);
// the sequence lines of the reads aligned on `contig`, in neighbour order (parse_reads_on_contig)
void load_read_sequences(std::ifstream& reads_file, const Store& st, int64_t contig, std::vector<std::string>& out);
// the same lines as views into the mapped reads file (no copies): (pointer, length) per aligned read
void view_read_sequences(const MappedFile& reads, const Store& st, int64_t contig,
                         std::vector<std::pair<const char*, size_t>>& out);
// .col and .vcf, contigs in the iteration order of the reference's own container
void write_outputs(const Store& st, const std::unordered_map<int, std::vector<Column>>& variants,
                   const std::string& col_file, const std::string& vcf_file);

}  // namespace hs
