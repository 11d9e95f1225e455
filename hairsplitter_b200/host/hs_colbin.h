// Binary sidecar of the .col file. The .col text stays the interface between HS_call_variants and
// HS_separate_reads (output_files, src/call_variants.cpp:1174-1213; parse_column_file, src/separate_reads.cpp:46-190);
// next to it the writer leaves "<col>.hsb" with the same content in the form the second executable needs: per contig
// the CONTIG / READ lines verbatim (they are short and parse_column_file's own code reads them) and the SNPS lines as
// flat arrays (positions, ref / second codes, CSR offsets, read indices, codes) -- no decimal text to tokenise for the
// 2 x depth numbers of every SNP. The reader applies parse_column_file's per-SNP filters (max_coverage, :157; rarest
// strain abundance, :167) to the arrays exactly as the text parser applies them to the tokens, so both routes leave the
// same ColContig structures. A sidecar is used only if it names the .col file it lies next to: byte size, modification
// time and a hash of the file's first and last 64 KiB must agree; anything else (no sidecar, a .col written by the
// reference's own executable, a copied or edited file, a truncated sidecar) falls back to the text.
// HS_SIDECAR=0 in the environment: neither written nor read.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "hs_sepreads.h"

namespace hs {

struct ColSidecarBlock {
    std::string head;                          // the CONTIG line and the READ lines of the block, as in the .col
    const std::vector<Column>* snps = nullptr;  // the SNPS lines, in file order
};

bool col_sidecar_enabled();
std::string col_sidecar_path(const std::string& col_path);

// writes <col_path>.hsb for the .col file that has just been written and closed; blocks in file order.
// Returns false (and leaves no sidecar) on any failure: the sidecar is an accelerator, never a requirement.
bool write_col_sidecar(const std::string& col_path, const std::vector<ColSidecarBlock>& blocks);

// fills `contigs` like parse_column_file from <col_path>.hsb; false = no usable sidecar (contigs untouched)
bool read_col_sidecar(const std::string& col_path, std::vector<ColContig>& contigs, int max_coverage,
                      float rarest_strain_abundance);

// CONTIG / READ lines of one block through parse_column_file's own line code (hs_sepreads.cpp)
void parse_col_head(const char* b, const char* e, ColContig& c);

}  // namespace hs
