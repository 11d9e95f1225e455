// Reference-side glue for HS_call_variants: what a HairSplitter maintainer adds to bind libhsgpu (INTEGRATION.md
// sections 0-3), as code that is compiled against the reference's own headers and run under the reference's own,
// unmodified main().
//
// This file defines generate_msa, call_variants and keep_only_robust_variants with exactly the reference's signatures
// (src/call_variants.h:12-44). oracle/Makefile builds oracle/_ref/HS_call_variants_glued from it: the reference's
// objects live in a shared library (libhsref_cv.so: main() renamed at compile time, everything else untouched), whose
// calls to the three functions go through the PLT and therefore bind to the definitions below. The reference keeps
// doing everything else: argv, parse_reads / parse_assembly / parse_SAM / parse_reads_on_contig, the OpenMP loop over
// contigs, the merge with the automatic SNPs, output_files. tests/test_gpu_callvariants.py runs it next to the
// reference executable: the .col, .vcf and error-rate files must be byte-identical.
//
// Only the C ABI of include/hsgpu.h is used for the GPU side (one context per OpenMP thread).
// Loops 1-2 of keep_only_robust_variants (sequential, mutating) run through this repository's host library
// (hairsplitter_b200/host/hs_partition.h), as they do in bin/HS_call_variants.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>

#include "call_variants.h"  // the reference's (src/), found through -I

#include "hsgpu.h"
#include "hs_partition.h"

namespace {

// what one OpenMP thread carries from generate_msa to keep_only_robust_variants of the same contig
struct Session {
    hsgpu_ctx* ctx = nullptr;
    hsgpu_pileup* pu = nullptr;
    int32_t L = 0;
    std::vector<uint8_t> ref_base, second_base;
    std::vector<int32_t> suspect_pos;
};
thread_local Session g_session;

[[noreturn]] void die(hsgpu_ctx* ctx, const char* what, int rc) {
    std::cout << "ERROR: " << what << " failed (" << rc << "): " << hsgpu_last_error(ctx) << std::endl;
    std::exit(1);  // the reference's own error convention (cout + exit(1)); there is no CPU fallback
}
#define GPU(call)                                            \
    do {                                                     \
        const int rc_ = (call);                              \
        if (rc_ != HSGPU_OK) die(g_session.ctx, #call, rc_); \
    } while (0)

hsgpu_ctx* context() {
    Session& s = g_session;
    if (!s.ctx) {
        int device = 0;
        if (const char* e = std::getenv("HSGPU_DEVICE")) device = std::atoi(e);
        const int rc = hsgpu_ctx_create(device, &s.ctx);
        if (rc != HSGPU_OK) die(nullptr, "hsgpu_ctx_create", rc);
    }
    return s.ctx;
}

// full columns of the given positions in the reference's own layout
void fetch_columns(const std::vector<int32_t>& pos, std::vector<Column>& out) {
    Session& s = g_session;
    out.clear();
    const int n = (int)pos.size();
    if (n == 0) return;
    std::vector<int64_t> off((size_t)n + 1, 0);
    int rc = hsgpu_pileup_extract_columns(s.pu, 0, n, pos.data(), 0, off.data(), nullptr, nullptr);  // sizes first
    if (rc != HSGPU_OK && rc != HSGPU_ERR_CAPACITY) die(s.ctx, "hsgpu_pileup_extract_columns", rc);
    std::vector<uint32_t> idx((size_t)std::max<int64_t>(off[n], 1));
    std::vector<uint8_t> code((size_t)std::max<int64_t>(off[n], 1));
    GPU(hsgpu_pileup_extract_columns(s.pu, 0, n, pos.data(), off[n], off.data(), idx.data(), code.data()));
    out.resize((size_t)n);
    for (int i = 0; i < n; i++) {
        Column& c = out[(size_t)i];
        c.pos = pos[(size_t)i];
        c.ref_base = s.ref_base[(size_t)pos[(size_t)i]];
        c.second_base = s.second_base[(size_t)pos[(size_t)i]];
        c.readIdxs.assign(idx.begin() + off[i], idx.begin() + off[i + 1]);
        c.content.assign(code.begin() + off[i], code.begin() + off[i + 1]);
    }
}

}  // namespace

// src/call_variants.cpp:50-437
float generate_msa(long int bbcontig, std::vector<Overlap>& allOverlaps, std::vector<Read>& allreads, std::vector<Column>& snps,
                   robin_hood::unordered_map<int, int>& /*insertionPos*/, int backboneReadIndex,
                   std::unordered_map<int, std::vector<std::pair<int, int>>>& readLimits, std::string& newref,
                   std::string& /*tmpFolder*/, bool /*DEBUG*/) {
    Session& s = g_session;
    hsgpu_ctx* ctx = context();
    if (s.pu) {
        hsgpu_pileup_destroy(s.pu);
        s.pu = nullptr;
    }
    Read& contig = allreads[bbcontig];
    const size_t n_neighbors = contig.neighbors_.size();
    contig.new_backbone(std::make_pair(backboneReadIndex, (int)n_neighbors), n_neighbors + 1);  // :71
    const std::string consensus = contig.sequence_.str();
    const int32_t L = (int32_t)consensus.size();
    s.L = L;

    // pack once per contig: 2-bit bases in the reference's own Sequence code, reads in ORIGINAL orientation (the
    // strand is applied on the device), CIGAR strings as BAM ops
    std::vector<uint32_t> contig_words(((size_t)L + 15) / 16 + 1, 0), read_words, cigar_ops;
    std::vector<int64_t> read_word_off(1, 0), cigar_off(1, 0);
    std::vector<int32_t> read_len, read_start;
    std::vector<uint8_t> strand;
    std::vector<std::pair<int, int>> positionOfReads;
    hsgpu_pack_bases_ascii(consensus.data(), L, contig_words.data());
    for (size_t n = 0; n < n_neighbors; n++) {
        const Overlap& o = allOverlaps[contig.neighbors_[n]];
        if (o.CIGAR == "" || (long int)o.sequence2 != bbcontig) {
            // the PAF branches of the reference (:95-147) need an aligner; HS_call_variants is fed a SAM file
            std::cout << "ERROR: hsgpu glue: only alignments with a CIGAR on the contig (SAM input) are supported" << std::endl;
            std::exit(1);
        }
        allreads[o.sequence1].new_backbone(std::make_pair(backboneReadIndex, (int)n), n_neighbors + 1);  // :105
        const std::string seq = allreads[o.sequence1].sequence_.str();
        const size_t w0 = read_words.size();
        read_words.resize(w0 + (seq.size() + 15) / 16, 0);
        hsgpu_pack_bases_ascii(seq.data(), (int64_t)seq.size(), read_words.data() + w0);
        read_word_off.push_back((int64_t)read_words.size());
        const size_t c0 = cigar_ops.size();
        cigar_ops.resize(c0 + o.CIGAR.size() + 1);
        const int64_t n_ops = hsgpu_parse_cigar(o.CIGAR.c_str(), cigar_ops.data() + c0, (int64_t)o.CIGAR.size() + 1);
        if (n_ops < 0) die(ctx, "hsgpu_parse_cigar", (int)n_ops);
        cigar_ops.resize(c0 + (size_t)n_ops);
        cigar_off.push_back((int64_t)cigar_ops.size());
        read_len.push_back((int32_t)seq.size());
        read_start.push_back(o.position_2_1);
        strand.push_back(o.strand ? 1 : 0);
        positionOfReads.push_back(std::make_pair(o.position_2_1, o.position_2_2));
    }
    if (read_words.empty()) read_words.push_back(0);
    if (cigar_ops.empty()) cigar_ops.push_back(0);
    const int64_t contig_word_off[2] = {0, ((int64_t)L + 15) / 16};
    const int64_t contig_read_off[2] = {0, (int64_t)read_len.size()};
    hsgpu_pileup_input in;
    std::memset(&in, 0, sizeof(in));
    in.n_contigs = 1;
    in.contig_len = &L;
    in.contig_bases = contig_words.data();
    in.contig_word_off = contig_word_off;
    in.contig_read_off = contig_read_off;
    in.n_reads = (int64_t)read_len.size();
    in.read_bases = read_words.data();
    in.read_word_off = read_word_off.data();
    in.read_len = read_len.data();
    in.cigar = cigar_ops.data();
    in.cigar_off = cigar_off.data();
    in.read_start = read_start.data();
    in.read_strand = strand.data();
    GPU(hsgpu_pileup_create(ctx, &in, &s.pu));
    GPU(hsgpu_pileup_build(s.pu));

    // the out-parameters main() hands on: the columns stay on the device (the two functions below fetch the ones they
    // need), so `snps` only carries the positions; readLimits (:354,363) and newref (:366-376) are filled as the
    // reference fills them
    snps = std::vector<Column>((size_t)L);
    for (int32_t c = 0; c < L; c++) snps[(size_t)c].pos = c;
    if (!positionOfReads.empty()) {
        std::vector<int32_t> ends(positionOfReads.size());
        GPU(hsgpu_pileup_read_ends(s.pu, ends.data()));
        for (size_t n = 0; n < ends.size(); n++) positionOfReads[n].second = ends[n];
        readLimits[(int)bbcontig] = positionOfReads;
    }
    {
        const std::string acgt = "ACGT-";
        unsigned char b2 = 'C', b1 = 'G';  // after the first shift: 'A','C','G' -> (b-2, b-1) = ('C','G')
        newref.clear();
        newref.reserve((size_t)L);
        for (const char base : consensus) {
            newref += (char)(unsigned char)('!' + 5 * acgt.find((char)b2) + acgt.find((char)b1) + 25 * acgt.find(base));
            b2 = b1;
            b1 = (unsigned char)base;
        }
    }
    int64_t cells = 0, dist = 0, alen = 0;
    GPU(hsgpu_pileup_stats(s.pu, &cells, &dist, &alen));
    return hsgpu_mean_distance(dist, alen);  // :434
}

// src/call_variants.cpp:447-567
std::vector<Column> call_variants(std::vector<Column>& snps, std::vector<Read>& allreads, std::vector<Overlap>& /*allOverlaps*/,
                                  long int contig, std::string& ref, std::vector<size_t>& /*suspectPostitions*/, float& meanError,
                                  float automatic_snp_threshold, std::vector<Column>& automatic_snps, std::string& /*tmpFolder*/,
                                  bool /*DEBUG*/) {
    Session& s = g_session;
    if (!s.pu) {
        std::cout << "ERROR: hsgpu glue: call_variants without the pileup of generate_msa" << std::endl;
        std::exit(1);
    }
    GPU(hsgpu_column_rank(s.pu, &meanError, automatic_snp_threshold));
    int32_t n_suspects = 0;
    int64_t depth_sum = 0;
    GPU(hsgpu_column_counts(s.pu, &n_suspects, &depth_sum));
    allreads[contig].depth = (double)depth_sum / ref.size();  // :565
    // ref_base / second_base of every column (:503-507): keep_only_robust_variants reads them from the msa
    s.ref_base.assign((size_t)std::max(s.L, 1), 0);
    s.second_base.assign((size_t)std::max(s.L, 1), 0);
    GPU(hsgpu_column_summary(s.pu, 0, s.ref_base.data(), s.second_base.data(), nullptr, nullptr));
    for (int32_t c = 0; c < s.L && (size_t)c < snps.size(); c++) {
        snps[(size_t)c].ref_base = s.ref_base[(size_t)c];
        snps[(size_t)c].second_base = s.second_base[(size_t)c];
    }
    s.suspect_pos.assign((size_t)std::max(n_suspects, 1), 0);
    std::vector<uint8_t> is_auto((size_t)std::max(n_suspects, 1), 0);
    GPU(hsgpu_suspects(s.pu, 0, n_suspects, s.suspect_pos.data(), is_auto.data()));
    s.suspect_pos.resize((size_t)n_suspects);
    std::vector<Column> suspicious;
    fetch_columns(s.suspect_pos, suspicious);
    for (int32_t i = 0; i < n_suspects; i++)
        if (is_auto[(size_t)i]) automatic_snps.push_back(suspicious[(size_t)i]);  // :531-533
    return suspicious;
}

// src/call_variants.cpp:577-768
void keep_only_robust_variants(std::vector<Column>& msa, std::vector<Column>& snps_in, std::vector<Column>& snps_out,
                               float mean_error, std::vector<Partition>& parts) {
    Session& s = g_session;
    snps_out = std::vector<Column>();
    if (!s.pu) {
        std::cout << "ERROR: hsgpu glue: keep_only_robust_variants without the pileup of generate_msa" << std::endl;
        std::exit(1);
    }
    // loops 1-2 (:590-708) on the host, over the suspect columns main() hands in
    std::vector<hs::Column> suspects(snps_in.size());
    std::vector<int32_t> suspect_pos(snps_in.size());
    for (size_t i = 0; i < snps_in.size(); i++) {
        suspects[i].pos = snps_in[i].pos;
        suspects[i].ref_base = snps_in[i].ref_base;
        suspects[i].second_base = snps_in[i].second_base;
        suspects[i].readIdxs = snps_in[i].readIdxs;
        suspects[i].content = snps_in[i].content;
        suspect_pos[i] = snps_in[i].pos;
    }
    std::vector<hs::Partition> finals;
    hs::build_partitions(suspects, mean_error, finals);
    if (!finals.empty()) {  // :640-642: no partition, nothing is kept
        // loops 3-4 (:721-764) in one call
        std::vector<int64_t> part_off(1, 0);
        std::vector<int32_t> p_idx, p_more, p_less;
        std::vector<int16_t> p_state;
        for (const hs::Partition& p : finals) {
            p_idx.insert(p_idx.end(), p.readIdx.begin(), p.readIdx.end());
            p_state.insert(p_state.end(), p.state.begin(), p.state.end());
            p_more.insert(p_more.end(), p.more.begin(), p.more.end());
            p_less.insert(p_less.end(), p.less.begin(), p.less.end());
            part_off.push_back((int64_t)p_idx.size());
        }
        hsgpu_partitions P;
        std::memset(&P, 0, sizeof(P));
        P.n_parts = (int32_t)finals.size();
        P.part_off = part_off.data();
        P.read_idx = p_idx.data();
        P.state = p_state.data();
        P.more = p_more.data();
        P.less = p_less.data();
        std::vector<int32_t> kept((size_t)std::max<size_t>(msa.size(), 1));
        int32_t n_kept = 0;
        GPU(hsgpu_robust_filter(s.pu, 0, &P, (int32_t)suspect_pos.size(), suspect_pos.data(), (int32_t)kept.size(), kept.data(),
                                &n_kept));
        kept.resize((size_t)n_kept);
        fetch_columns(kept, snps_out);
        // the final partitions, as far as the reference's class lets a caller set them (main() does not read them)
        parts.clear();
        for (const hs::Partition& p : finals) {
            Partition q;
            q.new_corrected_partition(p.state, p.readIdx, p.more, p.less);
            q.number_of_correlating_snps = p.correlating;
            parts.push_back(q);
        }
    }
    hsgpu_pileup_destroy(s.pu);
    s.pu = nullptr;
}

int hs_ref_call_variants_main(int argc, char* argv[]);  // the reference's main(), renamed when libhsref_cv.so is built

int main(int argc, char* argv[]) { return hs_ref_call_variants_main(argc, argv); }
