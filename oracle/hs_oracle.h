/* TEST INFRASTRUCTURE ONLY -- never linked, loaded or executed by the product path.
 *
 * Plain-C CPU restatement of the reference's hot path (RolandFaure/Hairsplitter v1.9.22), used as
 * the parity oracle for the CUDA library (libhsgpu.so). Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load liboracle.so.
 *
 * Parity status: PINNED. Every function here is checked against the unmodified reference compiled
 * from /root/reference (oracle/_ref/libhsref_*.so, built by oracle/Makefile) in tests/test_oracle_*.py,
 * and against the committed fixtures under tests/golden/ generated from that reference.
 */
#ifndef HS_ORACLE_H
#define HS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* generate_msa, src/call_variants.cpp:50-437. Bases are u8 codes A=0 C=1 G=2 T=3 (sequence.cpp:13-23),
 * reads in ORIGINAL orientation; cigar ops are len<<4|op, op = index in "MIDNSHP=X".
 * Output is the column-major pileup exactly as the reference holds it: for column q the cells
 * [col_off[q], col_off[q+1]) with ascending read index and the 3-mer code (33..157).
 * Pass cell_capacity = 0 to only count: returns the number of cells (col_off is still filled).
 * stats[0] = totalDistance numerator (mismatch + I + D), stats[1] = alignment length (without the
 * initial 1), both as exact integers; read_end[n] = positionOfReads[n].second. */
int64_t hso_pileup(const uint8_t* contig, int32_t L, int32_t n_reads, const uint8_t* read_bases,
                   const int64_t* read_off, const uint32_t* cigar, const int64_t* cigar_off,
                   const int32_t* start, const uint8_t* strand, int64_t cell_capacity, int64_t* col_off,
                   uint32_t* read_idx, uint8_t* code, int64_t* stats, int32_t* read_end);

/* float generate_msa() return value from the integer sums (float accumulator semantics, :67-68,434). */
float hso_mean_distance(int64_t distance_sum, int64_t aligned_sum);

/* newref of generate_msa (:366-376): 3-mer code of every contig position. */
void hso_ref_codes(const uint8_t* contig, int32_t L, uint8_t* out);

/* Iteration order of robin_hood::unordered_map<unsigned char,int> (src/robin_hood.h, flat table,
 * 80 % load, murmur-style hash) after inserting the DISTINCT keys `keys[0..n)` in that order.
 * Returns n; out receives the keys in iteration order. */
int hso_rh_order(const uint8_t* keys, int n, uint8_t* out);

/* libstdc++ std::sort (introsort + final insertion sort) with comparator "count descending",
 * restated; (keys, counts) are permuted in place exactly as std::sort would permute the pairs. */
void hso_sort_desc(uint8_t* keys, int32_t* counts, int n);

/* Per-column ranking of call_variants (:477-507): out[0]=k0 (ref_base), out[1]=k1 (second_base),
 * out[2..4] = c0,c1,c2. codes = the column's cell codes in the reference's order. */
void hso_column_rank(const uint8_t* codes, int n, int32_t* out);

/* call_variants (:447-567) over a CSR pileup. Fills ref_base/second_base for every column, the
 * suspect list (positions, ascending) and the "automatic" subset flags. Returns the number of suspects.
 * depth_sum = number of cells counted (depthOfCoverage numerator). */
int32_t hso_call_variants(const int64_t* col_off, const uint8_t* code, int32_t L, float mean_error,
                          float auto_threshold, uint8_t* ref_base, uint8_t* second_base,
                          int32_t* suspect_pos, uint8_t* suspect_is_auto, int32_t suspect_capacity,
                          int64_t* depth_sum);

/* distance(Partition&, Column&, char), src/call_variants.cpp:778-967.
 * Partition as parallel arrays (sorted read_idx, state in {1,-1,0,-2}, more, less); column as
 * (ascending read_idx, code). out = n00,n01,n10,n11,solid00,solid01,solid10,solid11,secondBase,augmented
 * (secondBase reported as 0 when not augmented). */
void hso_distance(int32_t np, const int32_t* p_idx, const int16_t* p_state, const int32_t* p_more,
                  const int32_t* p_less, int32_t nc, const uint32_t* c_idx, const uint8_t* c_code,
                  int32_t ref_base, int32_t* out);

/* computeChiSquare, src/call_variants.cpp:1135-1163 (float/double mix kept verbatim). */
float hso_chi_square(int32_t n00, int32_t n01, int32_t n10, int32_t n11);

/* Rescue predicate of keep_only_robust_variants loop 4 (:751-752): central base differs and not an
 * indel next to a homopolymer. */
int hso_rescue_prefilter(int32_t ref_base, int32_t second_base);

/* Loops 3+4 of keep_only_robust_variants (:718-764) for given final partitions (concatenated arrays,
 * partition p = [part_off[p], part_off[p+1])) and suspect positions; kept receives snps_out positions. */
int32_t hso_robust_filter(int32_t L, const int64_t* col_off, const uint32_t* read_idx, const uint8_t* code,
                          const uint8_t* ref_base, const uint8_t* second_base, int32_t n_parts,
                          const int64_t* part_off, const int32_t* p_idx, const int16_t* p_state,
                          const int32_t* p_more, const int32_t* p_less, int32_t n_suspects,
                          const int32_t* suspect_pos, int32_t* kept);

/* list_similarities_and_differences_between_reads3, src/separate_reads.cpp:374-433, dense output.
 * SNP columns given as CSR (snp_off, read_idx, code) with per-SNP ref_base/second_base.
 * sim/diff are n_reads x n_reads int32, row-major (symmetric, zero diagonal). */
void hso_read_pair_counts(int32_t n_reads, int32_t n_snps, const int64_t* snp_off, const uint32_t* read_idx,
                          const uint8_t* code, const uint8_t* ref_base, const uint8_t* second_base,
                          int32_t* sim, int32_t* diff);

/* Read clipping of modify_GFA (src/create_new_contigs.cpp:392-447): the part of a read and of its expanded CIGAR
 * that lies on the contig interval [left_to_polish, right_to_polish]. ops = BAM-encoded CIGAR (len << 4 | index in
 * "MIDNSHP=X"), pos_2_1 = Overlap.position_2_1. out = {posOnReadStart, posOnReadEnd, posOnCIGARStart, posOnCIGAREnd}.
 * Returns 0, or -2 when the reference marks the read "within a deletion" (interval.second[r] = -2). */
int32_t hso_clip_read(const uint32_t* ops, int64_t n_ops, int32_t pos_2_1, int32_t left_to_polish, int32_t right_to_polish,
                      int32_t* out);

/* edlibAlign (src/edlib/src/edlib.cpp:142-297) restated as a full dynamic program (hs_oracle_edlib.c).
 * mode 0 NW / 1 SHW / 2 HW, task 0 DISTANCE / 1 LOC / 2 PATH, k < 0 = unbounded. Output arrays must hold
 * n+1 locations and m+n alignment bytes. Returns edlib's status (0; 1 when its Hirschberg recursion finds no split row).
 * Paths at or above edlib's 1 MiB switch follow obtainAlignmentHirschberg's split rule (:1236-1401). */
int32_t hso_edlib_align(const char* query, int32_t m, const char* target, int32_t n, int32_t k, int32_t mode,
                        int32_t task, int32_t* edit_distance, int32_t* alphabet_length, int32_t* n_locations,
                        int32_t* end_locations, int32_t* start_locations, int32_t* alignment_length,
                        uint8_t* alignment);

#ifdef __cplusplus
}
#endif
#endif
