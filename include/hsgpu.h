/* libhsgpu -- B200 (sm_100a) implementation of HairSplitter's data-parallel core behind a C ABI.
 *
 * Reference = RolandFaure/Hairsplitter v1.9.22 (paths below are relative to its repository root).
 * The reference has no plugin API; the boundary this library replaces is the set of C++ module
 * functions on the hot path (src/call_variants.h:12-76, src/separate_reads.h:13-90) plus the one
 * real C ABI in the tree, edlib (src/edlib/include/edlib.h:146-271). Each entry point below names
 * the reference interface it stands in for. INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every function returns 0 (HSGPU_OK) or a negative hsgpu_status; no exceptions cross the ABI;
 *     hsgpu_last_error() gives the message of the last failure on that context.
 *   - plain pointers and sizes only. Input buffers are HOST memory owned by the caller (pageable or
 *     pinned; hsgpu_host_alloc gives pinned memory) unless a parameter says "device".
 *   - one hsgpu_ctx per (host thread, GPU); calls on one context are serialised on its CUDA stream,
 *     different contexts run concurrently, so the reference's OpenMP-over-contigs loops can call in.
 *   - there is NO CPU fallback: without a usable sm_100-class device hsgpu_ctx_create fails.
 *
 * Encodings (identical to the reference's in-memory forms)
 *   bases   2-bit code of `Sequence` (src/sequence.cpp:13-23): A=0 C=1 G=2 T/other=3, 16 bases per
 *           little-endian u32 word (base j in bits 2*(j%16)); reverse complement = reversed, 3-b.
 *   CIGAR   u32 per op = len<<4 | op, op = index in "MIDNSHP=X" (the BAM encoding of the SAM CIGAR
 *           the reference keeps as a string, src/read.h:23 / src/tools.cpp:27-57).
 *   codes   pileup cell = '!' + 5*i(c-2) + i(c-1) + 25*i(c0) over "ACGT-" (src/call_variants.cpp:238).
 */
#ifndef HSGPU_H
#define HSGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    HSGPU_OK = 0,
    HSGPU_ERR_CUDA = -1,       /* a CUDA runtime call failed */
    HSGPU_ERR_NO_DEVICE = -2,  /* no CUDA device / not compute capability 10.x */
    HSGPU_ERR_ARG = -3,        /* invalid argument */
    HSGPU_ERR_CAPACITY = -4,   /* caller buffer too small; sizes were written so the call can be retried */
    HSGPU_ERR_STATE = -5,      /* stage called out of order (e.g. rank before build) */
    HSGPU_ERR_LIMIT = -6       /* input exceeds a documented limit */
} hsgpu_status;

typedef struct hsgpu_ctx hsgpu_ctx;
typedef struct hsgpu_pileup hsgpu_pileup;

/* ---- context ------------------------------------------------------------------------------- */
int hsgpu_ctx_create(int device, hsgpu_ctx** out);
void hsgpu_ctx_destroy(hsgpu_ctx* ctx);
const char* hsgpu_last_error(hsgpu_ctx* ctx); /* ctx may be NULL: last error of the calling thread */
int hsgpu_sync(hsgpu_ctx* ctx);
/* number of kernels launched on this context so far (bench.py's gpu_launches) */
int64_t hsgpu_launch_count(hsgpu_ctx* ctx);
/* the CUDA stream of the context as an opaque handle (cudaStream_t), for event timing */
void* hsgpu_stream(hsgpu_ctx* ctx);
/* per-kernel timing: when enabled every kernel launch on the context is bracketed by CUDA events;
 * hsgpu_profile_report returns "name\tlaunches\ttotal_ms\n" lines accumulated since the last report
 * (the string is owned by the context and valid until the next call) */
int hsgpu_profile_enable(hsgpu_ctx* ctx, int on);
const char* hsgpu_profile_report(hsgpu_ctx* ctx);
/* pinned host memory for staging inputs/outputs */
int hsgpu_host_alloc(void** out, int64_t bytes);
void hsgpu_host_free(void* p);

/* ---- host-side packing (replaces Sequence::Sequence(string&), src/sequence.cpp:13-23, and the
 * string form of the CIGAR) ------------------------------------------------------------------ */
void hsgpu_pack_bases_ascii(const char* seq, int64_t n, uint32_t* out_words); /* ceil(n/16) words */
void hsgpu_pack_bases_codes(const uint8_t* codes, int64_t n, uint32_t* out_words);
/* parses a SAM CIGAR string into BAM ops; returns the number of ops or <0; "*" gives 0 ops */
int64_t hsgpu_parse_cigar(const char* cigar, uint32_t* out_ops, int64_t capacity);
/* BAM ops -> the 8-bit form of hsgpu_pileup_input.cigar8. Returns the number of bytes written (>= n_ops: ops longer
 * than 63 are split), HSGPU_ERR_CAPACITY, or HSGPU_ERR_ARG when an op has no 8-bit form (N, P: use cigar16).
 * out == NULL only counts. */
int64_t hsgpu_pack_cigar8(const uint32_t* ops, int64_t n_ops, uint8_t* out, int64_t capacity);

/* ---- pileup: generate_msa (src/call_variants.cpp:50-437) -------------------------------------
 * A batch of contig chunks with the reads aligned on them (what parse_SAM + parse_reads_on_contig
 * leave in allOverlaps/allreads for each backbone, src/input_output.cpp:274-569). Reads of contig c
 * are [contig_read_off[c], contig_read_off[c+1]) and keep the neighbour order n of the reference. */
typedef struct {
    int32_t n_contigs;
    const int32_t* contig_len;        /* [n_contigs] columns L */
    const uint32_t* contig_bases;     /* packed; contig c starts at word contig_word_off[c] */
    const int64_t* contig_word_off;   /* [n_contigs+1] */
    const int64_t* contig_read_off;   /* [n_contigs+1] */
    int64_t n_reads;
    const uint32_t* read_bases;       /* packed, ORIGINAL orientation; read r starts at word read_word_off[r] */
    const int64_t* read_word_off;     /* [n_reads+1] */
    const int32_t* read_len;          /* [n_reads] bases */
    const uint32_t* cigar;            /* all ops, concatenated */
    const int64_t* cigar_off;         /* [n_reads+1] */
    const int32_t* read_start;        /* [n_reads] Overlap.position_2_1 = POS-1 */
    const uint8_t* read_strand;       /* [n_reads] Overlap.strand, 1 = forward */
    /* optional compact CIGAR: u16 per op = len<<4 | op with len <= 4095 (longer ops are written as several
     * ops of the same kind, which every loop of the reference treats the same way). When non-NULL it replaces
     * `cigar` (which may then be NULL), cigar_off indexes it, and half as many bytes cross PCIe. */
    const uint16_t* cigar16;
    /* optional 8-bit CIGAR: u8 per op = len<<2 | kind, len <= 63, kind 0 = M/=/X, 1 = I, 2 = D, 3 = S/H (the
     * four classes generate_msa distinguishes, src/call_variants.cpp:226-342; longer ops are split). Made by
     * hsgpu_pack_cigar8. When non-NULL it replaces `cigar` and `cigar16`, cigar_off indexes it: a quarter of
     * the CIGAR bytes cross PCIe (the e2e path is bound by that link). */
    const uint8_t* cigar8;
} hsgpu_pileup_input;

/* copies the batch to the device and returns when the upload has landed (the caller's buffers are free again).
 * Uploads take turns per device: while one context uploads, the kernels of the other contexts keep running, which
 * is what overlaps transfer and compute when every host thread drives its own context. Pinned host buffers
 * (hsgpu_host_alloc) reach the full link rate. On failure nothing is left allocated. */
int hsgpu_pileup_create(hsgpu_ctx* ctx, const hsgpu_pileup_input* in, hsgpu_pileup** out);
void hsgpu_pileup_destroy(hsgpu_pileup* p);

/* runs the CIGAR walk of every read and writes the device-resident pileup (hot loop A) */
int hsgpu_pileup_build(hsgpu_pileup* p);

/* per contig: total cells, and the integer sums behind generate_msa's return value
 * (totalDistance numerator, totalLengthOfAlignment without its initial 1). Any pointer may be NULL. */
int hsgpu_pileup_stats(hsgpu_pileup* p, int64_t* n_cells, int64_t* distance_sum, int64_t* aligned_sum);
/* sizes behind the roofline figures of bench.py. info[8]: [0] CIGAR bytes on the device (one byte per op), [1] pileup
 * bytes, [2] (tile, read) index entries, [3] reads with clips inside the alignment; of the last hsgpu_robust_filter*
 * call: [4] active columns, [5] their cells, [6] partition state bytes read, [7] columns kept */
int hsgpu_pileup_info(hsgpu_pileup* p, int64_t* info);
/* generate_msa's float return value from the two sums (float accumulator, :67-68,434) */
float hsgpu_mean_distance(int64_t distance_sum, int64_t aligned_sum);
/* positionOfReads[n].second of every read (:354), i.e. readLimits */
int hsgpu_pileup_read_ends(hsgpu_pileup* p, int32_t* read_end);

/* the whole pileup of one contig in the reference's own layout (vector<Column>: per column the
 * ascending neighbour indices and their codes). col_off has L+1 entries and is always written;
 * cells are written when cell_capacity >= col_off[L], else HSGPU_ERR_CAPACITY. Needs column_rank. */
int hsgpu_pileup_export(hsgpu_pileup* p, int32_t contig, int64_t cell_capacity, int64_t* col_off,
                        uint32_t* read_idx, uint8_t* code);
/* the same for a sorted subset of columns (what output_files / Partition need for SNP columns) */
int hsgpu_pileup_extract_columns(hsgpu_pileup* p, int32_t contig, int32_t n_cols, const int32_t* pos,
                                 int64_t cell_capacity, int64_t* off, uint32_t* read_idx, uint8_t* code);

/* ---- allele counting: call_variants (src/call_variants.cpp:447-567) ---------------------------
 * Per column: histogram of codes, ranking with the reference's tie-breaking (robin_hood iteration
 * order + libstdc++ std::sort), ref_base/second_base, suspect predicate (:525-529) and "automatic"
 * flag (:531). mean_error may be NULL (then generate_msa's own value per contig is used, as in
 * main() :1307-1322), else [n_contigs] floats. Results stay on the device for the next stages. */
int hsgpu_column_rank(hsgpu_pileup* p, const float* mean_error, float automatic_snp_threshold);

/* per-contig results of call_variants: number of suspect columns, depth numerator (cells counted) */
int hsgpu_column_counts(hsgpu_pileup* p, int32_t* n_suspects, int64_t* depth_sum);
/* suspect positions of one contig (ascending) and whether each is an "automatic" SNP */
int hsgpu_suspects(hsgpu_pileup* p, int32_t contig, int32_t capacity, int32_t* pos, uint8_t* is_automatic);
/* the same for every contig of the batch in one call (two host round trips in all instead of two per contig):
 * contig c's suspects are pos[off[c] .. off[c+1]) (positions inside contig c, ascending); off has n_contigs+1
 * entries; depth_sum as in hsgpu_column_counts; pos / is_automatic / depth_sum may be NULL. HSGPU_ERR_CAPACITY
 * (with off filled in) when capacity < off[n_contigs]. */
int hsgpu_suspects_all(hsgpu_pileup* p, int64_t capacity, int32_t* pos, uint8_t* is_automatic, int64_t* off,
                       int64_t* depth_sum);
/* per-column summary of one contig; any pointer may be NULL. counts = c0,c1,c2 interleaved [3L] */
int hsgpu_column_summary(hsgpu_pileup* p, int32_t contig, uint8_t* ref_base, uint8_t* second_base,
                         uint32_t* counts, uint32_t* depth);

/* ---- partition x column contingency: distance(Partition&,Column&,char) + computeChiSquare
 * (src/call_variants.cpp:778-967,1135-1163), as used by keep_only_robust_variants (:577-768) -------
 * Partitions of one contig as the reference's parallel vectors (Partition::getReads/getPartition/
 * getMore/getLess, src/Partition.h:36-41): partition k = entries [part_off[k], part_off[k+1]). */
typedef struct {
    int32_t n_parts;
    const int64_t* part_off;   /* [n_parts+1] */
    const int32_t* read_idx;   /* ascending neighbour indices */
    const int16_t* state;      /* mostFrequentBases: 1, -1, 0, -2 (masked) */
    const int32_t* more;       /* moreFrequence */
    const int32_t* less;       /* lessFrequence */
} hsgpu_partitions;

typedef struct {
    int32_t n00, n01, n10, n11;
    int32_t solid00, solid01, solid10, solid11;
    uint8_t second_base; /* distancePartition.secondBase; 0 when not augmented */
    uint8_t augmented;
    uint8_t pad[2];
    float chi_square;    /* computeChiSquare of this table */
} hsgpu_distance;

/* tables for every (column in pos[], partition) pair against ref_base = the column's own ref_base
 * (how loops 3 and 4 call it, :725,755); out is [n_cols * n_parts], column-major by position. */
int hsgpu_partition_tables(hsgpu_pileup* p, int32_t contig, const hsgpu_partitions* parts, int32_t n_cols,
                           const int32_t* pos, hsgpu_distance* out);

/* loops 3+4 of keep_only_robust_variants fused (:721-764): given the suspect positions (snps_in) and
 * the final partitions, returns the ascending positions of snps_out. kept_capacity bounds `kept`;
 * *n_kept is always written. */
int hsgpu_robust_filter(hsgpu_pileup* p, int32_t contig, const hsgpu_partitions* parts, int32_t n_suspects,
                        const int32_t* suspect_pos, int32_t kept_capacity, int32_t* kept, int32_t* n_kept);

/* Batch form of the same filter (what HS_call_variants' loop over contigs becomes on one GPU, src/call_variants.cpp:
 * 1276-1330): hsgpu_partitions_set uploads the final partitions of EVERY contig of the pileup (parts[c] for contig c,
 * n_parts may be 0) and keeps them with the pileup until it is called again; hsgpu_robust_filter_all then runs loops
 * 3+4 for all contigs in one launch, with the pileup's own suspect columns (hsgpu_column_rank) as snps_in -- the way
 * main() chains call_variants and keep_only_robust_variants (:1322-1327). Contig c's snps_out are
 * kept[off[c] .. off[c+1]) (positions inside the contig, ascending); off has n_contigs+1 entries and is always written;
 * kept may be NULL (counts only); HSGPU_ERR_CAPACITY when capacity < off[n_contigs]. */
int hsgpu_partitions_set(hsgpu_pileup* p, const hsgpu_partitions* parts);
int hsgpu_robust_filter_all(hsgpu_pileup* p, int64_t capacity, int32_t* kept, int64_t* off);

/* ---- read x read counts: list_similarities_and_differences_between_reads3
 * (src/separate_reads.cpp:374-433) --------------------------------------------------------------
 * SNP columns of one contig as parsed from the .col file (parse_column_file, :46-190): CSR over SNPs
 * with neighbour indices and codes, plus ref_base/second_base per SNP. sim/diff are dense
 * n_reads x n_reads int32 (row-major): similarity = 3*A*At + R*Rt, difference = A*Rt + R*At, zero diagonal.
 * The dense form exists for the caller only (at most 46000 reads, HSGPU_ERR_LIMIT beyond): on the device the counts
 * are kept as 128 x 128 blocks over the band of reads that share SNPs, which is what hsgpu_graph_* read, so a
 * contig's memory follows its read overlaps, not n_reads^2. */
int hsgpu_read_pair_counts(hsgpu_ctx* ctx, int32_t n_reads, int32_t n_snps, const int64_t* snp_off,
                           const uint32_t* read_idx, const uint8_t* code, const uint8_t* ref_base,
                           const uint8_t* second_base, int32_t* sim, int32_t* diff);

/* The same contraction for a batch of contigs with the operands and results resident on the device
 * (HS_separate_reads calls ..._reads3 once per contig inside its OpenMP loop, src/separate_reads.cpp:1507-1560;
 * a batch is what one GPU receives from the contig sharding). SNPs of all contigs are concatenated:
 * contig c owns SNPs [snp_base[c], snp_base[c+1]) (snp_base has n_contigs+1 entries), snp_off has one
 * entry per SNP plus one and indexes read_idx/code; read_idx is local to the contig (< n_reads[c]).
 * create uploads the columns and builds the one-hot operands, compute runs the tensor-core kernel
 * (asynchronous on the context's stream, may be repeated), fetch copies one contig's n x n row-major
 * matrices to the host and synchronises. */
typedef struct hsgpu_pairs hsgpu_pairs;
#define HSGPU_PAIRS_DENSE 1      /* schedule every tile pair over every SNP block (no band pruning): measurement */
#define HSGPU_PAIRS_KEEP_ORDER 2 /* do not reorder reads by their first SNP */
#define HSGPU_PAIRS_SIMT 4       /* plain integer-pipe kernel instead of tcgen05: the A/B check used by the tests */
int hsgpu_pairs_create(hsgpu_ctx* ctx, int32_t n_contigs, const int32_t* n_reads, const int64_t* snp_base,
                       const int64_t* snp_off, const uint32_t* read_idx, const uint8_t* code, const uint8_t* ref_base,
                       const uint8_t* second_base, int32_t flags, hsgpu_pairs** out);
int hsgpu_pairs_compute(hsgpu_pairs* h);
int hsgpu_pairs_fetch(hsgpu_pairs* h, int32_t contig, int32_t* sim, int32_t* diff);
/* info[8]: tile pairs scheduled, tile pairs of the full upper triangles, 128-SNP blocks executed, the
 * same for the full triangles, operand rows, operand row stride (bytes), output elements per matrix,
 * 1 if no contig needed reordering */
int hsgpu_pairs_info(hsgpu_pairs* h, int64_t* info);
void hsgpu_pairs_destroy(hsgpu_pairs* h);

/* ---- read graph + chinese whispers of the windows of a batch: create_read_graph_matrix
 * (src/separate_reads.cpp:706-828) and chinese_whispers_high_memory (src/cluster_graph.cpp:240-310) ----
 * A window is the set of reads that span it (the reference's mask_at_this_position, :1590-1622): window w of
 * contig win_contig[w] owns the ascending read indices win_reads[win_off[w] .. win_off[w+1]). The similarity /
 * difference counts are read from the device-resident results of `pairs` (hsgpu_pairs_compute must have run;
 * `pairs` must outlive the graph). Everything below is expressed in LOCAL indices: position of a read in its
 * window's list. */
typedef struct hsgpu_graph hsgpu_graph;
int hsgpu_graph_create(hsgpu_pairs* pairs, int32_t n_windows, const int32_t* win_contig, const int64_t* win_off,
                       const int32_t* win_reads, float error_rate, hsgpu_graph** out);
/* The same with a per-window switch (win_low_memory[w] != 0, array may be NULL): the window's neighbour choice follows
 * create_read_graph_low_memory (src/separate_reads.cpp:538-693, what the reference runs with -l, for amplicons and
 * above 1000x) instead of create_read_graph_matrix: no `similarity > 0` guard on the distance, reads that appear in
 * no SNP column are skipped. The counts are the same contig-wide similarity / difference counts of `pairs`; the
 * reference's pairwise loop indexes a read's alleles by offset from its first SNP (:597-605), which gives those
 * counts whenever every read covers a run of consecutive SNP columns -- the caller checks that (HS_separate_reads
 * does, and falls back on its host loop otherwise). Limits: at most 6400 reads per window here, 3200 for
 * hsgpu_graph_whispers (HSGPU_ERR_LIMIT beyond). */
int hsgpu_graph_create_ex(hsgpu_pairs* pairs, int32_t n_windows, const int32_t* win_contig, const int64_t* win_off,
                          const int32_t* win_reads, const uint8_t* win_low_memory, float error_rate, hsgpu_graph** out);
/* builds the adjacency of every window. n_replayed (may be NULL) receives the number of reads whose neighbour
 * choice depended on std::sort's order of equal distances and was replayed with the reference's sort. */
int hsgpu_graph_build(hsgpu_graph* g, int64_t* n_replayed);
/* the symmetric 0/1 adjacency as a CSR over all masked reads of all windows (rows in window order): neighbours of
 * masked read i are adj[adj_off[i] .. adj_off[i+1]), local indices, ascending. adj_off has win_off[n_windows]+1
 * entries; *n_adj is always written; HSGPU_ERR_CAPACITY when capacity < *n_adj. */
int hsgpu_graph_adjacency(hsgpu_graph* g, int64_t* adj_off, int64_t capacity, int32_t* adj, int64_t* n_adj);
/* n_runs clusterings. Run i works on window run_window[i] and starts from init_labels (the runs' label vectors
 * concatenated, m entries each, labels = local indices or negative = "no cluster"); labels_out has the same
 * layout. Node order of sweep s = the masked reads sorted by order_rank[min(s, n_orders-1)], where order_rank
 * holds, per contig (concatenated in contig order), n_orders arrays of n_reads positions: order_rank[k][r] = place
 * of read r in the k-th shuffled order of all reads of the contig (the reference shuffles 0..n_reads-1 afresh in
 * every sweep, cluster_graph.cpp:255-259, and skips the reads outside the window). */
int hsgpu_graph_whispers(hsgpu_graph* g, int64_t n_runs, const int32_t* run_window, const int32_t* init_labels,
                         int32_t n_orders, const int32_t* order_rank, int32_t* labels_out);
void hsgpu_graph_destroy(hsgpu_graph* g);

/* ---- read clipping / window extraction of modify_GFA (src/create_new_contigs.cpp:383-447) ----------------------
 * For every window (interval) of a contig and every read with a cluster in it, the reference walks the read's
 * expanded CIGAR from its first character to find the part of the read (posOnReadStart/End) and of the CIGAR
 * (posOnCIGARStart/End) lying on [leftToPolish, rightToPolish] of the contig, then hands
 * seq.substr(read_start, read_end - read_start) and convert_cigar2(cigar.substr(cigar_start, cigar_end - cigar_start))
 * to the polisher (:449-461). hsgpu_clip_reads does the walks of a batch of (read, interval) items on the GPU.
 * CIGARs are BAM ops (len << 4 | index in "MIDNSHP=X") of n_reads alignments, concatenated, read r = ops
 * [cigar_off[r], cigar_off[r+1]); pos_2_1[r] = Overlap.position_2_1. Item i clips read item_read[i] to
 * [left_to_polish[i], right_to_polish[i]]. status -2 = the reference sets interval.second[r] = -2 ("within a
 * deletion", :442-447) and skips the read; the other fields are then unspecified. The clipped CIGAR is returned as a
 * range of the read's ops: op_first without its first op_first_skip characters, the ops up to op_last (exclusive),
 * and the first op_last_take characters of op_last (op_last may equal the number of ops, with op_last_take 0);
 * adjacent ops of the same letter merge into one, as convert_cigar2 does. */
typedef struct {
    int32_t status;
    int32_t read_start, read_end;
    int32_t cigar_start, cigar_end;
    int32_t op_first, op_first_skip;
    int32_t op_last, op_last_take;
} hsgpu_clip;
int hsgpu_clip_reads(hsgpu_ctx* ctx, int64_t n_reads, const uint32_t* cigar, const int64_t* cigar_off,
                     const int32_t* pos_2_1, int64_t n_items, const int64_t* item_read, const int32_t* left_to_polish,
                     const int32_t* right_to_polish, hsgpu_clip* out);

/* ---- realignment: edlibAlign (src/edlib/include/edlib.h:146-271, src/edlib/src/edlib.cpp:142-297)
 * Batch of (query, target) pairs, results with edlib's exact field semantics. Modes/tasks use edlib's
 * numeric values: mode 0 NW, 1 SHW, 2 HW; task 0 DISTANCE, 1 LOC, 2 PATH. Queries up to 2^20, targets up to 2^30.
 * Paths at or above edlib's 1 MiB switch ((20*ceil(q/64)+8)*columns >= 1 MiB, edlib.cpp:1193-1195) are split the
 * way obtainAlignmentHirschberg (edlib.cpp:1236-1401) splits them, so the ops are edlib's there too. Pairs with
 * queries of at most 2048 whose path lies below the switch take the throughput kernels; the others are redone by a
 * second, slower launch over just those pairs. */
typedef struct {
    int32_t status;               /* edlib's: 0 = EDLIB_STATUS_OK, 1 = EDLIB_STATUS_ERROR */
    int32_t edit_distance;        /* -1 if larger than k */
    int32_t n_locations;
    int32_t alignment_length;
    int32_t alphabet_length;      /* EdlibAlignResult.alphabetLength */
    int32_t has_start_locations;  /* 0 where edlib leaves startLocations NULL (DISTANCE task, empty sequence) */
    int64_t loc_off;              /* offset into end_locations / start_locations */
    int64_t aln_off;              /* offset into alignment */
} hsgpu_edlib_result;

int hsgpu_edlib_align_batch(hsgpu_ctx* ctx, int32_t n_pairs, const char* queries, const int64_t* query_off,
                            const char* targets, const int64_t* target_off, int32_t k, int32_t mode, int32_t task,
                            hsgpu_edlib_result* results, int32_t* end_locations, int32_t* start_locations,
                            int64_t loc_capacity, uint8_t* alignment, int64_t aln_capacity);

/* Single-pair form with edlib's own calling convention, so that the four call sites of the reference
 * (src/create_new_contigs.cpp:557-630, src/tools.cpp:508-536) change by one line:
 *     EdlibAlignResult r = edlibAlign(q, ql, t, tl, cfg);   ->   hsgpu_edlibAlign(ctx, q, ql, t, tl, cfg);
 * The two structs are layout-compatible with EdlibAlignConfig / EdlibAlignResult (src/edlib/include/edlib.h:
 * 100-106, 213-262): a maintainer may pass edlib's own objects through a cast; edlibAlignmentToCigar works on the
 * result unchanged. endLocations / startLocations / alignment are malloc'd by the library and released by
 * hsgpu_edlibFreeAlignResult (or free()), NULL where edlib leaves them NULL (distance above k, DISTANCE task,
 * empty sequence). status: 0 = EDLIB_STATUS_OK, 1 = EDLIB_STATUS_ERROR (also: additional equalities are not
 * supported). One pair per call costs a kernel launch and a round trip: batch the pairs where the caller can. */
typedef struct {
    int k;
    int mode; /* EdlibAlignMode: 0 NW, 1 SHW, 2 HW */
    int task; /* EdlibAlignTask: 0 DISTANCE, 1 LOC, 2 PATH */
    const void* additionalEqualities;
    int additionalEqualitiesLength;
} hsgpu_EdlibAlignConfig;
typedef struct {
    int status;
    int editDistance;
    int* endLocations;
    int* startLocations;
    int numLocations;
    unsigned char* alignment;
    int alignmentLength;
    int alphabetLength;
} hsgpu_EdlibAlignResult;
hsgpu_EdlibAlignResult hsgpu_edlibAlign(hsgpu_ctx* ctx, const char* query, int queryLength, const char* target,
                                        int targetLength, hsgpu_EdlibAlignConfig config);
void hsgpu_edlibFreeAlignResult(hsgpu_EdlibAlignResult result);

#ifdef __cplusplus
}
#endif
#endif /* HSGPU_H */
