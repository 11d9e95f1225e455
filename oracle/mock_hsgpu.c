/* TEST INFRASTRUCTURE ONLY -- never linked, loaded or executed by the product path.
 *
 * A CPU stand-in for the handful of libhsgpu entry points integration/glue_call_variants.cpp calls, with the semantics
 * include/hsgpu.h documents, computed by the oracle (hs_oracle.c). It exists for ONE test: tests/test_host.py runs the
 * glued executable (the reference's own main() + the glue) with this library in front of the real one
 * (LD_LIBRARY_PATH) in a container without a GPU, and compares the .col / .vcf / error-rate files with the reference
 * executable's. That checks the glue -- the conversions between the reference's structures and the flat arrays of the
 * C ABI, the order of the calls, what main() gets back -- where no B200 is at hand; the same test with the real library
 * is tests/test_gpu_callvariants.py::test_reference_main_on_libhsgpu_gives_the_reference_files.
 * Built as oracle/_ref/mock/libhsgpu.so (same soname as the product library so that the dynamic loader takes it).
 * One contig per pileup (what the glue creates), u32 CIGAR ops only. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hsgpu.h"
#include "hs_oracle.h"

struct hsgpu_ctx {
    int device;
};

struct hsgpu_pileup {
    hsgpu_ctx* ctx;
    int32_t L, n_reads;
    uint8_t* contig;      /* codes 0..3 */
    uint8_t* read_bases;  /* codes, concatenated */
    int64_t* read_off;    /* [n_reads+1] bases */
    uint32_t* cigar;
    int64_t* cigar_off;
    int32_t* start;
    uint8_t* strand;
    /* built */
    int built, ranked;
    int64_t n_cells, stats[2];
    int64_t* col_off;
    uint32_t* read_idx;
    uint8_t* code;
    int32_t* read_end;
    /* ranked */
    uint8_t *ref_base, *second_base;
    int32_t n_suspects;
    int32_t* suspect_pos;
    uint8_t* suspect_auto;
    int64_t depth_sum;
};

static const char* g_error = "mock libhsgpu: no error";

int hsgpu_ctx_create(int device, hsgpu_ctx** out) {
    *out = (hsgpu_ctx*)calloc(1, sizeof(hsgpu_ctx));
    (*out)->device = device;
    return HSGPU_OK;
}
void hsgpu_ctx_destroy(hsgpu_ctx* ctx) { free(ctx); }
const char* hsgpu_last_error(hsgpu_ctx* ctx) {
    (void)ctx;
    return g_error;
}

void hsgpu_pack_bases_ascii(const char* seq, int64_t n, uint32_t* out) {
    for (int64_t w = 0; w < (n + 15) / 16; w++) out[w] = 0;
    for (int64_t i = 0; i < n; i++) {
        const char c = seq[i];
        const uint32_t b = c == 'A' ? 0u : (c == 'C' ? 1u : (c == 'G' ? 2u : 3u)); /* src/sequence.cpp:16-21 */
        out[i >> 4] |= b << (2 * (i & 15));
    }
}

int64_t hsgpu_parse_cigar(const char* cigar, uint32_t* out_ops, int64_t capacity) {
    static const char kinds[] = "MIDNSHP=X";
    if (cigar[0] == '*' && cigar[1] == 0) return 0;
    int64_t n = 0;
    uint32_t len = 0;
    int have = 0;
    for (const char* p = cigar; *p; p++) {
        if (*p >= '0' && *p <= '9') {
            len = len * 10 + (uint32_t)(*p - '0');
            have = 1;
            continue;
        }
        const char* k = strchr(kinds, *p);
        if (!k || !have) return HSGPU_ERR_ARG;
        if (n >= capacity) return HSGPU_ERR_CAPACITY;
        out_ops[n++] = (len << 4) | (uint32_t)(k - kinds);
        len = 0;
        have = 0;
    }
    return have ? HSGPU_ERR_ARG : n;
}

static void unpack(const uint32_t* words, int64_t n, uint8_t* out) {
    for (int64_t i = 0; i < n; i++) out[i] = (uint8_t)((words[i >> 4] >> (2 * (i & 15))) & 3u);
}

int hsgpu_pileup_create(hsgpu_ctx* ctx, const hsgpu_pileup_input* in, hsgpu_pileup** out) {
    if (in->n_contigs != 1 || !in->cigar || in->cigar16 || in->cigar8) {
        g_error = "mock libhsgpu: one contig per pileup, u32 CIGAR ops";
        return HSGPU_ERR_ARG;
    }
    hsgpu_pileup* p = (hsgpu_pileup*)calloc(1, sizeof(hsgpu_pileup));
    p->ctx = ctx;
    p->L = in->contig_len[0];
    p->n_reads = (int32_t)in->n_reads;
    const int32_t R = p->n_reads;
    p->contig = (uint8_t*)malloc((size_t)p->L + 1);
    unpack(in->contig_bases + in->contig_word_off[0], p->L, p->contig);
    p->read_off = (int64_t*)calloc((size_t)R + 1, sizeof(int64_t));
    for (int32_t r = 0; r < R; r++) p->read_off[r + 1] = p->read_off[r] + in->read_len[r];
    p->read_bases = (uint8_t*)malloc((size_t)p->read_off[R] + 1);
    for (int32_t r = 0; r < R; r++) unpack(in->read_bases + in->read_word_off[r], in->read_len[r], p->read_bases + p->read_off[r]);
    p->cigar_off = (int64_t*)malloc(((size_t)R + 1) * sizeof(int64_t));
    memcpy(p->cigar_off, in->cigar_off, ((size_t)R + 1) * sizeof(int64_t));
    p->cigar = (uint32_t*)malloc(((size_t)p->cigar_off[R] + 1) * sizeof(uint32_t));
    memcpy(p->cigar, in->cigar, (size_t)p->cigar_off[R] * sizeof(uint32_t));
    p->start = (int32_t*)malloc(((size_t)R + 1) * sizeof(int32_t));
    memcpy(p->start, in->read_start, (size_t)R * sizeof(int32_t));
    p->strand = (uint8_t*)malloc((size_t)R + 1);
    memcpy(p->strand, in->read_strand, (size_t)R);
    *out = p;
    return HSGPU_OK;
}

void hsgpu_pileup_destroy(hsgpu_pileup* p) {
    if (!p) return;
    free(p->contig); free(p->read_bases); free(p->read_off); free(p->cigar); free(p->cigar_off); free(p->start);
    free(p->strand); free(p->col_off); free(p->read_idx); free(p->code); free(p->read_end); free(p->ref_base);
    free(p->second_base); free(p->suspect_pos); free(p->suspect_auto);
    free(p);
}

int hsgpu_pileup_build(hsgpu_pileup* p) {
    p->col_off = (int64_t*)calloc((size_t)p->L + 1, sizeof(int64_t));
    p->read_end = (int32_t*)calloc((size_t)p->n_reads + 1, sizeof(int32_t));
    p->n_cells = hso_pileup(p->contig, p->L, p->n_reads, p->read_bases, p->read_off, p->cigar, p->cigar_off, p->start, p->strand,
                            0, p->col_off, NULL, NULL, p->stats, p->read_end);
    p->read_idx = (uint32_t*)malloc(((size_t)p->n_cells + 1) * sizeof(uint32_t));
    p->code = (uint8_t*)malloc((size_t)p->n_cells + 1);
    hso_pileup(p->contig, p->L, p->n_reads, p->read_bases, p->read_off, p->cigar, p->cigar_off, p->start, p->strand, p->n_cells,
               p->col_off, p->read_idx, p->code, p->stats, p->read_end);
    p->built = 1;
    return HSGPU_OK;
}

int hsgpu_pileup_stats(hsgpu_pileup* p, int64_t* n_cells, int64_t* distance_sum, int64_t* aligned_sum) {
    if (!p->built) return HSGPU_ERR_STATE;
    if (n_cells) *n_cells = p->n_cells;
    if (distance_sum) *distance_sum = p->stats[0];
    if (aligned_sum) *aligned_sum = p->stats[1];
    return HSGPU_OK;
}

float hsgpu_mean_distance(int64_t distance_sum, int64_t aligned_sum) { return hso_mean_distance(distance_sum, aligned_sum); }

int hsgpu_pileup_read_ends(hsgpu_pileup* p, int32_t* read_end) {
    if (!p->built) return HSGPU_ERR_STATE;
    memcpy(read_end, p->read_end, (size_t)p->n_reads * sizeof(int32_t));
    return HSGPU_OK;
}

int hsgpu_column_rank(hsgpu_pileup* p, const float* mean_error, float automatic_snp_threshold) {
    if (!p->built) return HSGPU_ERR_STATE;
    const float me = mean_error ? mean_error[0] : hso_mean_distance(p->stats[0], p->stats[1]);
    const int32_t cap = p->L / 6 + 2;  /* suspects are more than 5 columns apart (:529) */
    free(p->ref_base); free(p->second_base); free(p->suspect_pos); free(p->suspect_auto);
    p->ref_base = (uint8_t*)calloc((size_t)p->L + 1, 1);
    p->second_base = (uint8_t*)calloc((size_t)p->L + 1, 1);
    p->suspect_pos = (int32_t*)calloc((size_t)cap, sizeof(int32_t));
    p->suspect_auto = (uint8_t*)calloc((size_t)cap, 1);
    p->n_suspects = hso_call_variants(p->col_off, p->code, p->L, me, automatic_snp_threshold, p->ref_base, p->second_base,
                                      p->suspect_pos, p->suspect_auto, cap, &p->depth_sum);
    p->ranked = 1;
    return HSGPU_OK;
}

int hsgpu_column_counts(hsgpu_pileup* p, int32_t* n_suspects, int64_t* depth_sum) {
    if (!p->ranked) return HSGPU_ERR_STATE;
    if (n_suspects) *n_suspects = p->n_suspects;
    if (depth_sum) *depth_sum = p->depth_sum;
    return HSGPU_OK;
}

int hsgpu_suspects(hsgpu_pileup* p, int32_t contig, int32_t capacity, int32_t* pos, uint8_t* is_automatic) {
    if (!p->ranked || contig != 0) return HSGPU_ERR_STATE;
    if (capacity < p->n_suspects) return HSGPU_ERR_CAPACITY;
    memcpy(pos, p->suspect_pos, (size_t)p->n_suspects * sizeof(int32_t));
    memcpy(is_automatic, p->suspect_auto, (size_t)p->n_suspects);
    return HSGPU_OK;
}

int hsgpu_column_summary(hsgpu_pileup* p, int32_t contig, uint8_t* ref_base, uint8_t* second_base, uint32_t* counts,
                         uint32_t* depth) {
    if (!p->ranked || contig != 0 || counts || depth) return HSGPU_ERR_STATE;
    if (ref_base) memcpy(ref_base, p->ref_base, (size_t)p->L);
    if (second_base) memcpy(second_base, p->second_base, (size_t)p->L);
    return HSGPU_OK;
}

int hsgpu_pileup_extract_columns(hsgpu_pileup* p, int32_t contig, int32_t n_cols, const int32_t* pos, int64_t cell_capacity,
                                 int64_t* off, uint32_t* read_idx, uint8_t* code) {
    if (!p->ranked || contig != 0) return HSGPU_ERR_STATE;
    off[0] = 0;
    for (int32_t i = 0; i < n_cols; i++) {
        if (pos[i] < 0 || pos[i] >= p->L) return HSGPU_ERR_ARG;
        off[i + 1] = off[i] + (p->col_off[pos[i] + 1] - p->col_off[pos[i]]);
    }
    if (off[n_cols] > cell_capacity || !read_idx || !code) return off[n_cols] > 0 ? HSGPU_ERR_CAPACITY : HSGPU_OK;
    for (int32_t i = 0; i < n_cols; i++) {
        const int64_t n = off[i + 1] - off[i];
        memcpy(read_idx + off[i], p->read_idx + p->col_off[pos[i]], (size_t)n * sizeof(uint32_t));
        memcpy(code + off[i], p->code + p->col_off[pos[i]], (size_t)n);
    }
    return HSGPU_OK;
}

int hsgpu_robust_filter(hsgpu_pileup* p, int32_t contig, const hsgpu_partitions* parts, int32_t n_suspects,
                        const int32_t* suspect_pos, int32_t kept_capacity, int32_t* kept, int32_t* n_kept) {
    if (!p->ranked || contig != 0) return HSGPU_ERR_STATE;
    *n_kept = 0;
    if (parts->n_parts == 0 || p->L == 0) return HSGPU_OK;
    int32_t* all = (int32_t*)malloc(((size_t)p->L + 1) * sizeof(int32_t));
    const int32_t n = hso_robust_filter(p->L, p->col_off, p->read_idx, p->code, p->ref_base, p->second_base, parts->n_parts,
                                        parts->part_off, parts->read_idx, parts->state, parts->more, parts->less, n_suspects,
                                        suspect_pos, all);
    *n_kept = n;
    if (n > kept_capacity) {
        free(all);
        return HSGPU_ERR_CAPACITY;
    }
    memcpy(kept, all, (size_t)n * sizeof(int32_t));
    free(all);
    return HSGPU_OK;
}

/* integration/glue_separate_reads.cpp */
int hsgpu_read_pair_counts(hsgpu_ctx* ctx, int32_t n_reads, int32_t n_snps, const int64_t* snp_off, const uint32_t* read_idx,
                           const uint8_t* code, const uint8_t* ref_base, const uint8_t* second_base, int32_t* sim,
                           int32_t* diff) {
    (void)ctx;
    hso_read_pair_counts(n_reads, n_snps, snp_off, read_idx, code, ref_base, second_base, sim, diff);
    return HSGPU_OK;
}
