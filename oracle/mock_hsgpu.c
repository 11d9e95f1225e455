/* TEST INFRASTRUCTURE ONLY -- never linked, loaded or executed by the product path.
 *
 * A CPU stand-in for the libhsgpu entry points that the reference-side glue (integration/glue_*.cpp) and the host side of
 * bin/HS_call_variants call, with the semantics include/hsgpu.h documents, computed by the oracle (hs_oracle.c). It exists
 * for tests/test_host.py only: in a container without a GPU those tests run the glued executables (the reference's own
 * main() + the glue) and the drop-in executable with this library in front of the real one (LD_LIBRARY_PATH in that one
 * subprocess) and compare the .col / .vcf / error-rate / .gro files with the reference executable's. What that checks is
 * HOST code -- the conversions between the reference's structures and the flat arrays of the C ABI, the parsers, the
 * partition building, the merge, the writers -- where no B200 is at hand. The same comparisons with the real library
 * are the GPU tests (tests/test_gpu_callvariants.py, tests/test_gpu_sepreads.py).
 * Built as oracle/_ref/mock_for_glue_test/libhsgpu.so (same soname as the product library so that the dynamic loader
 * takes it). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hsgpu.h"
#include "hs_oracle.h"

struct hsgpu_ctx {
    int device;
};

typedef struct {
    int32_t L, n_reads;
    uint8_t* contig;      /* codes 0..3 */
    uint8_t* read_bases;  /* codes, concatenated */
    int64_t* read_off;    /* [n_reads+1] bases */
    uint32_t* cigar;
    int64_t* cigar_off;
    int32_t* start;
    uint8_t* strand;
    /* built */
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
    /* partitions (hsgpu_partitions_set) */
    int32_t n_parts;
    int64_t* part_off;
    int32_t *p_idx, *p_more, *p_less;
    int16_t* p_state;
} mock_contig;

struct hsgpu_pileup {
    hsgpu_ctx* ctx;
    int32_t n_contigs;
    mock_contig* c;
    int built, ranked, have_parts;
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
    if (in->cigar8 || (!in->cigar && !in->cigar16)) {
        g_error = "mock libhsgpu: u32 or u16 CIGAR ops";
        return HSGPU_ERR_ARG;
    }
    hsgpu_pileup* p = (hsgpu_pileup*)calloc(1, sizeof(hsgpu_pileup));
    p->ctx = ctx;
    p->n_contigs = in->n_contigs;
    p->c = (mock_contig*)calloc((size_t)in->n_contigs + 1, sizeof(mock_contig));
    for (int32_t ci = 0; ci < in->n_contigs; ci++) {
        mock_contig* c = &p->c[ci];
        const int64_t r0 = in->contig_read_off[ci], r1 = in->contig_read_off[ci + 1];
        const int32_t R = (int32_t)(r1 - r0);
        c->L = in->contig_len[ci];
        c->n_reads = R;
        c->contig = (uint8_t*)malloc((size_t)c->L + 1);
        unpack(in->contig_bases + in->contig_word_off[ci], c->L, c->contig);
        c->read_off = (int64_t*)calloc((size_t)R + 1, sizeof(int64_t));
        c->cigar_off = (int64_t*)calloc((size_t)R + 1, sizeof(int64_t));
        for (int32_t r = 0; r < R; r++) {
            c->read_off[r + 1] = c->read_off[r] + in->read_len[r0 + r];
            c->cigar_off[r + 1] = c->cigar_off[r] + (in->cigar_off[r0 + r + 1] - in->cigar_off[r0 + r]);
        }
        c->read_bases = (uint8_t*)malloc((size_t)c->read_off[R] + 1);
        c->cigar = (uint32_t*)malloc(((size_t)c->cigar_off[R] + 1) * sizeof(uint32_t));
        c->start = (int32_t*)malloc(((size_t)R + 1) * sizeof(int32_t));
        c->strand = (uint8_t*)malloc((size_t)R + 1);
        for (int32_t r = 0; r < R; r++) {
            unpack(in->read_bases + in->read_word_off[r0 + r], in->read_len[r0 + r], c->read_bases + c->read_off[r]);
            const int64_t k0 = in->cigar_off[r0 + r], n = in->cigar_off[r0 + r + 1] - k0;
            for (int64_t k = 0; k < n; k++)  /* the 16-bit form is len << 4 | op too */
                c->cigar[c->cigar_off[r] + k] = in->cigar16 ? (uint32_t)in->cigar16[k0 + k] : in->cigar[k0 + k];
            c->start[r] = in->read_start[r0 + r];
            c->strand[r] = in->read_strand[r0 + r];
        }
    }
    *out = p;
    return HSGPU_OK;
}

static void free_parts(mock_contig* c) {
    free(c->part_off); free(c->p_idx); free(c->p_more); free(c->p_less); free(c->p_state);
    c->part_off = NULL; c->p_idx = c->p_more = c->p_less = NULL; c->p_state = NULL;
    c->n_parts = 0;
}

void hsgpu_pileup_destroy(hsgpu_pileup* p) {
    if (!p) return;
    for (int32_t ci = 0; ci < p->n_contigs; ci++) {
        mock_contig* c = &p->c[ci];
        free(c->contig); free(c->read_bases); free(c->read_off); free(c->cigar); free(c->cigar_off); free(c->start);
        free(c->strand); free(c->col_off); free(c->read_idx); free(c->code); free(c->read_end); free(c->ref_base);
        free(c->second_base); free(c->suspect_pos); free(c->suspect_auto);
        free_parts(c);
    }
    free(p->c);
    free(p);
}

int hsgpu_pileup_build(hsgpu_pileup* p) {
    for (int32_t ci = 0; ci < p->n_contigs; ci++) {
        mock_contig* c = &p->c[ci];
        c->col_off = (int64_t*)calloc((size_t)c->L + 1, sizeof(int64_t));
        c->read_end = (int32_t*)calloc((size_t)c->n_reads + 1, sizeof(int32_t));
        c->n_cells = hso_pileup(c->contig, c->L, c->n_reads, c->read_bases, c->read_off, c->cigar, c->cigar_off, c->start,
                                c->strand, 0, c->col_off, NULL, NULL, c->stats, c->read_end);
        c->read_idx = (uint32_t*)malloc(((size_t)c->n_cells + 1) * sizeof(uint32_t));
        c->code = (uint8_t*)malloc((size_t)c->n_cells + 1);
        hso_pileup(c->contig, c->L, c->n_reads, c->read_bases, c->read_off, c->cigar, c->cigar_off, c->start, c->strand,
                   c->n_cells, c->col_off, c->read_idx, c->code, c->stats, c->read_end);
    }
    p->built = 1;
    return HSGPU_OK;
}

int hsgpu_pileup_stats(hsgpu_pileup* p, int64_t* n_cells, int64_t* distance_sum, int64_t* aligned_sum) {
    if (!p->built) return HSGPU_ERR_STATE;
    for (int32_t ci = 0; ci < p->n_contigs; ci++) {
        if (n_cells) n_cells[ci] = p->c[ci].n_cells;
        if (distance_sum) distance_sum[ci] = p->c[ci].stats[0];
        if (aligned_sum) aligned_sum[ci] = p->c[ci].stats[1];
    }
    return HSGPU_OK;
}

float hsgpu_mean_distance(int64_t distance_sum, int64_t aligned_sum) { return hso_mean_distance(distance_sum, aligned_sum); }

int hsgpu_pileup_read_ends(hsgpu_pileup* p, int32_t* read_end) {
    if (!p->built) return HSGPU_ERR_STATE;
    for (int32_t ci = 0; ci < p->n_contigs; ci++) {
        memcpy(read_end, p->c[ci].read_end, (size_t)p->c[ci].n_reads * sizeof(int32_t));
        read_end += p->c[ci].n_reads;
    }
    return HSGPU_OK;
}

int hsgpu_column_rank(hsgpu_pileup* p, const float* mean_error, float automatic_snp_threshold) {
    if (!p->built) return HSGPU_ERR_STATE;
    for (int32_t ci = 0; ci < p->n_contigs; ci++) {
        mock_contig* c = &p->c[ci];
        const float me = mean_error ? mean_error[ci] : hso_mean_distance(c->stats[0], c->stats[1]);
        const int32_t cap = c->L / 6 + 2; /* suspects are more than 5 columns apart (:529) */
        free(c->ref_base); free(c->second_base); free(c->suspect_pos); free(c->suspect_auto);
        c->ref_base = (uint8_t*)calloc((size_t)c->L + 1, 1);
        c->second_base = (uint8_t*)calloc((size_t)c->L + 1, 1);
        c->suspect_pos = (int32_t*)calloc((size_t)cap, sizeof(int32_t));
        c->suspect_auto = (uint8_t*)calloc((size_t)cap, 1);
        c->n_suspects = hso_call_variants(c->col_off, c->code, c->L, me, automatic_snp_threshold, c->ref_base, c->second_base,
                                          c->suspect_pos, c->suspect_auto, cap, &c->depth_sum);
    }
    p->ranked = 1;
    return HSGPU_OK;
}

int hsgpu_column_counts(hsgpu_pileup* p, int32_t* n_suspects, int64_t* depth_sum) {
    if (!p->ranked) return HSGPU_ERR_STATE;
    for (int32_t ci = 0; ci < p->n_contigs; ci++) {
        if (n_suspects) n_suspects[ci] = p->c[ci].n_suspects;
        if (depth_sum) depth_sum[ci] = p->c[ci].depth_sum;
    }
    return HSGPU_OK;
}

int hsgpu_suspects(hsgpu_pileup* p, int32_t contig, int32_t capacity, int32_t* pos, uint8_t* is_automatic) {
    if (!p->ranked || contig < 0 || contig >= p->n_contigs) return HSGPU_ERR_STATE;
    const mock_contig* c = &p->c[contig];
    if (capacity < c->n_suspects) return HSGPU_ERR_CAPACITY;
    memcpy(pos, c->suspect_pos, (size_t)c->n_suspects * sizeof(int32_t));
    memcpy(is_automatic, c->suspect_auto, (size_t)c->n_suspects);
    return HSGPU_OK;
}

int hsgpu_column_summary(hsgpu_pileup* p, int32_t contig, uint8_t* ref_base, uint8_t* second_base, uint32_t* counts,
                         uint32_t* depth) {
    if (!p->ranked || contig < 0 || contig >= p->n_contigs || counts || depth) return HSGPU_ERR_STATE;
    const mock_contig* c = &p->c[contig];
    if (ref_base) memcpy(ref_base, c->ref_base, (size_t)c->L);
    if (second_base) memcpy(second_base, c->second_base, (size_t)c->L);
    return HSGPU_OK;
}

int hsgpu_pileup_extract_columns(hsgpu_pileup* p, int32_t contig, int32_t n_cols, const int32_t* pos, int64_t cell_capacity,
                                 int64_t* off, uint32_t* read_idx, uint8_t* code) {
    if (!p->ranked || contig < 0 || contig >= p->n_contigs) return HSGPU_ERR_STATE;
    const mock_contig* c = &p->c[contig];
    off[0] = 0;
    for (int32_t i = 0; i < n_cols; i++) {
        if (pos[i] < 0 || pos[i] >= c->L) return HSGPU_ERR_ARG;
        off[i + 1] = off[i] + (c->col_off[pos[i] + 1] - c->col_off[pos[i]]);
    }
    if (off[n_cols] > cell_capacity || !read_idx || !code) return off[n_cols] > 0 ? HSGPU_ERR_CAPACITY : HSGPU_OK;
    for (int32_t i = 0; i < n_cols; i++) {
        const int64_t n = off[i + 1] - off[i];
        memcpy(read_idx + off[i], c->read_idx + c->col_off[pos[i]], (size_t)n * sizeof(uint32_t));
        memcpy(code + off[i], c->code + c->col_off[pos[i]], (size_t)n);
    }
    return HSGPU_OK;
}

static int32_t filter_contig(const mock_contig* c, const hsgpu_partitions* parts, int32_t n_suspects, const int32_t* suspect_pos,
                             int32_t* kept /* room for L */) {
    if (parts->n_parts == 0 || c->L == 0) return 0; /* :640-642 */
    return hso_robust_filter(c->L, c->col_off, c->read_idx, c->code, c->ref_base, c->second_base, parts->n_parts, parts->part_off,
                             parts->read_idx, parts->state, parts->more, parts->less, n_suspects, suspect_pos, kept);
}

int hsgpu_robust_filter(hsgpu_pileup* p, int32_t contig, const hsgpu_partitions* parts, int32_t n_suspects,
                        const int32_t* suspect_pos, int32_t kept_capacity, int32_t* kept, int32_t* n_kept) {
    if (!p->ranked || contig < 0 || contig >= p->n_contigs) return HSGPU_ERR_STATE;
    const mock_contig* c = &p->c[contig];
    int32_t* all = (int32_t*)malloc(((size_t)c->L + 1) * sizeof(int32_t));
    const int32_t n = filter_contig(c, parts, n_suspects, suspect_pos, all);
    *n_kept = n;
    if (n > kept_capacity) {
        free(all);
        return HSGPU_ERR_CAPACITY;
    }
    memcpy(kept, all, (size_t)n * sizeof(int32_t));
    free(all);
    return HSGPU_OK;
}

int hsgpu_partitions_set(hsgpu_pileup* p, const hsgpu_partitions* parts) {
    if (!p->ranked) return HSGPU_ERR_STATE;
    for (int32_t ci = 0; ci < p->n_contigs; ci++) {
        mock_contig* c = &p->c[ci];
        free_parts(c);
        const int32_t np = parts[ci].n_parts;
        const int64_t n = np > 0 ? parts[ci].part_off[np] : 0;
        c->n_parts = np;
        c->part_off = (int64_t*)calloc((size_t)np + 1, sizeof(int64_t));
        c->p_idx = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
        c->p_more = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
        c->p_less = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
        c->p_state = (int16_t*)malloc(((size_t)n + 1) * sizeof(int16_t));
        if (np > 0) {
            memcpy(c->part_off, parts[ci].part_off, ((size_t)np + 1) * sizeof(int64_t));
            memcpy(c->p_idx, parts[ci].read_idx, (size_t)n * sizeof(int32_t));
            memcpy(c->p_more, parts[ci].more, (size_t)n * sizeof(int32_t));
            memcpy(c->p_less, parts[ci].less, (size_t)n * sizeof(int32_t));
            memcpy(c->p_state, parts[ci].state, (size_t)n * sizeof(int16_t));
        }
    }
    p->have_parts = 1;
    return HSGPU_OK;
}

int hsgpu_robust_filter_all(hsgpu_pileup* p, int64_t capacity, int32_t* kept, int64_t* off) {
    if (!p->ranked || !p->have_parts) return HSGPU_ERR_STATE;
    off[0] = 0;
    int rc = HSGPU_OK;
    for (int32_t ci = 0; ci < p->n_contigs; ci++) {
        const mock_contig* c = &p->c[ci];
        const hsgpu_partitions parts = {c->n_parts, c->part_off, c->p_idx, c->p_state, c->p_more, c->p_less};
        int32_t* all = (int32_t*)malloc(((size_t)c->L + 1) * sizeof(int32_t));
        const int32_t n = filter_contig(c, &parts, c->n_suspects, c->suspect_pos, all);
        if (kept && off[ci] + n <= capacity) memcpy(kept + off[ci], all, (size_t)n * sizeof(int32_t));
        else if (kept) rc = HSGPU_ERR_CAPACITY;
        off[ci + 1] = off[ci] + n;
        free(all);
    }
    return rc;
}

/* integration/glue_separate_reads.cpp */
int hsgpu_read_pair_counts(hsgpu_ctx* ctx, int32_t n_reads, int32_t n_snps, const int64_t* snp_off, const uint32_t* read_idx,
                           const uint8_t* code, const uint8_t* ref_base, const uint8_t* second_base, int32_t* sim,
                           int32_t* diff) {
    (void)ctx;
    hso_read_pair_counts(n_reads, n_snps, snp_off, read_idx, code, ref_base, second_base, sim, diff);
    return HSGPU_OK;
}
