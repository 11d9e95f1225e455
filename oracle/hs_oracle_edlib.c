/* TEST INFRASTRUCTURE ONLY -- see hs_oracle.h.
 *
 * Plain-C restatement of what edlibAlign returns (reference src/edlib/src/edlib.cpp:142-297), written
 * as a full O(m*n) dynamic program instead of Myers' banded bit-vectors: edit distance, all end
 * locations, start locations and the alignment path with edlib's tie-breaking. The restated rules:
 *
 *   - D[i][0] = i. Top row: NW/SHW D[0][j] = j (startHout = 1, :618), HW D[0][j] = 0.
 *   - HW/SHW (:549-706): best = min_j D[m][j]; candidates are the target positions j-1 for j = 1..n and,
 *     ONLY when the query length is not a multiple of 64, also j = 0 (position -1): edlib pads the
 *     query to a whole number of 64-bit blocks with W wildcard rows and reads column c-W from the
 *     padded bottom row (:665-681,690-702), so with W > 0 the boundary column becomes visible.
 *     HW clamps k to the query length (:565-567). All positions with the best score, ascending.
 *   - NW (:732-933): distance D[m][n], -1 when k < |n - m| or distance > k; end location n-1.
 *   - k < 0: edlib doubles k from 64 until found (:195-213) -> same as unbounded.
 *   - start locations (:226-259): HW: for each end location e != -1 the prefix problem on the reversed
 *     query against reversed target[0..e] in SHW mode with k = distance, taking the LAST position p,
 *     start = e - p; for e == -1 start = 0. NW/SHW: 0.
 *   - path (:265-283, obtainAlignmentTraceback :947-1146): NW alignment of the query against
 *     target[start0..end0], traced back from the bottom-right cell with priority up (1 = insertion)
 *     > left (2 = deletion) > diagonal (0 match / 3 mismatch), then reversed -- below edlib's 1 MiB switch
 *     (:1193-1195). At or above it obtainAlignmentHirschberg (:1236-1401): the target is halved, the score columns of
 *     the left half and of the reversed right half are compared, the FIRST query row (ascending) whose two scores
 *     add up to the distance is the split row (:1312-1323), then the row "-1" boundary, then the last row
 *     (:1325-1343); both parts recurse through obtainAlignment with their own scores and their own 1 MiB test.
 *     edlib computes the columns inside its Ukkonen band; every cell on an optimal path lies inside both bands, so
 *     the full-matrix columns used here give the same split (pinned against the vendored edlib by
 *     tests/test_oracle_edlib.py on random pairs in that regime).
 *   - empty query or target: the special case of :162-180.
 */
#include <stdlib.h>
#include <string.h>

#include "hs_oracle.h"

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* last row of the DP: out[j] = D[m][j], j = 0..n. free_start: D[0][j] = 0 */
static void last_row(const unsigned char* q, int m, const unsigned char* t, int n, int free_start, int rev, int* out) {
    int* prev = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    int* cur = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    for (int i = 0; i <= m; i++) prev[i] = i;
    out[0] = m;
    for (int j = 1; j <= n; j++) {
        unsigned char tc = rev ? t[n - j] : t[j - 1];
        cur[0] = free_start ? 0 : j;
        for (int i = 1; i <= m; i++) {
            unsigned char qc = rev ? q[m - i] : q[i - 1];
            int d = prev[i - 1] + (qc != tc);
            int u = cur[i - 1] + 1;
            int l = prev[i] + 1;
            cur[i] = imin(d, imin(u, l));
        }
        out[j] = cur[m];
        int* tmp = prev; prev = cur; cur = tmp;
    }
    free(prev);
    free(cur);
}

/* semi-global search over a last row: best score and positions (j-1), honouring k and the W quirk.
 * returns number of positions (0 when nothing <= k), writes best (-1 if none). */
static int semiglobal_positions(const int* row, int m, int n, int k, int clamp_k_to_m, int* best_out, int* pos) {
    const int W = (64 - m % 64) % 64;
    if (clamp_k_to_m) k = imin(k, m);
    int best = -1, np = 0;
    for (int j = (W > 0 ? 0 : 1); j <= n; j++) {
        int s = row[j];
        if (s <= k && (best == -1 || s <= best)) {
            if (s != best) { np = 0; best = s; k = best; }
            pos[np++] = j - 1;
        }
    }
    *best_out = best;
    return np;
}

/* last column of the NW matrix: out[i] = D[i][n], i = 0..m (rev: both sequences read backwards) */
static void nw_last_column(const unsigned char* q, int m, const unsigned char* t, int n, int rev, int* out) {
    int* prev = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    int* cur = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    for (int i = 0; i <= m; i++) prev[i] = i;
    for (int j = 1; j <= n; j++) {
        unsigned char tc = rev ? t[n - j] : t[j - 1];
        cur[0] = j;
        for (int i = 1; i <= m; i++) {
            unsigned char qc = rev ? q[m - i] : q[i - 1];
            cur[i] = imin(prev[i - 1] + (qc != tc), imin(cur[i - 1], prev[i]) + 1);
        }
        int* tmp = prev; prev = cur; cur = tmp;
    }
    memcpy(out, prev, sizeof(int) * (size_t)(m + 1));
    free(prev);
    free(cur);
}

/* obtainAlignment (:1168-1230): NW path of q against t with known score `best`; ops appended in order at out.
 * Returns 0, or 1 when no split row exists (edlib's EDLIB_STATUS_ERROR). */
static int nw_path(const unsigned char* q, int m, const unsigned char* t, int n, int best, uint8_t* out, int* len_out) {
    if (m == 0 || n == 0) { /* :1173-1180 */
        for (int i = 0; i < m + n; i++) out[i] = m == 0 ? 2 : 1;
        *len_out = m + n;
        return 0;
    }
    const long long blocks = (m + 63) / 64;
    if ((2ll * 8 + 4) * blocks * n + 2ll * 4 * n < 1024 * 1024) { /* traceback (:947-1146) */
        int* D = (int*)malloc(sizeof(int) * (size_t)(m + 1) * (size_t)(n + 1));
#define DD(i, j) D[(size_t)(i) * (n + 1) + (j)]
        for (int j = 0; j <= n; j++) DD(0, j) = j;
        for (int i = 1; i <= m; i++) {
            DD(i, 0) = i;
            for (int j = 1; j <= n; j++)
                DD(i, j) = imin(DD(i - 1, j - 1) + (q[i - 1] != t[j - 1]), imin(DD(i - 1, j), DD(i, j - 1)) + 1);
        }
        int i = m, j = n, len = 0;
        while (i > 0 || j > 0) {
            if (i > 0 && DD(i - 1, j) + 1 == DD(i, j)) { out[len++] = 1; i--; }
            else if (j > 0 && DD(i, j - 1) + 1 == DD(i, j)) { out[len++] = 2; j--; }
            else { out[len++] = DD(i - 1, j - 1) == DD(i, j) ? 0 : 3; i--; j--; }
        }
        for (int a = 0, b = len - 1; a < b; a++, b--) { uint8_t x = out[a]; out[a] = out[b]; out[b] = x; }
        *len_out = len;
        free(D);
#undef DD
        return 0;
    }
    /* Hirschberg (:1236-1401) */
    const int left_w = n / 2, right_w = n - left_w;
    int* L = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    int* R = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    nw_last_column(q, m, t, left_w, 0, L);            /* L[i] = D[i][left_w]: scoresLeft[r] = L[r + 1] */
    nw_last_column(q, m, t + left_w, right_w, 1, R);  /* R[i] = reversed D[i][right_w]: scoresRight[r] = R[m - r] */
    int row = -2, left_score = -1, right_score = -1;
    for (int r = 0; r <= m - 2; r++) { /* :1312-1323 */
        if (L[r + 1] + R[m - (r + 1)] == best) { row = r; left_score = L[r + 1]; right_score = R[m - (r + 1)]; break; }
    }
    if (row == -2 && left_w + R[m] == best) { row = -1; left_score = left_w; right_score = R[m]; }           /* :1325-1333 */
    if (row == -2 && L[m] + right_w == best) { row = m - 1; left_score = L[m]; right_score = right_w; }       /* :1334-1343 */
    free(L);
    free(R);
    if (row == -2) { *len_out = 0; return 1; }
    const int ul_h = row + 1;
    int l1 = 0, l2 = 0;
    int rc = nw_path(q, ul_h, t, left_w, left_score, out, &l1);
    if (!rc) rc = nw_path(q + ul_h, m - ul_h, t + left_w, right_w, right_score, out + l1, &l2);
    *len_out = l1 + l2;
    return rc;
}

int32_t hso_edlib_align(const char* query, int32_t m, const char* target, int32_t n, int32_t k, int32_t mode,
                        int32_t task, int32_t* edit_distance, int32_t* alphabet_length, int32_t* n_locations,
                        int32_t* end_locations, int32_t* start_locations, int32_t* alignment_length,
                        uint8_t* alignment) {
    const unsigned char* q = (const unsigned char*)query;
    const unsigned char* t = (const unsigned char*)target;
    *edit_distance = -1;
    *n_locations = 0;
    *alignment_length = 0;
    /* alphabet (:1422-1462) */
    int seen[256] = {0}, na = 0;
    for (int i = 0; i < m; i++) if (!seen[q[i]]) { seen[q[i]] = 1; na++; }
    for (int i = 0; i < n; i++) if (!seen[t[i]]) { seen[t[i]] = 1; na++; }
    *alphabet_length = na;
    if (m == 0 || n == 0) { /* :162-180 */
        if (mode == 0) { *edit_distance = imax(m, n); end_locations[0] = n - 1; *n_locations = 1; }
        else if (mode == 1 || mode == 2) { *edit_distance = m; end_locations[0] = -1; *n_locations = 1; }
        else return 1;
        return 0;
    }
    const int unbounded = k < 0;
    if (unbounded) k = 0x3fffffff;
    int* row = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    int* pos = (int*)malloc(sizeof(int) * (size_t)(n + 2));
    int best = -1, np = 0;
    if (mode == 1 || mode == 2) {
        last_row(q, m, t, n, mode == 2, 0, row);
        np = semiglobal_positions(row, m, n, k, mode == 2, &best, pos);
    } else {
        int kk = k;
        if (!(kk < abs(n - m))) {
            kk = imin(kk, imax(m, n));
            last_row(q, m, t, n, 0, 0, row);
            if (row[n] <= kk) best = row[n];
        }
        if (best >= 0) { pos[0] = n - 1; np = 1; }
    }
    *edit_distance = best;
    if (best < 0) { free(row); free(pos); return 0; }
    *n_locations = np;
    for (int i = 0; i < np; i++) end_locations[i] = pos[i];
    if (task >= 1) {
        for (int i = 0; i < np; i++) {
            if (mode != 2 || pos[i] == -1) { start_locations[i] = 0; continue; }
            int e = pos[i];
            int* rrow = (int*)malloc(sizeof(int) * (size_t)(e + 2));
            int* rpos = (int*)malloc(sizeof(int) * (size_t)(e + 3));
            int rbest, rn;
            last_row(q, m, t, e + 1, 0, 1, rrow); /* reversed query vs reversed target[0..e], SHW */
            rn = semiglobal_positions(rrow, m, e + 1, best, 0, &rbest, rpos);
            start_locations[i] = rn > 0 ? e - rpos[rn - 1] : 0;
            free(rrow);
            free(rpos);
        }
    }
    int status = 0;
    if (task == 2) {
        const int s0 = start_locations[0], e0 = end_locations[0];
        const int an = e0 - s0 + 1;
        int len = 0;
        status = nw_path(q, m, t + s0, an < 0 ? 0 : an, best, alignment, &len);
        *alignment_length = len;
    }
    free(row);
    free(pos);
    return status;
}
