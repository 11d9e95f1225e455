// Host-side entry points that run the SAME ranking code the kernels use (rank.cuh compiled for the
// host), so that the tie-breaking logic can be fuzzed on a machine without a GPU. Not part of
// include/hsgpu.h; used by tests only.
#include "common.cuh"
#include "rank.cuh"

extern "C" {

// out = k0, k1, c0, c1, c2 of one column (codes in the reference's in-column order)
void hsgpu_debug_rank_column(const uint8_t* codes, int n, int32_t* out) {
    uint32_t cnt[256] = {0};
    uint8_t order[HS_RH_MAXKEYS];
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (cnt[codes[i]]++ == 0) order[m++] = codes[i];
    }
    HsRhTable t;
    hs_rh_new(t);
    for (int k = 0; k < m; k++) hs_rh_insert(t, order[k]);
    hs_rh_insert(t, 0);
    hs_rh_insert(t, 1);
    hs_rh_insert(t, 2);
    uint8_t it[HS_RH_MAXKEYS];
    uint32_t kc[HS_RH_MAXKEYS];
    const int k = hs_rh_iterate(t, it);
    for (int i = 0; i < k; i++) kc[i] = (cnt[it[i]] << 8) | it[i];
    hs_kc_std_sort(kc, k);
    out[0] = kc[0] & 0xff;
    out[1] = kc[1] & 0xff;
    out[2] = kc[0] >> 8;
    out[3] = kc[1] >> 8;
    out[4] = kc[2] >> 8;
}

int hsgpu_debug_rh_order(const uint8_t* keys, int n, uint8_t* out) {
    HsRhTable t;
    hs_rh_new(t);
    for (int k = 0; k < n; k++) hs_rh_insert(t, keys[k]);
    return hs_rh_iterate(t, out);
}

void hsgpu_debug_sort_desc(uint8_t* keys, int32_t* counts, int n) {
    uint32_t kc[256];
    for (int i = 0; i < n; i++) kc[i] = ((uint32_t)counts[i] << 8) | keys[i];
    hs_kc_std_sort(kc, n);
    for (int i = 0; i < n; i++) {
        keys[i] = kc[i] & 0xff;
        counts[i] = (int32_t)(kc[i] >> 8);
    }
}

}  // extern "C"
