// Host-side entry points that run the SAME ranking code the kernels use (rank.cuh compiled for the
// host), so that the tie-breaking logic can be fuzzed on a machine without a GPU. Not part of
// include/hsgpu.h; used by tests only.
#include "common.cuh"
#include "rank.cuh"

extern "C" {

struct HostAcc {
    const uint8_t* order;
    const uint32_t* cnt;
    HS_HD int key(int k) const { return order[k]; }
    HS_HD unsigned count(int key) const { return cnt[key]; }
};

// out = k0, k1, c0, c1, c2 of one column (codes in the reference's in-column order), out[5] = 1 when the
// literal replay was needed, 0 when the bucket table decided. mode 0 = fast path with fallback (what the
// kernels do), 1 = always literal, 2 = the slot-order path alone (falls back on the literal replay). out[5]: 0 = decided by the bucket / hash-bit tables, 2 = literal replay
// after both table passes gave up, 1 = forced literal.
void hsgpu_debug_rank_column(const uint8_t* codes, int n, int32_t* out, int mode) {
    static HsRankLut lut;
    static bool have = false;
    if (!have) {
        hs_build_rank_lut(lut);
        have = true;
    }
    uint32_t cnt[256] = {0};
    uint8_t order[HS_RH_MAXKEYS];
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (cnt[codes[i]]++ == 0) order[m++] = codes[i];
    }
    HostAcc acc{order, cnt};
    int k0 = 0, k1 = 0;
    unsigned c0 = 0, c1 = 0, c2 = 0;
    int lit = mode == 1 ? 1 : hs_rank_fast(acc, m, &lut, k0, k1, c0, c1, c2);
    if (mode == 2) lit = 2 * hs_rank_slotorder(acc, m, &lut, k0, k1, c0, c1, c2);  // slot-order path alone
    else if (lit && mode != 1) {
        lit = 2 * hs_rank_hashbits(acc, m, &lut, k0, k1, c0, c1, c2);  // 0 resolved, 2 literal
        if (lit) lit = 2 * hs_rank_slotorder(acc, m, &lut, k0, k1, c0, c1, c2);
    }
    if (lit) hs_rank_literal(acc, m, k0, k1, c0, c1, c2);
    out[0] = k0;
    out[1] = k1;
    out[2] = (int32_t)c0;
    out[3] = (int32_t)c1;
    out[4] = (int32_t)c2;
    out[5] = lit;
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
