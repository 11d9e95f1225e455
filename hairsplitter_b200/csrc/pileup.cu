// Pileup construction: the body of generate_msa (reference src/call_variants.cpp:163-364) on the GPU.
//
// Layout in HBM. The reference keeps a vector<Column> (column-major, (readIdx u32, code u8) pairs,
// ~5 B/cell, pointer-chasing). Here the pileup is a ragged, READ-major byte matrix: read r owns one
// contiguous row of codes, one byte per contig column it covers, addressed codes[row_base[r] + q]
// with row_base % 16 == 0, zero-padded to 16-byte boundaries in column space. Rows are written with
// fully coalesced stores by the warp that walks the read's CIGAR, need no read index per cell
// (1 B/cell instead of 5), and a 128-column tile of any row is eight aligned 16-byte vectors, which
// is what the column kernels (column.cu, contingency.cu) stage through shared memory. A per-tile
// index lists the reads overlapping each tile in ascending neighbour order, which is exactly the
// in-column order of the reference.
#include <algorithm>
#include <cstring>
#include <mutex>

#include <atomic>

#include "common.cuh"

static std::mutex g_upload_mutex[64];  // per device, see hsgpu_pileup_create

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };  // BAM ops (host side)

// ---- CIGAR on the device --------------------------------------------------------------------------
// One byte per op: len << 2 | kind, len <= 63, kind 0 = M/=/X, 1 = I, 2 = D, 3 = S/H -- the four classes
// generate_msa distinguishes (:226-342); N and P ops move neither cursor there and are dropped, longer ops are
// split. This is hsgpu_pileup_input.cigar8 as the caller gives it; the u32 / u16 forms are converted on the host
// (hs_cigar_to_bytes). Four ops sit in one 32-bit word, so lengths and kinds are handled four at a time.
#define HS_READ_IRREGULAR 1
#define HS_SUPER_TILES 4  // tiles per super-tile of the two-level tile index
#define HS_SUPER_COLS (HS_SUPER_TILES * HS_TILE)

// the aligned word holding cigar bytes [kb, kb+4) of a read whose ops are bytes [k0, k1): bytes outside are zeroed
// (a zero byte is an M op of length 0: it takes part in nothing)
__device__ __forceinline__ uint32_t cg_word(const uint8_t* __restrict__ cigar, int64_t kb, int64_t k0, int64_t k1) {
    uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(cigar + kb));
    if (kb < k0) w &= 0xffffffffu << (8 * (int)(k0 - kb));
    if (kb + 4 > k1) w &= (k1 - kb <= 0) ? 0u : (0xffffffffu >> (8 * (int)(kb + 4 - k1)));
    return w;
}

// ---- K1: contig span of every read (sum of M/=/X/D lengths, clipped at the contig end, :217) ----
// Also classifies the CIGAR: a read is REGULAR when its clips (S/H) only lead or trail the aligned part,
// which is what SAM allows; its leading clip length is the read offset of the first aligned base
// (S and H both advance the read cursor in the reference, :269-273). Anything else is IRREGULAR and
// goes through pileup_generic_kernel.
__global__ void __launch_bounds__(256) span_kernel(int64_t n_reads, const uint8_t* __restrict__ cigar,
                                                   const int64_t* __restrict__ cigar_off,
                                                   const int32_t* __restrict__ read_start,
                                                   const int32_t* __restrict__ read_contig,
                                                   const int32_t* __restrict__ contig_len,
                                                   int32_t* __restrict__ read_end, int64_t* __restrict__ row_alloc,
                                                   int32_t* __restrict__ read_tlead, uint8_t* __restrict__ read_flags,
                                                   unsigned long long* __restrict__ totals) {
    // the three totals are summed per CTA first: one global atomic per counter and CTA instead of one per read
    __shared__ unsigned long long s_tot[3];
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const bool valid = r < n_reads;
    const int64_t k0 = valid ? cigar_off[r] : 0, k1 = valid ? cigar_off[r + 1] : 0;
    const int64_t kbase = k0 & ~(int64_t)3;
    int sum = 0;
    long long first = k1, last = -1;  // first / last op that holds alignment positions (M, I, D with a length)
    int clips = 0;
    // four independent loads per lane in flight
    for (int64_t kb = kbase + 4 * lane; kb < k1; kb += 4 * 32 * 4) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t k = kb + 128 * j;
            w[j] = k < k1 ? cg_word(cigar, k, k0, k1) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t k = kb + 128 * j;
            const uint32_t lens = (w[j] >> 2) & 0x3f3f3f3fu;
            const uint32_t b0 = w[j] & 0x01010101u, b1 = (w[j] >> 1) & 0x01010101u;
            sum = (int)__dp4a(lens & ((b0 ^ 0x01010101u) * 255u), 0x01010101u, (unsigned int)sum);  // kinds 0 and 2 consume the contig
            const uint32_t nz = ((lens + 0x3f3f3f3fu) >> 6) & 0x01010101u;       // len > 0
            const uint32_t is_s = b0 & b1;
            const uint32_t al = nz & ~is_s;
            if (al) {
                const long long f = k + ((__ffs(al) - 1) >> 3), l = k + ((31 - __clz(al)) >> 3);
                if (f < first) first = f;
                if (l > last) last = l;
            }
            if (nz & is_s) clips = 1;
        }
    }
    long long sum64 = hs_warp_sum64((long long)sum);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
    }
    long long tlead = 0;
    int irregular = 0;
    // the second pass places the clips; most reads have none
    if (__any_sync(0xffffffffu, clips))
    for (int64_t kb = kbase + 4 * lane; kb < k1; kb += 128) {
        const uint32_t w = cg_word(cigar, kb, k0, k1);
        if (((w & (w >> 1)) & 0x01010101u) == 0) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b = (w >> (8 * j)) & 0xffu;
            if ((b & 3u) == 3u && (b >> 2) > 0) {
                if (kb + j < first) tlead += b >> 2;
                else if (kb + j < last) irregular = 1;
            }
        }
    }
    tlead = hs_warp_sum64(tlead);
    irregular = __any_sync(0xffffffffu, irregular) || tlead > 0x3fffffff;
    if (valid && lane == 0) {
        const int L = contig_len[read_contig[r]];
        const int start = read_start[r];
        long long end = start;
        if (start < L) end = (start + sum64 < (long long)L) ? start + sum64 : L;
        read_end[r] = (int)end;
        long long alloc = 0;
        if (end > start) alloc = ((end + HS_ALIGN - 1) & ~(long long)(HS_ALIGN - 1)) - (start & ~(HS_ALIGN - 1));
        row_alloc[r] = alloc;
        read_tlead[r] = (int32_t)tlead;
        read_flags[r] = irregular ? HS_READ_IRREGULAR : 0;
        if (irregular) atomicAdd(&s_tot[2], 1ull);
        if (end > start) {  // exact sizes of the two index levels (tile_index_kernel), known before they are built
            atomicAdd(&s_tot[0], (unsigned long long)((end - 1) / HS_TILE - start / HS_TILE + 1));
            atomicAdd(&s_tot[1], (unsigned long long)((end - 1) / HS_SUPER_COLS - start / HS_SUPER_COLS + 1));
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 && s_tot[threadIdx.x])
        atomicAdd(totals + (threadIdx.x == 0 ? 1 : (threadIdx.x == 1 ? 4 : 2)), s_tot[threadIdx.x]);
}

struct PileupArgs {
    int64_t n_reads;
    const int32_t* contig_len;
    const uint32_t* contig_bases;
    const int64_t* contig_word_off;
    const int32_t* read_contig;
    const uint32_t* read_bases;
    const int64_t* read_word_off;
    const int32_t* read_len;
    const uint8_t* cigar;   // one byte per op (see above)
    const int64_t* cigar_off;
    const int32_t* read_start;
    const uint8_t* read_strand;
    const int32_t* read_end;
    const int32_t* read_tlead;
    const uint8_t* read_flags;
    const int64_t* row_off;
    int64_t* row_base;
    uint8_t* codes;
    unsigned long long* stats;
    unsigned int* next_read;  // work counters of the persistent warps, one per length class
    int ops_long, ops_mid;    // CIGAR ops from which a read counts as long / medium
    int k_one, k_four;        // 1 and 4, unknown to the assembler: see PW_STEP
};

// ---- K3: the CIGAR walk. Persistent warps, one read at a time, one LANE per run of 30 positions. ----
// The alignment of a read is the sequence of its M/=/X, I and D positions (:226-342). A warp takes the
// CIGAR in windows of up to 192 ops (six per lane, kept in registers), prefix-scans their lengths and
// turns the window's first 960 positions into two bit masks in shared memory (is-insertion,
// is-deletion; warp-private, set with shared-memory atomics). Prefix popcounts of the masks give every
// lane the contig column and the read offset at which its run of 30 positions starts, so no lane ever
// searches for "its" op. Each lane then walks 2 warm-up positions (they rebuild the 3-mer context, the
// two symbols pushed before, :234-238) and its 30 positions with a branch-free step of ~13
// instructions: the 16 next read symbols and contig symbols sit in two registers that all lanes refill
// together, the op kind of step u is bit u of the two masks. Codes go to a warp-private staging row in
// shared memory and leave for HBM as whole aligned 16-byte vectors: every row byte is written exactly
// once, fully coalesced.
#define PW_WARPS 8
#define PW_OPL 6                  // CIGAR ops per lane and window (byte-sized: four in one register, two in another)
#define PW_NOPS (32 * PW_OPL)
#define PW_P 30                   // positions per lane and window (+2 warm-up = 32 steps)
#define PW_E (32 * PW_P)
#define PW_BUF 1024               // staging row: a carried partial vector + one window of columns (<= 16 + 960 + 16)
enum { CG_M = 0, CG_I = 1, CG_D = 2, CG_S = 3 };  // kinds of the byte-sized CIGAR ops

// CHECKED = false: the caller knows that every word touched lies inside the sequence (windows away from the ends
// of the read and of the contig: all but the first and last window of a read)
template <bool CHECKED>
__device__ __forceinline__ uint32_t pw_word(const uint32_t* __restrict__ w, int i, int n) {
    if (!CHECKED) return __ldg(w + i);
    return (i >= 0 && i < n) ? __ldg(w + i) : 0u;
}
// sixteen 2-bit symbols starting at base i (i may be negative or run past the end: zeros)
template <bool CHECKED>
__device__ __forceinline__ uint32_t pw_window(const uint32_t* __restrict__ w, int nw, int i) {
    const int wi = i >> 4;
    return __funnelshift_r(pw_word<CHECKED>(w, wi, nw), pw_word<CHECKED>(w, wi + 1, nw), (i & 15) * 2);
}
// the 16 read symbols of alignment read offsets tp .. tp+15 (reverse strand: complement, read backwards)
template <bool CHECKED>
__device__ __forceinline__ uint32_t pw_read_window(const uint32_t* __restrict__ rb, int nw, int rlen, int tp, int strand) {
    uint32_t x;
    int nvalid;
    if (strand) {
        x = pw_window<CHECKED>(rb, nw, tp);
        nvalid = rlen - tp;
    } else {
        const int j = rlen - 1 - tp;  // first base wanted; the window is bases j-15 .. j, reversed
        x = pw_window<CHECKED>(rb, nw, j - 15);
        x = __brev(x);
        x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
        x = ~x;
        nvalid = j + 1;
    }
    // a CIGAR longer than the read is malformed (the reference reads past the string); those symbols are 0
    if (CHECKED && nvalid < 16) x = (nvalid <= 0) ? 0u : (x & ((1u << (2 * nvalid)) - 1u));
    return x;
}

// The same windows with symbol 0 in the TOP two bits, for the fast walk: taking a symbol is a shift by 30, moving on
// is a multiplication by 4 -- IMAD on the FMA pipe instead of SHF on the ALU pipe, which bounds the kernel.
__device__ __forceinline__ uint32_t pw_rev2(uint32_t x) {
    x = __brev(x);
    return ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
}
template <bool CHECKED>
__device__ __forceinline__ uint32_t pw_window_top(const uint32_t* __restrict__ w, int nw, int i) {
    return pw_rev2(pw_window<CHECKED>(w, nw, i));
}
template <bool CHECKED>
__device__ __forceinline__ uint32_t pw_read_window_top(const uint32_t* __restrict__ rb, int nw, int rlen, int tp, int strand) {
    uint32_t x;
    int nvalid;
    if (strand) {
        x = pw_rev2(pw_window<CHECKED>(rb, nw, tp));
        nvalid = rlen - tp;
    } else {
        const int j = rlen - 1 - tp;  // bases j-15 .. j: the last one is the first symbol, and already on top
        x = ~pw_window<CHECKED>(rb, nw, j - 15);
        nvalid = j + 1;
    }
    if (CHECKED && nvalid < 16) x = (nvalid <= 0) ? 0u : (x & ~((1u << (32 - 2 * nvalid)) - 1u));
    return x;
}

// set bits [a, a + n) of a warp-private bit array in shared memory, 0 < n <= 63 (one CIGAR op): at most three words.
// Straight-line code on purpose: a loop around the atomics makes the compiler insert YIELDs and a non-reconvergent
// barrier, and the warp then walks the rest of the window in two halves (measured: 16 active lanes per instruction).
__device__ __forceinline__ void pw_set_bits(unsigned int* __restrict__ bits, int a, int n) {
    const int s = a & 31;
    unsigned int* const w = bits + (a >> 5);
    const unsigned long long m = (1ull << n) - 1ull;  // n <= 63
    const unsigned long long lo = m << s;             // bits 0..63 of the shifted run
    const unsigned int w2 = s ? (unsigned int)(m >> (64 - s)) : 0u;  // bits 64.. (s + n <= 94)
    atomicOr(w, (unsigned int)lo);
    const unsigned int w1 = (unsigned int)(lo >> 32);
    if (w1) atomicOr(w + 1, w1);
    if (w2) atomicOr(w + 2, w2);
}

// the PW_OPL ops [kop + PW_OPL*lane, ...) of a read: four in `lo`, two in the low half of `hi` (bytes past nops
// read as zero: empty M ops)
__device__ __forceinline__ void pw_load_ops(const uint8_t* __restrict__ cig, int kop, int nops, int lane, uint32_t& lo,
                                            uint32_t& hi) {
    const uint8_t* p = cig + kop + PW_OPL * lane;
    const int sh = (int)(reinterpret_cast<uintptr_t>(p) & 3);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(p - sh);
    const int nv = min(PW_OPL, nops - kop - PW_OPL * lane);  // valid bytes from p on
    // three aligned words cover the six bytes; a word is read only if it holds a valid byte
    const uint32_t w0 = nv > 0 ? __ldg(wp) : 0u;
    const uint32_t w1 = nv + sh > 4 ? __ldg(wp + 1) : 0u;
    const uint32_t w2 = nv + sh > 8 ? __ldg(wp + 2) : 0u;
    lo = __funnelshift_r(w0, w1, 8 * sh);
    hi = __funnelshift_r(w1, w2, 8 * sh) & 0xffffu;
    if (nv < PW_OPL) {
        if (nv <= 0) { lo = 0; hi = 0; }
        else if (nv <= 4) { hi = 0; if (nv < 4) lo &= (1u << (8 * nv)) - 1u; }
        else hi &= (1u << (8 * (nv - 4))) - 1u;
    }
}

// One step of the FAST walk, as PTX so that it stays the 14 predicated instructions it is meant to be
// (the compiler otherwise re-derives the column from the mask bits and branches around the store).
//   pI/pD = this step is an insertion / a deletion (bit BIT of the masks)
//   sym   = '-' for a deletion, else the next read symbol;  code = ctx + 25*sym, ctx = '!' + 5*b(-2) + b(-1)  (:238,287)
//   a mismatch is counted for M and D positions ('-' never matches, :254-256,305); insertions are counted from the mask
//   p1x   = 5*b(-1) + '!' of the next step
// The kernel is bound by the ALU pipe (LOP3 / SHF / ISETP / SEL / IADD: 64 lanes per clock and SM), not by issue
// slots, while the FMA pipe is almost idle. The windows keep their next symbol on top, so that advancing is a
// multiplication by 4; the +1 of the staging address and of the mismatch count are multiply-adds by 1. %8 = 1 and
// %9 = 4 come from the kernel parameters so that the assembler cannot turn them back into adds and shifts:
// 6 ALU + 7 FMA + 1 LSU instructions per step instead of 10 + 3 + 1.
#define PW_STEP_HEAD(BIT)                                   \
    "and.b32 t, %6, " #BIT ";\n\t"                          \
    "setp.ne.u32 pI, t, 0;\n\t"                             \
    "and.b32 t, %7, " #BIT ";\n\t"                          \
    "setp.ne.u32 pD, t, 0;\n\t"                             \
    "shr.u32 b, %0, 30;\n\t"                                \
    "selp.b32 sym, 4, b, pD;\n\t"
#define PW_STEP_EMIT                                        \
    "shr.u32 c, %1, 30;\n\t"                                \
    "mad.lo.s32 code, sym, 25, %3;\n\t"                     \
    "@!pI st.shared.u8 [%2], code;\n\t"                     \
    "setp.ne.and.s32 pm, sym, c, !pI;\n\t"                  \
    "@pm mad.lo.s32 %5, %8, %8, %5;\n\t"
#define PW_STEP_TAIL                                        \
    "@!pD mul.lo.u32 %0, %0, %9;\n\t"                       \
    "@!pI mul.lo.u32 %1, %1, %9;\n\t"                       \
    "@!pI mad.lo.s32 %2, %8, %8, %2;\n\t"                   \
    "add.s32 %3, %4, sym;\n\t"                              \
    "mad.lo.s32 %4, sym, 5, 33;\n\t"
#define PW_STEP_REGS "{\n\t.reg .pred pI, pD, pm;\n\t.reg .b32 t, b, c, sym, code;\n\t"
#define PW_STEP_OPS : "+r"(rw), "+r"(cw), "+r"(qa), "+r"(ctx), "+r"(p1x), "+r"(dist) : "r"(mI), "r"(mD), "r"(k_one), "r"(k_four)
#define PW_WARM(BIT) asm volatile(PW_STEP_REGS PW_STEP_HEAD(BIT) PW_STEP_TAIL "}" PW_STEP_OPS)
#define PW_STEP(BIT) asm volatile(PW_STEP_REGS PW_STEP_HEAD(BIT) PW_STEP_EMIT PW_STEP_TAIL "}" PW_STEP_OPS)

// the 32 steps of one lane of a FAST window (a full window that cannot reach the contig end: every lane
// emits exactly PW_P positions, nothing is predicated on counts or on the column)
template <bool CHECKED>
__device__ __forceinline__ void pw_walk_fast(const uint32_t mI, const uint32_t mD, const int lane, const int carry_ctx,
                                             const int carry_p1x, const uint32_t* __restrict__ rb, const int nrw,
                                             const int rlen, const int strand, int tp, const uint32_t* __restrict__ cb,
                                             const int ncw, const int q, unsigned int qa, int& ctx_out, int& p1x_out,
                                             unsigned int& dist, const int k_one, const int k_four) {
    int ctx = 0, p1x = 0;  // rebuilt by the two warm-up steps
    uint32_t rw = pw_read_window_top<CHECKED>(rb, nrw, rlen, tp, strand);
    uint32_t cw = pw_window_top<CHECKED>(cb, ncw, q);
    PW_WARM(0x1); PW_WARM(0x2);
    if (lane == 0) {  // lane 0 warmed up on two virtual positions: its context is the carry
        ctx = carry_ctx;
        p1x = carry_p1x;
    }
    PW_STEP(0x4); PW_STEP(0x8); PW_STEP(0x10); PW_STEP(0x20); PW_STEP(0x40); PW_STEP(0x80);
    PW_STEP(0x100); PW_STEP(0x200); PW_STEP(0x400); PW_STEP(0x800); PW_STEP(0x1000); PW_STEP(0x2000);
    PW_STEP(0x4000); PW_STEP(0x8000);
    const int lowI = __popc(mI & 0xffffu), lowD = __popc(mD & 0xffffu);
    rw = pw_read_window_top<CHECKED>(rb, nrw, rlen, tp + 16 - lowD, strand);
    cw = pw_window_top<CHECKED>(cb, ncw, q + 16 - lowI);
    PW_STEP(0x10000); PW_STEP(0x20000); PW_STEP(0x40000); PW_STEP(0x80000); PW_STEP(0x100000); PW_STEP(0x200000);
    PW_STEP(0x400000); PW_STEP(0x800000); PW_STEP(0x1000000); PW_STEP(0x2000000); PW_STEP(0x4000000);
    PW_STEP(0x8000000); PW_STEP(0x10000000); PW_STEP(0x20000000); PW_STEP(0x40000000); PW_STEP(0x80000000);
    ctx_out = ctx;
    p1x_out = p1x;
}

// the general walk: partial windows and windows that reach the contig end
__device__ __forceinline__ void pw_walk_slow(const uint32_t mI, const uint32_t mD, const int cnt, const int lane,
                                             const int carry_ctx, const int carry_p1x, const uint32_t* __restrict__ rb,
                                             const int nrw, const int rlen, const int strand, int tp,
                                             const uint32_t* __restrict__ cb, const int ncw, int q, const int L,
                                             uint8_t* __restrict__ out, int& ctx_out, int& p1x_out, unsigned int& dist,
                                             unsigned int& alen) {
    int ctx = 0, p1x = 0;
#pragma unroll 1
    for (int blk = 0; blk < 2; blk++) {
        uint32_t rw = pw_read_window<true>(rb, nrw, rlen, tp, strand);
        uint32_t cw = pw_window<true>(cb, ncw, q);
#pragma unroll 4
        for (int uu = 0; uu < 16; uu++) {
            const int u = 16 * blk + uu;
            if (u == 2 && lane == 0) {
                ctx = carry_ctx;
                p1x = carry_p1x;
            }
            const bool iI = (mI >> u) & 1u, iD = (mD >> u) & 1u;
            const int b = (int)(rw & 3u), c = (int)(cw & 3u);
            const int sym = iD ? 4 : b;
            const bool mine = u - 2 < cnt;
            if (u >= 2 && mine && q < L) {
                if (!iI) out[q] = (uint8_t)(ctx + 25 * sym);
                alen++;
                if (iI || sym != c) dist++;  // :254-256, :305, :337-338
            }
            if (mine) {  // the context freezes after the lane's last position (it may be the carry)
                ctx = p1x + sym;
                p1x = 5 * sym + HS_CODE0;
            }
            if (!iD) { rw >>= 2; tp++; }
            if (!iI) { cw >>= 2; q++; }
        }
    }
    ctx_out = ctx;
    p1x_out = p1x;
}

__global__ void __launch_bounds__(32 * PW_WARPS, 4) pileup_kernel(PileupArgs a) {
    __shared__ unsigned int s_bits[PW_WARPS][2][32];
    __shared__ __align__(16) uint8_t s_buf[PW_WARPS][PW_BUF];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned int* __restrict__ sI = s_bits[wid][0];
    unsigned int* __restrict__ sD = s_bits[wid][1];
    uint8_t* __restrict__ buf = s_buf[wid];
    // Longest reads first (three sweeps over the read list, each with its own counter): a warp walks its read
    // alone, so a 60 kb read picked up late would run on long after every other warp has left.
    for (int sweep = 0; sweep < 3; sweep++)
    for (;;) {
        unsigned int grab = 0;
        if (lane == 0) grab = atomicAdd(a.next_read + sweep, 1u);
        const int64_t r = __shfl_sync(0xffffffffu, grab, 0);
        if (r >= a.n_reads) break;  // whole warps leave together; only __syncwarp is used
        {
            const int n_ops = (int)(a.cigar_off[r + 1] - a.cigar_off[r]);
            if ((n_ops >= a.ops_long ? 0 : (n_ops >= a.ops_mid ? 1 : 2)) != sweep) continue;
        }
        const int start = a.read_start[r];
        const int64_t row_base = a.row_off[r] - (int64_t)(start & ~(HS_ALIGN - 1));
        if (lane == 0) a.row_base[r] = row_base;
        if (a.read_flags[r] & HS_READ_IRREGULAR) continue;  // pileup_generic_kernel writes this row
        const int c = a.read_contig[r];
        const int L = a.contig_len[c];
        const uint32_t* __restrict__ cb = a.contig_bases + a.contig_word_off[c];
        const int ncw = (L + 15) >> 4;
        const uint32_t* __restrict__ rb = a.read_bases + a.read_word_off[r];
        const int rlen = a.read_len[r];
        const int nrw = (rlen + 15) >> 4;
        const int strand = a.read_strand[r];
        const int end = a.read_end[r];
        uint8_t* __restrict__ row = a.codes + row_base;
        const uint8_t* __restrict__ cig = a.cigar + a.cigar_off[r];
        const int nops = (int)(a.cigar_off[r + 1] - a.cigar_off[r]);
        int kop = 0;                   // first op of the window
        int off0 = 0;                  // positions of op kop that earlier windows already consumed
        int q0 = start, t0 = a.read_tlead[r];  // contig column / read offset of the window's first position
        int carry_ctx = HS_CODE0 + 5 * 1 + 2, carry_p1x = 5 * 2 + HS_CODE0;  // context 'A','C','G' (:212-214)
        int bufbase = start & ~(HS_ALIGN - 1);               // column of buf[0]
        unsigned int dist = 0, alen = 0;    // per lane
        unsigned int udist = 0, ualen = 0;  // per warp (same value in every lane)
        if (lane < HS_ALIGN) buf[lane] = 0;  // the pad in front of the first cell
        __syncwarp();
        // the window's ops live in two registers (six byte-sized ops per lane); the next window's are requested
        // before the walk of this one
        uint32_t oplo, ophi;
        pw_load_ops(cig, kop, nops, lane, oplo, ophi);
        while (kop < nops && q0 < L) {
            // ---- the window's ops, four at a time: lengths, kinds, positions taken (S/H take none) ------------
            const uint32_t len_lo = (oplo >> 2) & 0x3f3f3f3fu, len_hi = (ophi >> 2) & 0x3f3f3f3fu;
            const uint32_t s_lo = oplo & (oplo >> 1) & 0x01010101u, s_hi = ophi & (ophi >> 1) & 0x01010101u;
            uint32_t pl_lo = len_lo & ~(s_lo * 255u), pl_hi = len_hi & ~(s_hi * 255u);
            if (lane == 0) pl_lo -= min((uint32_t)off0, pl_lo & 0xffu);  // op kop was partly consumed by the last window
            const int es_lo = (int)__dp4a(pl_lo, 0x01010101u, 0u), es = (int)__dp4a(pl_hi, 0x01010101u, (unsigned int)es_lo);
            // exclusive prefix of the positions inside each word (byte k = positions of the bytes below k; <= 189)
            const uint32_t pre_lo = pl_lo * 0x01010100u, pre_hi = pl_hi * 0x01010100u;
            const int ei = hs_warp_incl_scan(es, lane);
            const int ebase = ei - es;
            const int Etot = __shfl_sync(0xffffffffu, ei, 31);
            const int nload = min(PW_NOPS, nops - kop);
            const int Enew = min(Etot, PW_E);
            // ---- where the next window starts; its ops are requested now and used after the walk --------
            int kop_next = kop + nload, off0_next = 0;
            if (Enew < Etot) {
                int myslot = -1, myd = 0;
                if (es > 0 && ebase <= Enew && Enew < ei) {  // exactly one lane: the op that holds position Enew
                    const int rel = Enew - ebase;
                    const bool hi = rel >= es_lo;
                    const int rel2 = hi ? rel - es_lo : rel;
                    const uint32_t pl = hi ? pl_hi : pl_lo, pre = hi ? pre_hi : pre_lo;
                    int j = 0;
#pragma unroll
                    for (int k = 1; k < 4; k++)
                        if ((int)((pre >> (8 * k)) & 0xffu) <= rel2 && ((pl >> (8 * k)) & 0xffu) > 0) j = k;
                    // the last byte whose range starts at or before rel2 and is not empty holds it
                    myslot = PW_OPL * lane + (hi ? 4 : 0) + j;
                    myd = rel2 - (int)((pre >> (8 * j)) & 0xffu);
                }
                const unsigned int who = __ballot_sync(0xffffffffu, myslot >= 0);
                const int src = __ffs(who) - 1;
                const int slot = __shfl_sync(0xffffffffu, myslot, src);
                const int d = __shfl_sync(0xffffffffu, myd, src);
                off0_next = (slot == 0 ? off0 : 0) + d;
                kop_next = kop + slot;
            }
            const uint32_t cur_lo = oplo, cur_hi = ophi;
            pw_load_ops(cig, kop_next, nops, lane, oplo, ophi);
            kop = kop_next;
            off0 = off0_next;
            if (Etot == 0) continue;  // nothing but clips / padding
            const bool full = Enew == PW_E;
            const int P = full ? PW_P : max(2, (Enew + 31) >> 5);
            // ---- insertion / deletion masks over window positions (bit 0,1 = virtual warm-up of lane 0) ---
            sI[lane] = 0;
            sD[lane] = 0;
            __syncwarp();
            if (ebase < Enew) {
                // kind 1 = I, kind 2 = D: exactly one of the two low bits
                const uint32_t id_lo = (cur_lo ^ (cur_lo >> 1)) & 0x01010101u, id_hi = (cur_hi ^ (cur_hi >> 1)) & 0x01010101u;
#pragma unroll
                for (int j = 0; j < PW_OPL; j++) {
                    const uint32_t idm = j < 4 ? id_lo : id_hi, pl = j < 4 ? pl_lo : pl_hi, pre = j < 4 ? pre_lo : pre_hi;
                    const uint32_t cur = j < 4 ? cur_lo : cur_hi;
                    const int sh = 8 * (j & 3);
                    const int ln = (int)((pl >> sh) & 0xffu);
                    if (((idm >> sh) & 1u) && ln > 0) {
                        const int pa = ebase + (j < 4 ? 0 : es_lo) + (int)((pre >> sh) & 0xffu);
                        const int n = min(ln, Enew - pa);
                        if (n > 0) pw_set_bits(((cur >> sh) & 1u) ? sI : sD, pa + 2, n);
                    }
                }
            }
            __syncwarp();
            const unsigned int wI = sI[lane], wD = sD[lane];
            const int pk = __popc(wI) | (__popc(wD) << 16);
            const int pki = hs_warp_incl_scan(pk, lane);
            const int tot = __shfl_sync(0xffffffffu, pki, 31);
            const int totI = tot & 0xffff, totD = tot >> 16;
            const int ws = lane * P;  // window position of the lane's first warm-up step
            const int wi = ws >> 5, bit = ws & 31;
            const unsigned int loI = __shfl_sync(0xffffffffu, wI, wi), hiI = __shfl_sync(0xffffffffu, wI, (wi + 1) & 31);
            const unsigned int loD = __shfl_sync(0xffffffffu, wD, wi), hiD = __shfl_sync(0xffffffffu, wD, (wi + 1) & 31);
            const int pre = __shfl_sync(0xffffffffu, pki - pk, wi);
            const unsigned int below = (1u << bit) - 1u;
            const int preI = (pre & 0xffff) + __popc(loI & below), preD = (pre >> 16) + __popc(loD & below);
            const uint32_t mI = __funnelshift_r(loI, hiI, bit), mD = __funnelshift_r(loD, hiD, bit);
            const int q = q0 - 2 + ws - preI;   // the two virtual positions count as plain matches
            const int tp = t0 - 2 + ws - preD;
            const int cnt = max(0, min(P, Enew - ws));
            const int qend = q0 + Enew - totI;
            int ctx, p1x;
            if (full && qend <= L) {
                const unsigned int qa = (unsigned int)__cvta_generic_to_shared(buf) + (unsigned int)(q - bufbase);
                // every window of 16 symbols the lanes fetch stays inside the read and the contig (64 bases of margin:
                // a fetch reaches at most 47 bases past the position it is made for)
                const bool inside = q0 >= 64 && qend + 64 < L && t0 >= 64 && t0 + Enew + 64 < rlen;
                if (inside) pw_walk_fast<false>(mI, mD, lane, carry_ctx, carry_p1x, rb, nrw, rlen, strand, tp, cb, ncw, q, qa, ctx, p1x, dist, a.k_one, a.k_four);
                else pw_walk_fast<true>(mI, mD, lane, carry_ctx, carry_p1x, rb, nrw, rlen, strand, tp, cb, ncw, q, qa, ctx, p1x, dist, a.k_one, a.k_four);
                ualen += Enew;
                udist += totI;  // deletions count in the walk ('-' never equals the contig base)
            } else {
                pw_walk_slow(mI, mD, cnt, lane, carry_ctx, carry_p1x, rb, nrw, rlen, strand, tp, cb, ncw, q, L,
                             buf - bufbase, ctx, p1x, dist, alen);
            }
            // the lane that pushed the window's last symbol: (Enew - 1) / P without the division
            const int last = full ? (PW_E - 1) / PW_P : 31 - __clz(__ballot_sync(0xffffffffu, cnt > 0));
            carry_ctx = __shfl_sync(0xffffffffu, ctx, last);
            carry_p1x = __shfl_sync(0xffffffffu, p1x, last);
            q0 = qend;
            t0 += Enew - totD;
            // ---- flush the complete 16-byte vectors, keep the partial one ----------------------------
            __syncwarp();
            const int wr_end = min(q0, L);
            const int nvec = ((wr_end & ~(HS_ALIGN - 1)) - bufbase) >> 4;
            for (int v = lane; v < nvec; v += 32)
                *reinterpret_cast<uint4*>(row + bufbase + 16 * v) = *reinterpret_cast<const uint4*>(buf + 16 * v);
            uint8_t keep = 0;
            if (lane < HS_ALIGN) keep = buf[16 * nvec + lane];
            __syncwarp();
            if (lane < HS_ALIGN) buf[lane] = keep;
            bufbase += 16 * nvec;
            __syncwarp();
        }
        if (end > start && (end & (HS_ALIGN - 1))) {  // the last, zero-padded vector
            const int o = end - bufbase;              // 0 < o < 16: every complete vector has been flushed
            if (lane >= o && lane < HS_ALIGN) buf[lane] = 0;
            __syncwarp();
            if (lane < 4) reinterpret_cast<uint32_t*>(row + bufbase)[lane] = reinterpret_cast<const uint32_t*>(buf)[lane];
        }
        __syncwarp();
        long long d64 = hs_warp_sum64((long long)dist) + udist, a64 = hs_warp_sum64((long long)alen) + ualen;
        if (lane == 0) {
            if (d64) atomicAdd(a.stats + 3 * c + 0, (unsigned long long)d64);
            if (a64) atomicAdd(a.stats + 3 * c + 1, (unsigned long long)a64);
            if (end > start) atomicAdd(a.stats + 3 * c + 2, (unsigned long long)(end - start));
        }
    }
}

// ---- K3b: generic walk for irregular CIGARs (clips between aligned parts). One warp per read; the ----
// expanded alignment is processed 32 positions at a time, each lane locating its op by binary search in
// shared memory, context via ballot + shuffle. Slow, exact for any op sequence.
__global__ void __launch_bounds__(256) pileup_generic_kernel(PileupArgs a) {
    __shared__ int s_e[8][33];
    __shared__ int s_q[8][32];
    __shared__ int s_t[8][32];
    __shared__ int s_ty[8][32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + wid;
    if (r >= a.n_reads) return;  // whole warps leave together; only __syncwarp is used below
    if (!(a.read_flags[r] & HS_READ_IRREGULAR)) return;
    const int c = a.read_contig[r];
    const int L = a.contig_len[c];
    const uint32_t* __restrict__ cb = a.contig_bases + a.contig_word_off[c];
    const uint32_t* __restrict__ rb = a.read_bases + a.read_word_off[r];
    const int rlen = a.read_len[r];
    const int strand = a.read_strand[r];
    const int start = a.read_start[r];
    const int end = a.read_end[r];
    const int64_t row_base = a.row_off[r] - (int64_t)(start & ~(HS_ALIGN - 1));
    if (lane == 0) a.row_base[r] = row_base;
    uint8_t* __restrict__ row = a.codes + row_base;
    if (end > start) {  // zero the alignment pads so that tiles can be staged with whole 16-byte vectors
        for (int q = (start & ~(HS_ALIGN - 1)) + lane; q < start; q += 32) row[q] = 0;
        const int b = (end + HS_ALIGN - 1) & ~(HS_ALIGN - 1);
        for (int q = end + lane; q < b; q += 32) row[q] = 0;
    }
    const int64_t k0 = a.cigar_off[r], k1 = a.cigar_off[r + 1];
    int q = start, t = 0;
    int carry1 = 2, carry2 = 1;  // context 'A','C','G': most recent = G, before it C (:212-214)
    unsigned int dist = 0, alen = 0;
    for (int64_t kb = k0; kb < k1 && q < L; kb += 32) {
        const int64_t k = kb + lane;
        const uint32_t op = (k < k1) ? (uint32_t)__ldg(a.cigar + k) : 0u;  // a zero byte: an M op without positions
        const int len = (int)(op >> 2), ty = (int)(op & 3u);
        const bool cq = ty == CG_M || ty == CG_D, ct = ty != CG_D;
        const int qa = cq ? len : 0, ta = ct ? len : 0, ea = len;
        const int qi = hs_warp_incl_scan(qa, lane), ti = hs_warp_incl_scan(ta, lane), ei = hs_warp_incl_scan(ea, lane);
        s_q[wid][lane] = q + qi - qa;
        s_t[wid][lane] = t + ti - ta;
        s_e[wid][lane] = ei - ea;
        s_ty[wid][lane] = ty;
        const int E = __shfl_sync(0xffffffffu, ei, 31);
        __syncwarp();
        for (int e0 = 0; e0 < E; e0 += 32) {
            const int e = e0 + lane;
            int lo = 0, hi = 31;
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_e[wid][mid] <= e) lo = mid; else hi = mid - 1;
            }
            const int off = e - s_e[wid][lo];
            const int oty = s_ty[wid][lo];
            const bool ocq = oty == CG_M || oty == CG_D, oct = oty != CG_D;
            const int qp = s_q[wid][lo] + (ocq ? off : 0);
            const int tp = s_t[wid][lo] + (oct ? off : 0);
            const bool active = (e < E) && (qp < L);
            const bool push = active && (ocq || oty == CG_I);
            int sym = 0;
            if (push) {
                if (oty == CG_D) sym = 4;
                else if (tp < rlen) {  // a CIGAR longer than the read is malformed; the reference reads past the string
                    sym = strand ? hs_base2(rb, tp) : 3 - hs_base2(rb, (int64_t)rlen - 1 - tp);
                }
            }
            const unsigned pm = __ballot_sync(0xffffffffu, push);
            const unsigned below = pm & ((1u << lane) - 1u);
            const int l1 = below ? 31 - __clz(below) : -1;
            const unsigned below2 = (l1 >= 0) ? (below & ~(1u << l1)) : 0u;
            const int l2 = below2 ? 31 - __clz(below2) : -1;
            const int s1 = __shfl_sync(0xffffffffu, sym, l1 < 0 ? 0 : l1);
            const int s2 = __shfl_sync(0xffffffffu, sym, l2 < 0 ? 0 : l2);
            const int prev1 = (l1 < 0) ? carry1 : s1;
            const int prev2 = (l1 < 0) ? carry2 : ((l2 < 0) ? carry1 : s2);
            if (active) {
                if (ocq) {
                    row[qp] = (uint8_t)(HS_CODE0 + 5 * prev2 + prev1 + 25 * sym);  // :238,287
                    alen++;
                    if (oty == CG_D) dist++;
                    else if (sym != hs_base2(cb, qp)) dist++;  // :254-256
                } else if (oty == CG_I) {
                    dist++;  // :337-338
                    alen++;
                }
            }
            if (pm) {  // warp-uniform
                const int last = 31 - __clz(pm);
                const unsigned rest = pm & ~(1u << last);
                const int sl = __shfl_sync(0xffffffffu, sym, last);
                const int sr = __shfl_sync(0xffffffffu, sym, rest ? 31 - __clz(rest) : 0);
                carry2 = rest ? sr : carry1;
                carry1 = sl;
            }
        }
        q += __shfl_sync(0xffffffffu, qi, 31);
        t += __shfl_sync(0xffffffffu, ti, 31);
        __syncwarp();
    }
    long long d64 = hs_warp_sum64((long long)dist), a64 = hs_warp_sum64((long long)alen);
    if (lane == 0) {
        if (d64) atomicAdd(a.stats + 3 * c + 0, (unsigned long long)d64);
        if (a64) atomicAdd(a.stats + 3 * c + 1, (unsigned long long)a64);
        if (end > start) atomicAdd(a.stats + 3 * c + 2, (unsigned long long)(end - start));
    }
}


// ---- tile index: for every 128-column tile, the reads overlapping it in ascending order ----------
// Two levels, both an ordered ballot-compaction by one warp per unit (pass 0 counts, pass 1 fills; no
// atomics, no sort, deterministic): first the reads over every super-tile of 4 tiles, found by scanning the
// reads of the contig (a few thousand); then the reads over every tile, found by scanning the list of its
// super-tile (about depth x 1.05 entries) instead of the whole contig.
template <bool FILL>
__global__ void __launch_bounds__(256) super_index_kernel(int64_t n_super, int n_contigs,
                                                          const int64_t* __restrict__ super_base,
                                                          const int64_t* __restrict__ contig_read_off,
                                                          const int32_t* __restrict__ read_start,
                                                          const int32_t* __restrict__ read_end,
                                                          int64_t* __restrict__ super_cnt_or_off,
                                                          int32_t* __restrict__ super_reads) {
    const int lane = threadIdx.x & 31;
    const int64_t su = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (su >= n_super) return;
    int lo = 0, hi = n_contigs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (super_base[mid] <= su) lo = mid; else hi = mid - 1;
    }
    const int c = lo;
    const int q0 = (int)(su - super_base[c]) * HS_SUPER_COLS, q1 = q0 + HS_SUPER_COLS;
    const int64_t r0 = contig_read_off[c], r1 = contig_read_off[c + 1];
    int64_t out = FILL ? super_cnt_or_off[su] : 0;
#pragma unroll 4
    for (int64_t rb = r0; rb < r1; rb += 32) {
        const int64_t r = rb + lane;
        bool hit = false;
        if (r < r1) {
            const int s = __ldg(read_start + r), e = __ldg(read_end + r);
            hit = (e > s) && (s < q1) && (e > q0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (FILL && hit) super_reads[out + __popc(m & ((1u << lane) - 1u))] = (int32_t)r;
        out += __popc(m);
    }
    if (!FILL && lane == 0) super_cnt_or_off[su] = out;
}

template <bool FILL>
__global__ void __launch_bounds__(256) tile_index_kernel(int64_t n_tiles, const int32_t* __restrict__ tile_contig,
                                                         const int64_t* __restrict__ tile_base,
                                                         const int64_t* __restrict__ super_base,
                                                         const int64_t* __restrict__ super_off,
                                                         const int32_t* __restrict__ super_reads,
                                                         const int32_t* __restrict__ read_start,
                                                         const int32_t* __restrict__ read_end,
                                                         int64_t* __restrict__ tile_cnt_or_off,
                                                         int32_t* __restrict__ tile_reads,
                                                         unsigned long long* __restrict__ max_count) {
    __shared__ unsigned long long s_max;  // deepest tile of the CTA: one global atomicMax per CTA
    if (!FILL) {
        if (threadIdx.x == 0) s_max = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const bool valid = tile < n_tiles;
    if (FILL && !valid) return;
    const int c = valid ? tile_contig[tile] : 0;
    const int64_t lt = valid ? tile - tile_base[c] : 0;
    const int q0 = (int)lt * HS_TILE, q1 = q0 + HS_TILE;
    const int64_t su = super_base[c] + lt / HS_SUPER_TILES;
    const int64_t l0 = valid ? super_off[su] : 0, l1 = valid ? super_off[su + 1] : 0;
    int64_t out = FILL ? tile_cnt_or_off[tile] : 0;
    for (int64_t lb = l0; lb < l1; lb += 32) {
        const int64_t l = lb + lane;
        bool hit = false;
        int32_t r = 0;
        if (l < l1) {
            r = __ldg(super_reads + l);
            const int s = __ldg(read_start + r), e = __ldg(read_end + r);
            hit = (s < q1) && (e > q0);  // e > s holds for every listed read
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (FILL && hit) tile_reads[out + __popc(m & ((1u << lane) - 1u))] = r;
        out += __popc(m);
    }
    if (!FILL) {
        if (valid && lane == 0) {
            tile_cnt_or_off[tile] = out;
            if (out > 0) atomicMax(&s_max, (unsigned long long)out);
        }
        __syncthreads();
        if (threadIdx.x == 0 && max_count && s_max > 0) atomicMax(max_count, s_max);
    }
}

// the contig every read / every tile belongs to: last c with off[c] <= i (empty contigs repeat an offset)
__device__ __forceinline__ int hs_owner(const int64_t* __restrict__ off, int n, int64_t i) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__global__ void __launch_bounds__(256) owner_fill_kernel(int nc, int64_t n_reads, const int64_t* __restrict__ contig_read_off,
                                                         int32_t* __restrict__ read_contig, int64_t n_tiles,
                                                         const int64_t* __restrict__ tile_base, int32_t* __restrict__ tile_contig) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_reads) read_contig[i] = hs_owner(contig_read_off, nc, i);
    else if (i - n_reads < n_tiles) tile_contig[i - n_reads] = hs_owner(tile_base, nc, i - n_reads);
}

// ---- host side --------------------------------------------------------------------------------
// u32 (BAM) or u16 ops -> the byte form the kernels read. N and P ops are dropped (generate_msa moves neither cursor
// for them), ops longer than 63 are split, =/X become M and H becomes S. Reads in parallel: a counting pass, the
// offsets, a filling pass.
template <typename OP>
static void hs_cigar_to_bytes(const OP* ops, const int64_t* off, int64_t n_reads, std::vector<int64_t>& out_off,
                              std::vector<uint8_t>& out) {
    static const int8_t kind_of[16] = {0, 1, 2, -1, 3, 3, -1, 0, 0, -1, -1, -1, -1, -1, -1, -1};
    out_off.assign((size_t)n_reads + 1, 0);
#pragma omp parallel for schedule(static, 256)
    for (int64_t r = 0; r < n_reads; r++) {
        int64_t n = 0;
        for (int64_t k = off[r]; k < off[r + 1]; k++) {
            const uint32_t op = (uint32_t)ops[k];
            if (kind_of[op & 15u] < 0) continue;
            const uint32_t len = op >> 4;
            n += len <= 63u ? 1 : (len + 62u) / 63u;
        }
        out_off[r + 1] = n;
    }
    for (int64_t r = 0; r < n_reads; r++) out_off[r + 1] += out_off[r];
    out.resize((size_t)out_off[n_reads] + 16);
#pragma omp parallel for schedule(static, 256)
    for (int64_t r = 0; r < n_reads; r++) {
        uint8_t* w = out.data() + out_off[r];
        for (int64_t k = off[r]; k < off[r + 1]; k++) {
            const uint32_t op = (uint32_t)ops[k];
            const int kind = kind_of[op & 15u];
            if (kind < 0) continue;
            uint32_t len = op >> 4;
            do {
                const uint32_t piece = len > 63u ? 63u : len;
                *w++ = (uint8_t)((piece << 2) | (uint32_t)kind);
                len -= piece;
            } while (len > 0);
        }
    }
}

// hsgpu_pileup_build leaves the deepest tile's read count in flight (pinned scratch of the context + event)
int hs_resolve_max_tile_reads(hsgpu_pileup* p) {
    hsgpu_ctx* ctx = p->ctx;
    if (p->max_tile_reads >= 0) return HSGPU_OK;
    if (ctx->scratch_owner != p) HS_FAIL(ctx, HSGPU_ERR_STATE, "libhsgpu: tile depth of this pileup was lost");
    HS_CUDA(ctx, cudaEventSynchronize(ctx->scratch_event));
    p->max_tile_reads = ctx->h_scratch[0];
    ctx->scratch_owner = nullptr;
    return HSGPU_OK;
}

extern "C" {

void hsgpu_pileup_destroy(hsgpu_pileup* p) {
    if (!p) return;
    hsgpu_ctx* ctx = p->ctx;
    cudaSetDevice(ctx->device);
    hs_free(ctx, p->d_create_block);
    hs_free(ctx, p->d_rank_block);
    hs_free(ctx, p->d_codes);
    hs_free(ctx, p->d_tile_reads);
    if (ctx->scratch_owner == p) ctx->scratch_owner = nullptr;
    hs_free(ctx, p->d_col_off);
    hs_free(ctx, p->d_filter_block);
    hs_free(ctx, p->d_filter_work);
    delete p;
}

static int pileup_create_impl(hsgpu_ctx* ctx, const hsgpu_pileup_input* in, hsgpu_pileup* p);

int hsgpu_pileup_create(hsgpu_ctx* ctx, const hsgpu_pileup_input* in, hsgpu_pileup** out) {
    if (!ctx || !in || !out) return HSGPU_ERR_ARG;
    *out = nullptr;
    if (in->n_contigs <= 0 || in->n_reads < 0) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: empty batch");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (in->contig_read_off[0] != 0 || in->contig_read_off[in->n_contigs] != in->n_reads)
        HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: contig_read_off must span [0, n_reads]");
    hsgpu_pileup* p = new hsgpu_pileup();
    p->ctx = ctx;
    const int rc = pileup_create_impl(ctx, in, p);
    if (rc != HSGPU_OK) {  // whatever was allocated so far goes back to the pool
        hsgpu_pileup_destroy(p);
        return rc;
    }
    *out = p;
    return HSGPU_OK;
}

static int pileup_create_impl(hsgpu_ctx* ctx, const hsgpu_pileup_input* in, hsgpu_pileup* p) {
    const int32_t nc = in->n_contigs;
    const int64_t nr = in->n_reads;
    p->n_contigs = nc;
    p->n_reads = nr;
    p->h_contig_len.assign(in->contig_len, in->contig_len + nc);
    p->h_contig_read_off.assign(in->contig_read_off, in->contig_read_off + nc + 1);
    p->h_col_base.resize(nc + 1);
    p->h_tile_base.resize(nc + 1);
    p->h_suspect_base.resize(nc + 1);
    p->h_stats.assign((size_t)3 * nc, 0);
    int64_t cols = 0, tiles = 0, sus = 0, supers = 0;
    std::vector<int64_t> super_base((size_t)nc + 1);
    for (int c = 0; c < nc; c++) {
        if (in->contig_len[c] < 0 || in->contig_read_off[c + 1] < in->contig_read_off[c])
            HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: negative contig length or read range");
        p->h_col_base[c] = cols;
        p->h_tile_base[c] = tiles;
        p->h_suspect_base[c] = sus;
        super_base[c] = supers;
        cols += in->contig_len[c];
        tiles += (in->contig_len[c] + HS_TILE - 1) / HS_TILE;
        supers += (in->contig_len[c] + HS_SUPER_COLS - 1) / HS_SUPER_COLS;
        sus += in->contig_len[c] / 6 + 2;  // suspects are > 5 columns apart (:529)
    }
    p->h_col_base[nc] = cols;
    p->h_tile_base[nc] = tiles;
    p->h_suspect_base[nc] = sus;
    super_base[nc] = supers;
    p->n_super = supers;
    p->n_cols = cols;
    p->n_tiles = tiles;
    // the CIGARs as bytes: the caller's cigar8 as it is, the wider forms converted here
    std::vector<int64_t> cig_off_conv;
    std::vector<uint8_t> cig_conv;
    const uint8_t* h_cigar = in->cigar8;
    const int64_t* h_cigar_off = in->cigar_off;
    if (!h_cigar) {
        if (in->cigar16) hs_cigar_to_bytes(in->cigar16, in->cigar_off, nr, cig_off_conv, cig_conv);
        else if (in->cigar) hs_cigar_to_bytes(in->cigar, in->cigar_off, nr, cig_off_conv, cig_conv);
        else if (in->cigar_off[nr] > 0) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: no CIGAR array given");
        else cig_off_conv.assign((size_t)nr + 1, 0), cig_conv.assign(16, 0);
        h_cigar = cig_conv.data();
        h_cigar_off = cig_off_conv.data();
    }
    p->n_cigar = h_cigar_off[nr];
    for (int64_t r = 0; r < nr; r++) {
        if (in->read_start[r] < 0 || in->read_len[r] < 0)
            HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: negative read start or length");
    }
    const int64_t contig_words = in->contig_word_off[nc];
    const int64_t read_words = in->read_word_off[nr];
    {
        HsCarve cv;
#define A(ptr, n) cv.add(&p->ptr, (n))
        A(d_contig_len, nc); A(d_contig_bases, contig_words); A(d_contig_word_off, nc + 1);
        A(d_contig_read_off, nc + 1); A(d_col_base, 4 * (nc + 1)); A(d_tile_contig, tiles);
        A(d_read_contig, nr); A(d_read_bases, read_words); A(d_read_word_off, nr + 1); A(d_read_len, nr);
        A(d_cigar, p->n_cigar + 16); A(d_cigar_off, nr + 1); A(d_read_start, nr); A(d_read_strand, nr);
        A(d_read_end, nr); A(d_read_tlead, nr); A(d_read_flags, nr); A(d_next_read, 4); A(d_row_alloc, nr + 1); A(d_row_base, nr); A(d_stats, 3 * nc);
        A(d_tile_off, tiles + 1); A(d_super_off, supers + 1);
#undef A
        HS_CUDA(ctx, cv.alloc(ctx, &p->d_create_block));
    }
    // the four per-contig offset tables travel as one array
    p->d_tile_base = p->d_col_base + (nc + 1);
    p->d_suspect_base = p->d_col_base + 2 * (nc + 1);
    p->d_super_base = p->d_col_base + 3 * (nc + 1);
    p->h_tables.resize((size_t)4 * (nc + 1));
    std::copy(p->h_col_base.begin(), p->h_col_base.end(), p->h_tables.begin());
    std::copy(p->h_tile_base.begin(), p->h_tile_base.end(), p->h_tables.begin() + (nc + 1));
    std::copy(p->h_suspect_base.begin(), p->h_suspect_base.end(), p->h_tables.begin() + 2 * (nc + 1));
    std::copy(super_base.begin(), super_base.end(), p->h_tables.begin() + 3 * (nc + 1));
    // One batch upload at a time per device: the e2e path is bound by the host link, and two contexts copying
    // at once only finish both later. Taking turns lets the kernels of the context that has its data overlap
    // the upload of the next one (callers drive one context per host thread, include/hsgpu.h).
    std::lock_guard<std::mutex> upload_turn(g_upload_mutex[ctx->device & 63]);
    HS_CUDA(ctx, hs_h2d(ctx, p->d_contig_len, in->contig_len, nc));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_contig_bases, in->contig_bases, contig_words));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_contig_word_off, in->contig_word_off, nc + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_contig_read_off, in->contig_read_off, nc + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_col_base, p->h_tables.data(), 4 * (nc + 1)));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_bases, in->read_bases, read_words));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_word_off, in->read_word_off, nr + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_len, in->read_len, nr));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_cigar_off, h_cigar_off, nr + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_start, in->read_start, nr));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_strand, in->read_strand, nr));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_cigar, h_cigar, p->n_cigar));
    // contig of every read and of every tile: looked up on the device (two sorted offset tables)
    HS_KERNEL(ctx, "owner_fill_kernel", owner_fill_kernel<<<(unsigned)((nr + tiles + 255) / 256 + 1), 256, 0, ctx->stream>>>(
        nc, nr, p->d_contig_read_off, p->d_read_contig, tiles, p->d_tile_base, p->d_tile_contig));
    // The upload turn ends when the stream is idle: the copies have landed (the caller's buffers are free again) and
    // the two small kernels behind them are done. Releasing the turn earlier, on an event behind the last copy, was
    // measured slower and unstable (2.5-4.6 ms per e2e step against a steady 2.45 ms): the next context's upload
    // then competes with this context's first kernels and its allocation-size round trip.
    HS_CUDA(ctx, hs_stream_sync(ctx));
    return HSGPU_OK;
}

int hsgpu_pileup_build(hsgpu_pileup* p) {
    if (!p) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nr = p->n_reads;
    const unsigned rblocks = (unsigned)((nr + 7) / 8);
    static_assert(PW_WARPS == 8, "span_kernel and pileup_kernel share the 8-reads-per-CTA grid");
    // rebuilt from scratch on every call (bench steps call this repeatedly on resident inputs)
    hs_free(ctx, p->d_codes);
    hs_free(ctx, p->d_tile_reads);
    p->built = p->ranked = p->have_col_off = false;
    HS_CUDA(ctx, cudaMemsetAsync(p->d_stats, 0, sizeof(unsigned long long) * 3 * p->n_contigs, ctx->stream));
    // d_totals: codes bytes, tile index entries, irregular reads, most reads over one tile, super-tile index entries
    int64_t* d_totals = nullptr;
    HsTemps temps(ctx);
    temps.own(d_totals);
    HS_CUDA(ctx, hs_alloc(ctx, &d_totals, 5));
    HS_CUDA(ctx, cudaMemsetAsync(d_totals, 0, 5 * sizeof(int64_t), ctx->stream));
    HS_CUDA(ctx, cudaMemsetAsync(p->d_next_read, 0, 4 * sizeof(unsigned int), ctx->stream));
    if (nr > 0) {
        HS_KERNEL(ctx, "span_kernel", span_kernel<<<rblocks, 256, 0, ctx->stream>>>(
            nr, p->d_cigar, p->d_cigar_off, p->d_read_start, p->d_read_contig, p->d_contig_len, p->d_read_end,
            p->d_row_alloc, p->d_read_tlead, p->d_read_flags, reinterpret_cast<unsigned long long*>(d_totals)));
    }
    int rc = hs_exclusive_scan_i64(ctx, p->d_row_alloc, p->d_row_alloc, nr, d_totals);
    if (rc) return rc;
    int64_t totals[5] = {0, 0, 0, 0, 0};
    HS_CUDA(ctx, hs_d2h(ctx, totals, d_totals, 5));
    HS_CUDA(ctx, hs_stream_sync(ctx));  // the only host round trip of the build: the allocation sizes
    p->codes_bytes = totals[0];
    p->tile_entries = totals[1];
    p->n_irregular = totals[2];
    HS_CUDA(ctx, hs_alloc(ctx, &p->d_codes, p->codes_bytes + HS_ALIGN));
    HS_CUDA(ctx, hs_alloc(ctx, &p->d_tile_reads, p->tile_entries));
    if (p->n_tiles > 0) {
        int32_t* d_super_reads = nullptr;
        HsTemps index_temps(ctx);
        index_temps.own(d_super_reads);
        HS_CUDA(ctx, hs_alloc(ctx, &d_super_reads, totals[4]));
        const unsigned sblocks = (unsigned)((p->n_super + 7) / 8), tblocks = (unsigned)((p->n_tiles + 7) / 8);
        HS_KERNEL(ctx, "super_index_kernel<false>", super_index_kernel<false><<<sblocks, 256, 0, ctx->stream>>>(
            p->n_super, p->n_contigs, p->d_super_base, p->d_contig_read_off, p->d_read_start, p->d_read_end,
            p->d_super_off, nullptr));
        rc = hs_exclusive_scan_i64(ctx, p->d_super_off, p->d_super_off, p->n_super, p->d_super_off + p->n_super);
        if (rc) return rc;
        HS_KERNEL(ctx, "super_index_kernel<true>", super_index_kernel<true><<<sblocks, 256, 0, ctx->stream>>>(
            p->n_super, p->n_contigs, p->d_super_base, p->d_contig_read_off, p->d_read_start, p->d_read_end,
            p->d_super_off, d_super_reads));
        HS_KERNEL(ctx, "tile_index_kernel<false>", tile_index_kernel<false><<<tblocks, 256, 0, ctx->stream>>>(
            p->n_tiles, p->d_tile_contig, p->d_tile_base, p->d_super_base, p->d_super_off, d_super_reads,
            p->d_read_start, p->d_read_end, p->d_tile_off, nullptr, reinterpret_cast<unsigned long long*>(d_totals + 3)));
        rc = hs_exclusive_scan_i64(ctx, p->d_tile_off, p->d_tile_off, p->n_tiles, p->d_tile_off + p->n_tiles);
        if (rc) return rc;
        HS_KERNEL(ctx, "tile_index_kernel<true>", tile_index_kernel<true><<<tblocks, 256, 0, ctx->stream>>>(
            p->n_tiles, p->d_tile_contig, p->d_tile_base, p->d_super_base, p->d_super_off, d_super_reads,
            p->d_read_start, p->d_read_end, p->d_tile_off, p->d_tile_reads, nullptr));
        hs_free(ctx, d_super_reads);
        // the deepest tile picks the histogram width of hsgpu_column_rank: read back without stalling the stream
        if (ctx->scratch_owner && ctx->scratch_owner != p) {
            rc = hs_resolve_max_tile_reads(ctx->scratch_owner);
            if (rc) return rc;
        }
        HS_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_totals + 3, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        HS_CUDA(ctx, cudaEventRecord(ctx->scratch_event, ctx->stream));
        ctx->scratch_owner = p;
        p->max_tile_reads = -1;
    } else {
        p->max_tile_reads = 0;
    }
    hs_free(ctx, d_totals);
    if (nr > 0) {
        PileupArgs a;
        a.n_reads = nr;
        a.contig_len = p->d_contig_len;
        a.contig_bases = p->d_contig_bases;
        a.contig_word_off = p->d_contig_word_off;
        a.read_contig = p->d_read_contig;
        a.read_bases = p->d_read_bases;
        a.read_word_off = p->d_read_word_off;
        a.read_len = p->d_read_len;
        a.cigar = p->d_cigar;
        a.cigar_off = p->d_cigar_off;
        a.read_start = p->d_read_start;
        a.read_strand = p->d_read_strand;
        a.read_end = p->d_read_end;
        a.read_tlead = p->d_read_tlead;
        a.read_flags = p->d_read_flags;
        a.row_off = p->d_row_alloc;
        a.row_base = p->d_row_base;
        a.codes = p->d_codes;
        a.stats = p->d_stats;
        a.next_read = p->d_next_read;
        {
            const int64_t mean_ops = p->n_cigar / std::max<int64_t>(nr, 1);
            a.ops_long = (int)std::min<int64_t>(2 * mean_ops + 1, 0x7fffffff);
            a.ops_mid = (int)std::min<int64_t>(mean_ops + 1, 0x7fffffff);
            a.k_one = 1;
            a.k_four = 4;
        }
        // persistent warps pull reads from a counter: read lengths vary by an order of magnitude
        static std::atomic<int> ctas_cache{0};  // a property of the kernel, the same on every sm_100 device
        int ctas_per_sm = ctas_cache.load();
        if (!ctas_per_sm) {
            HS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, pileup_kernel, 32 * PW_WARPS, 0));
            if (ctas_per_sm < 1) ctas_per_sm = 1;
            ctas_cache.store(ctas_per_sm);
        }
        const unsigned pgrid = (unsigned)std::min<int64_t>(rblocks, (int64_t)ctx->sm_count * ctas_per_sm);
        HS_KERNEL(ctx, "pileup_kernel", pileup_kernel<<<pgrid, 32 * PW_WARPS, 0, ctx->stream>>>(a));
        if (p->n_irregular > 0)
            HS_KERNEL(ctx, "pileup_generic_kernel", pileup_generic_kernel<<<rblocks, 256, 0, ctx->stream>>>(a));
    }
    p->built = true;
    return HSGPU_OK;
}

int hsgpu_pileup_stats(hsgpu_pileup* p, int64_t* n_cells, int64_t* distance_sum, int64_t* aligned_sum) {
    if (!p) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->built) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_pileup_stats: call hsgpu_pileup_build first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "");
    HS_CUDA(ctx, hs_d2h(ctx, (unsigned long long*)p->h_stats.data(), p->d_stats, 3 * p->n_contigs));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    for (int c = 0; c < p->n_contigs; c++) {
        if (distance_sum) distance_sum[c] = p->h_stats[3 * c + 0];
        if (aligned_sum) aligned_sum[c] = p->h_stats[3 * c + 1];
        if (n_cells) n_cells[c] = p->h_stats[3 * c + 2];
    }
    return HSGPU_OK;
}

float hsgpu_mean_distance(int64_t distance_sum, int64_t aligned_sum) {
    // `float totalDistance` grows by 1.0f per event and therefore saturates at 2^24; the length is a
    // double that starts at 1 (src/call_variants.cpp:67-68); the quotient is returned as float (:434)
    float total_distance = (float)(distance_sum > 16777216 ? 16777216 : distance_sum);
    double total_length = 1.0 + (double)aligned_sum;
    return (float)(total_distance / total_length);
}

int hsgpu_pileup_read_ends(hsgpu_pileup* p, int32_t* read_end) {
    if (!p || !read_end) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->built) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_pileup_read_ends: call hsgpu_pileup_build first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    HS_CUDA(ctx, hs_d2h(ctx, read_end, p->d_read_end, p->n_reads));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    return HSGPU_OK;
}

}  // extern "C"
