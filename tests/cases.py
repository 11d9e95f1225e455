"""Seeded synthetic cases shared by the CPU and GPU parity tests."""
import numpy as np

from hairsplitter_b200 import synth


def small_case(seed=11, length=6000, depth=30, mean_len=1500, error=0.08, n_strains=3, indel_frac=0.25,
               hard=0.2, eqx=False):
    rng = np.random.default_rng(seed)
    st = synth.make_strains(rng, length, n_strains, [0, 0.012, 0.02][:n_strains], indel_frac=indel_frac)
    return synth.simulate_contig(rng, st, depth, mean_len, error, hard_clip_prob=hard, use_eqx=eqx, name=f"c{seed}")


def medium_case(seed=21):
    rng = np.random.default_rng(seed)
    st = synth.make_strains(rng, 40000, 3, [0, 0.01, 0.02], indel_frac=0.2)
    return synth.simulate_contig(rng, st, 40, 4000, 0.08, hard_clip_prob=0.3, name=f"m{seed}")


def hifi_case(seed=31):
    rng = np.random.default_rng(seed)
    st = synth.make_strains(rng, 30000, 4, [0, 0.005, 0.03, 0.03], indel_frac=0.1)
    return synth.simulate_contig(rng, st, 60, 6000, 0.005, name=f"h{seed}")


def deep_case(seed=41):
    """amplicon-like: short contig, very deep (exercises multi-batch row staging)"""
    rng = np.random.default_rng(seed)
    st = synth.make_strains(rng, 1500, 3, [0, 0.01, 0.01], indel_frac=0.3)
    return synth.simulate_contig(rng, st, 700, 1400, 0.06, sigma=0.05, name=f"d{seed}")


def ragged_cases():
    """edge cases: empty contig batch members, a contig with no reads, reads running off the contig end"""
    rng = np.random.default_rng(51)
    a = small_case(seed=52, length=700, depth=12, mean_len=300)
    # contig without reads
    b = small_case(seed=53, length=300, depth=5, mean_len=100)
    b = synth.ContigBatch(contig=b.contig, read_bases=np.zeros(0, np.uint8), read_off=np.zeros(1, np.int64),
                          cigar=np.zeros(0, np.uint32), cigar_off=np.zeros(1, np.int64), start=np.zeros(0, np.int32),
                          strand=np.zeros(0, np.uint8), strain=np.zeros(0, np.int32), name="empty")
    # reads whose CIGAR runs past the contig end / starts at the very end
    c = small_case(seed=54, length=900, depth=15, mean_len=400)
    c.start[:] = np.minimum(c.start + 300, c.length - 1).astype(np.int32)
    # a 1-column contig
    d = small_case(seed=55, length=130, depth=8, mean_len=60)
    return [a, b, c, d]


def walk_edge_case(seed=71, length=5000):
    """hand-made CIGARs for the windowed CIGAR walk of pileup_kernel: ops longer than one window, windows
    ending inside I/D ops, > 128 ops, long and mid-alignment clips, N/P ops, a CIGAR longer than its read,
    reads that run off the contig end or start at it, reads without any cell"""
    rng = np.random.default_rng(seed)
    letters = "MIDNSHP=X"
    consumes_read = set("MIS=X")
    spec = [
        (10, 1, "3000S2500M", 0), (100, 0, "10H5S1500M1600D50M1700I100M7S", 0), (0, 1, "1M1I1M1D" * 300, 0),
        (4000, 0, "100M50S100M", 0), (4900, 1, "300M", -200), (2000, 1, "20I5N30M4P2D10M", 0),
        (1000, 0, "200=1X200=", 0), (300, 1, "50S20I", 0), (length, 1, "10M", 0), (length - 1, 0, "10M", 0),
        (0, 1, "1472M", 0), (16, 0, "1473M", 0), (32, 1, "960M", 0), (33, 0, "959M", 0), (47, 1, "961M", 0),
        (5, 1, "1M", 0), (6, 0, "2M", 0), (7, 1, "3M", 0), (64, 1, "1470M10I10M", 0), (65, 0, "1470M10D10M", 0),
        (66, 1, "958M5I5M", 0), (67, 0, "958M5D5M", 0), (1, 1, "2I1D2I1D1M" * 150, 0), (3, 0, "31M1D" * 140, 0),
        (2, 1, "1471M1I1D1M", 0), (15, 0, "2943M1I1M", 0), (200, 1, "4000S40M", 0), (0, 1, "4999M", 0),
        (1, 0, "4200S4100M4300S", 0),
    ]
    import re
    starts, strands, cig, cig_off, bases, read_off = [], [], [], [0], [], [0]
    for start, strand, s, extra in spec:
        ops = [(int(n), letters.index(c)) for n, c in re.findall(r"(\d+)([MIDNSHP=X])", s)]
        rl = max(1, sum(n for n, c in ops if letters[c] in consumes_read) + extra)
        cig += [(n << 4) | c for n, c in ops]
        cig_off.append(len(cig))
        bases.append(rng.integers(0, 4, rl).astype(np.uint8))
        read_off.append(read_off[-1] + rl)
        starts.append(start)
        strands.append(strand)
    contig = rng.integers(0, 4, length).astype(np.uint8)
    n = len(spec)
    return synth.ContigBatch(contig=contig, read_bases=np.concatenate(bases), read_off=np.array(read_off, np.int64),
                             cigar=np.array(cig, np.uint32), cigar_off=np.array(cig_off, np.int64),
                             start=np.array(starts, np.int32), strand=np.array(strands, np.uint8),
                             strain=np.zeros(n, np.int32), name="walk_edges")


# ---- separate_reads stages: SNP columns and windows -------------------------------------------------------------
def snp_columns(oracle, cb, mean_error=0.08):
    """the suspect columns of a synthetic contig as (n_reads, snp_off, read_idx, code, ref_base, second_base)"""
    o = oracle.pileup(cb)
    oc = oracle.call_variants(o["col_off"], o["code"], mean_error)
    pos = oc["suspect_pos"]
    snp_off = np.zeros(pos.size + 1, np.int64)
    idx, code = [np.zeros(0, np.uint32)], [np.zeros(0, np.uint8)]
    for j, q in enumerate(pos):
        a, b = o["col_off"][q], o["col_off"][q + 1]
        idx.append(o["read_idx"][a:b])
        code.append(o["code"][a:b])
        snp_off[j + 1] = snp_off[j] + (b - a)
    return (cb.n_reads, snp_off, np.concatenate(idx).astype(np.uint32), np.concatenate(code).astype(np.uint8),
            oc["ref_base"][pos].astype(np.uint8), oc["second_base"][pos].astype(np.uint8)), pos


def coarse_columns(rng, n_reads, n_snps, p_alt=0.4, p_other=0.05):
    """every read covers every SNP, few SNPs: distances take few distinct values, so groups of equal distances
    straddle the "first five neighbours" cut (the case the CUDA path replays with the reference's sort)"""
    hap = rng.integers(0, 2, n_reads)
    rb = rng.integers(33, 158, n_snps).astype(np.uint8)
    sb = ((rb - 33 + rng.integers(1, 124, n_snps)) % 125 + 33).astype(np.uint8)
    snp_off = np.arange(n_snps + 1, dtype=np.int64) * n_reads
    idx = np.tile(np.arange(n_reads, dtype=np.uint32), n_snps)
    code = np.zeros(n_snps * n_reads, np.uint8)
    for s in range(n_snps):
        u = rng.random(n_reads)
        alt = (hap == 1) ^ (u < p_alt * 0.3)
        c = np.where(alt, sb[s], rb[s])
        c = np.where(rng.random(n_reads) < p_other, np.uint8(33 + (int(rb[s]) + 7) % 125), c)
        code[s * n_reads:(s + 1) * n_reads] = c
    return (n_reads, snp_off, idx, code, rb, sb)


def windows_of(col, pos, size=2000):
    """windows like the reference's walk: reads present at the first and at the last SNP of each 2 kb window;
    returns [(masked reads ascending, indices of the window's SNPs)]"""
    n_reads, snp_off, idx = col[0], col[1], col[2]
    out = []
    if pos.size == 0:
        return out
    for w0 in range(0, int(pos.max()) + 1, size):
        inside = np.nonzero((pos >= w0) & (pos < w0 + size))[0]
        if inside.size == 0:
            continue
        first = idx[snp_off[inside[0]]:snp_off[inside[0] + 1]]
        last = idx[snp_off[inside[-1]]:snp_off[inside[-1] + 1]]
        masked = np.intersect1d(first, last).astype(np.int32)
        if masked.size:
            out.append((masked, inside))
    return out


def start_labels(col, s, masked):
    """labels a clustering run starts from at SNP s (src/separate_reads.cpp:1680-1693), as LOCAL indices"""
    n_reads, snp_off, idx, code = col[0], col[1], col[2], col[3]
    local = np.full(n_reads, -1, np.int64)
    local[masked] = np.arange(masked.size)
    lab = np.arange(masked.size, dtype=np.int32)
    first = {}
    for r, c in zip(idx[snp_off[s]:snp_off[s + 1]], code[snp_off[s]:snp_off[s + 1]]):
        if local[r] >= 0:
            first.setdefault(int(c), int(local[r]))
            lab[local[r]] = first[int(c)]
    return lab
