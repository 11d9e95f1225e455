"""Seeded synthetic inputs for the hot path.

There is no minimap2 in the build image or on the GPU box, so the generator emits the alignment
itself: genome -> strains -> reads with mismatch/insertion/deletion errors -> exact CIGAR, in the
form the reference consumes after ``parse_SAM`` (/root/reference/src/input_output.cpp:274-536):
one record per read with contig start (POS-1), strand, the read in its ORIGINAL orientation and the
SAM CIGAR (which applies to the reverse complement when the 0x10 flag is set).

Encodings (shared with include/hsgpu.h):
  bases : u8, A=0 C=1 G=2 T=3 -- the 2-bit code of the reference's ``Sequence``
          (/root/reference/src/sequence.cpp:13-23); reverse complement = reversed, 3 - b.
  CIGAR : u32 per op, ``len << 4 | op`` with op = index in "MIDNSHP=X" (the BAM encoding).
"""
from __future__ import annotations

import dataclasses

import numpy as np

CIGAR_OPS = "MIDNSHP=X"
OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X = range(9)
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclasses.dataclass
class ContigBatch:
    """Reads aligned on one contig chunk, as flat arrays."""

    contig: np.ndarray        # u8 [L] base codes
    read_bases: np.ndarray    # u8 [sum read_len] base codes, original orientation, concatenated
    read_off: np.ndarray      # i64 [R+1]
    cigar: np.ndarray         # u32 [sum n_ops]
    cigar_off: np.ndarray     # i64 [R+1]
    start: np.ndarray         # i32 [R] contig column of the first aligned base (POS-1)
    strand: np.ndarray        # u8  [R] 1 = forward
    strain: np.ndarray        # i32 [R] strain of origin (ground truth, not an input of the path)
    name: str = "ctg"

    @property
    def n_reads(self) -> int:
        return int(self.start.shape[0])

    @property
    def length(self) -> int:
        return int(self.contig.shape[0])

    def read_len(self) -> np.ndarray:
        return np.diff(self.read_off).astype(np.int32)

    def read_str(self, i: int) -> str:
        return BASES[self.read_bases[self.read_off[i]:self.read_off[i + 1]]].tobytes().decode()

    def cigar_str(self, i: int) -> str:
        ops = self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]]
        return "".join(f"{int(o) >> 4}{CIGAR_OPS[int(o) & 15]}" for o in ops) or "*"

    def contig_str(self) -> str:
        return BASES[self.contig].tobytes().decode()


def make_strains(rng: np.random.Generator, length: int, n_strains: int, divergence, indel_frac: float = 0.0):
    """Strain 0 is the contig itself; strain k differs from it at rate divergence[k]: substitutions,
    except a fraction `indel_frac` of the variant sites which are deletions (code 4 = base absent in
    that strain, its reads carry a D there) -- deletion-majority columns exercise the reference's
    signed-char quirk for codes >= 128 (call_variants.cpp:838)."""
    base = rng.integers(0, 4, size=length, dtype=np.uint8)
    div = np.broadcast_to(np.asarray(divergence, dtype=np.float64), (n_strains,))
    strains = [base]
    for k in range(1, n_strains):
        s = base.copy()
        m = rng.random(length) < div[k]
        s[m] = (s[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
        if indel_frac > 0:
            d = m & (rng.random(length) < indel_frac)
            s[d] = 4
        strains.append(s)
    return np.stack(strains)


def simulate_contig(
    rng: np.random.Generator,
    strains: np.ndarray,
    depth: float,
    mean_len: float,
    error: float,
    abundances=None,
    sigma: float = 0.35,
    clip_prob: float = 0.1,
    hard_clip_prob: float = 0.0,
    use_eqx: bool = False,
    name: str = "ctg",
    full_length: bool = False,
) -> ContigBatch:
    """Reads sampled from `strains` (strain 0 = contig), aligned on strain 0 with their true CIGAR."""
    n_strains, L = strains.shape
    if abundances is None:
        abundances = np.full(n_strains, 1.0 / n_strains)
    abundances = np.asarray(abundances, dtype=np.float64)
    abundances = abundances / abundances.sum()
    n_reads = max(1, int(round(depth * L / mean_len)))
    span = np.clip(rng.lognormal(np.log(mean_len) - sigma * sigma / 2, sigma, n_reads), 50, None).astype(np.int64)
    span = np.minimum(span, L)
    # starts uniform over [-span/2, L - span/2) then clipped so that the ends of the contig are covered too
    start = rng.integers(-(span // 2), L - span // 2, dtype=np.int64)
    end = np.minimum(start + span, L)
    start = np.maximum(start, 0)
    if full_length:  # amplicon reads: every read covers the whole contig
        start = np.zeros(n_reads, dtype=np.int64)
        end = np.full(n_reads, L, dtype=np.int64)
    span = end - start
    keep = span >= 30
    start, span = start[keep], span[keep]
    n_reads = int(start.shape[0])
    strain = rng.choice(n_strains, size=n_reads, p=abundances).astype(np.int32)
    strand = (rng.random(n_reads) < 0.5).astype(np.uint8)

    # one entry per (read, contig column)
    N = int(span.sum())
    rid = np.repeat(np.arange(n_reads, dtype=np.int64), span)
    first = np.cumsum(span) - span
    pos = np.arange(N, dtype=np.int64) - np.repeat(first, span) + np.repeat(start, span)
    base = strains[strain[rid], pos]
    u = rng.random(N)
    is_first = np.zeros(N, dtype=bool)
    is_first[first] = True
    is_last = np.zeros(N, dtype=bool)
    is_last[first + span - 1] = True
    edge = is_first | is_last
    gap = base == 4
    base = np.where(gap, strains[0][pos], base)
    mism = (u < error / 3) & ~edge
    dele = ((u >= error / 3) & (u < 2 * error / 3) | gap) & ~edge
    ins = (u >= 2 * error / 3) & (u < error) & ~is_last
    obs = base.copy()
    obs[mism] = (obs[mism] + rng.integers(1, 4, size=int(mism.sum()), dtype=np.uint8)) & 3

    # expanded alignment: per column one of M/D, optionally followed by one I
    n_emit = (~dele).astype(np.int64) + ins.astype(np.int64)      # read bases produced per column
    n_ops = 1 + ins.astype(np.int64)                               # expanded ops per column
    E = int(n_ops.sum())
    op_first = np.cumsum(n_ops) - n_ops
    ops = np.full(E, OP_M, dtype=np.uint8)
    if use_eqx:
        ops[op_first] = np.where(obs == strains[0][pos], OP_EQ, OP_X)
    ops[op_first[dele]] = OP_D
    ops[op_first[ins] + 1] = OP_I
    op_rid = np.repeat(rid, n_ops)

    # aligned-orientation read bases
    T = int(n_emit.sum())
    b_first = np.cumsum(n_emit) - n_emit
    aligned = np.empty(T, dtype=np.uint8)
    aligned[b_first[~dele]] = obs[~dele]
    ins_slot = b_first[ins] + (~dele[ins]).astype(np.int64)
    aligned[ins_slot] = rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)
    aligned_len = np.bincount(rid, weights=n_emit, minlength=n_reads).astype(np.int64)

    # clips: soft (bases present, unaligned) and optionally hard (bases present in the FASTA too,
    # because the reference always loads the full read from the reads file)
    clip5 = np.where(rng.random(n_reads) < clip_prob, rng.integers(1, 200, n_reads), 0).astype(np.int64)
    clip3 = np.where(rng.random(n_reads) < clip_prob, rng.integers(1, 200, n_reads), 0).astype(np.int64)
    hard = rng.random(n_reads) < hard_clip_prob
    read_len = aligned_len + clip5 + clip3
    read_off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(read_len, out=read_off[1:])
    oriented = rng.integers(0, 4, size=int(read_off[-1]), dtype=np.uint8)
    a_first = np.cumsum(aligned_len) - aligned_len
    dst = np.arange(T, dtype=np.int64) - np.repeat(a_first, aligned_len) + np.repeat(read_off[:-1] + clip5, aligned_len)
    oriented[dst] = aligned

    # original orientation = reverse complement of the aligned orientation for reverse-strand reads
    read_bases = oriented.copy()
    rev_reads = np.nonzero(strand == 0)[0]
    if rev_reads.size:
        rl = read_len[rev_reads]
        idx_first = np.cumsum(rl) - rl
        k = np.arange(int(rl.sum()), dtype=np.int64) - np.repeat(idx_first, rl)
        src = np.repeat(read_off[rev_reads], rl) + k
        dstr = np.repeat(read_off[rev_reads] + rl - 1, rl) - k
        read_bases[dstr] = 3 - oriented[src]

    # run-length encode the expanded ops per read
    brk = np.ones(E, dtype=bool)
    brk[1:] = (ops[1:] != ops[:-1]) | (op_rid[1:] != op_rid[:-1])
    run_start = np.nonzero(brk)[0]
    run_len = np.diff(np.append(run_start, E))
    run_op = ops[run_start].astype(np.uint32)
    run_rid = op_rid[run_start]
    runs_per_read = np.bincount(run_rid, minlength=n_reads).astype(np.int64)
    has5, has3 = clip5 > 0, clip3 > 0
    cig_n = runs_per_read + has5 + has3
    cigar_off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(cig_n, out=cigar_off[1:])
    cigar = np.empty(int(cigar_off[-1]), dtype=np.uint32)
    run_first = np.cumsum(runs_per_read) - runs_per_read
    dst_run = np.arange(run_start.shape[0], dtype=np.int64) - np.repeat(run_first, runs_per_read) + np.repeat(
        cigar_off[:-1] + has5, runs_per_read
    )
    cigar[dst_run] = (run_len.astype(np.uint32) << 4) | run_op
    clip_op = np.where(hard, OP_H, OP_S).astype(np.uint32)
    cigar[cigar_off[:-1][has5]] = (clip5[has5].astype(np.uint32) << 4) | clip_op[has5]
    cigar[(cigar_off[1:] - 1)[has3]] = (clip3[has3].astype(np.uint32) << 4) | clip_op[has3]

    return ContigBatch(
        contig=np.where(strains[0] == 4, 0, strains[0]).astype(np.uint8),
        read_bases=read_bases,
        read_off=read_off,
        cigar=cigar,
        cigar_off=cigar_off,
        start=start.astype(np.int32),
        strand=strand,
        strain=strain,
        name=name,
    )


def empty_contig(contig: np.ndarray, name: str) -> ContigBatch:
    """a contig no read aligns on"""
    return ContigBatch(contig=contig.astype(np.uint8), read_bases=np.zeros(0, np.uint8), read_off=np.zeros(1, np.int64),
                       cigar=np.zeros(0, np.uint32), cigar_off=np.zeros(1, np.int64), start=np.zeros(0, np.int32),
                       strand=np.zeros(0, np.uint8), strain=np.zeros(0, np.int32), name=name)


def _unpack2(packed: np.ndarray, n: int) -> np.ndarray:
    b = packed.astype(np.uint8)
    return np.stack([b & 3, (b >> 2) & 3, (b >> 4) & 3, (b >> 6) & 3], axis=1).reshape(-1)[:n]


def simple_mock_config(rng: np.random.Generator, n_chunks=None):
    """BASELINE configs[0]: the contigs of the reference's test/simple_mock/assembly.gfa (4 segments, 2 links,
    one reverse) with reads simulated from the 3 haplotypes of mock_reference.fasta at equal abundance (8 kb
    mean, 8 % error, 30x each) -- mock_reads.fasta is absent from the reference repository.
    The sequences come from tests/golden/simple_mock.npz (made by tests/golden/make_simple_mock.py)."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "simple_mock.npz")
    g = np.load(path)
    haps = [_unpack2(g[f"hap{i}"], int(g["hap_length"][i])) for i in range(3)]
    chunks = []
    for i, name in enumerate(g["names"]):
        L, at = int(g["lengths"][i]), int(g["offsets"][i])
        contig = _unpack2(g[f"seg{i}"], L)
        if at < 0:  # consensus_2: 500 bp that belong to none of the haplotypes -> no read aligns on it
            chunks.append(empty_contig(contig, str(name)))
            continue
        strains = np.stack([contig] + [h[at:at + L] for h in haps])
        chunks.append(simulate_contig(rng, strains, 90, 8000, 0.08, abundances=[0, 1, 1, 1], name=str(name)))
    if n_chunks is not None:
        chunks = chunks[:n_chunks]
    return chunks, [str(l) for l in g["links"]]


def amplicon_contig(rng: np.random.Generator, length=10_000, n_reads=2000, n_variants=8, error=0.10):
    """the 10 kb / 2000x amplicon contig of BASELINE configs[4]: full-length reads, 8 planted variants shared out
    over 3 haplotypes (the contig itself is haplotype 0)"""
    base = rng.integers(0, 4, size=length, dtype=np.uint8)
    sites = np.sort(rng.choice(np.arange(200, length - 200), size=n_variants, replace=False))
    strains = np.stack([base, base.copy(), base.copy()])
    for k, p in enumerate(sites):
        alt = (base[p] + 1 + (k % 3)) & 3
        if k % 3 != 0:
            strains[1, p] = alt   # haplotype 1 carries variants 1,2,4,5,7
        if k % 3 != 1:
            strains[2, p] = alt   # haplotype 2 carries variants 0,2,3,5,6
    return simulate_contig(rng, strains, n_reads, length, error, abundances=[0.4, 0.35, 0.25], clip_prob=0.0,
                           name="amplicon", full_length=True)


CONFIG_SPEC = {
    2: dict(genome=5_000_000, strains=2, div=[0, 0.01], depth=60, mean_len=10000, err=0.10, ab=None,
            what="bacterial genome, 2 strains 1% apart, ONT-like reads 10 kb mean, 10% error, 60x"),
    3: dict(genome=60_000_000, strains=5, div=None, depth=100, mean_len=10000, err=0.10, ab="log",
            what="metagenome 20 x 3 Mb, 5 strains each at 0.3-3% divergence and 1-40% abundance, ONT 100x"),
    4: dict(genome=20_000_000, strains=4, div=[0, 0.005, 0.03, 0.03], depth=80, mean_len=15000, err=0.005, ab=None,
            what="allotetraploid (2 sub-genomes 3% apart, homologues 0.5% apart), HiFi-like reads 15 kb, 0.5% error, 4x20x"),
    5: dict(genome=64_000_000, strains=2, div=[0, 0.001], depth=40, mean_len=10000, err=0.10, ab=None,
            what="diploid chromosome 0.1% heterozygous, ONT 40x, plus a 10 kb amplicon contig at 2000x with 8 variants in 3 haplotypes"),
}


def _make_chunk(spec, rng, ci, l):
    div = spec["div"]
    ab = spec["ab"]
    if div is None:
        div = np.concatenate([[0.0], np.exp(rng.uniform(np.log(0.003), np.log(0.03), spec["strains"] - 1))])
    if ab == "log":
        ab = np.exp(rng.uniform(np.log(0.01), np.log(0.4), spec["strains"]))
    strains = make_strains(rng, l, spec["strains"], div)
    return simulate_contig(rng, strains, spec["depth"], spec["mean_len"], spec["err"], abundances=ab, name=f"ctg@{ci}")


def _make_chunk_job(job):
    config, seed, ci, l = job
    return _make_chunk(CONFIG_SPEC[config], np.random.default_rng([seed, ci]), ci, l)


def make_config(config: int, scale: float = 1.0, seed: int | None = None, n_chunks: int | None = None, workers: int = 1):
    """The five BASELINE.json configs as lists of <=300 kb contig chunks.

    `scale` shrinks the genome length (not depth, read length or error), `n_chunks` caps the number
    of chunks generated. Returns (list[ContigBatch], dict describing the realisation; info["amplicon"] is the
    <amplicon> argument the two executables take for this config, info["links"] the L lines of the GFA).
    """
    seed = config if seed is None else seed
    rng = np.random.default_rng(seed)
    CH = 300_000
    if config == 1:
        chunks, links = simple_mock_config(rng, n_chunks)
        info = dict(config=1, genome=sum(c.length for c in chunks), chunks=len(chunks), depth=90, mean_len=8000,
                    error=0.08, strains=3, seed=seed, amplicon=0, links=links,
                    description="test/simple_mock/assembly.gfa (4 segments, 2 links) + reads simulated from the 3 "
                                "haplotypes of mock_reference.fasta, 8 kb mean, 8% error, 30x each")
        return chunks, info
    spec = CONFIG_SPEC[config]
    total = int(spec["genome"] * scale)
    lengths = []
    left = total
    while left > 0:
        l = min(CH, left)
        lengths.append(l)
        left -= l
    if n_chunks is not None:
        lengths = lengths[:n_chunks]
    if config == 2:  # one generator through all chunks (the realisation every round-1 number was taken on)
        chunks = [_make_chunk(spec, rng, ci, l) for ci, l in enumerate(lengths)]
    else:            # one generator per chunk, so that chunks can be made in any order / in parallel
        jobs = [(config, seed, ci, l) for ci, l in enumerate(lengths)]
        if workers > 1 and len(jobs) > 1:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
                chunks = pool.map(_make_chunk_job, jobs, chunksize=1)
        else:
            chunks = [_make_chunk_job(j) for j in jobs]
    if config == 5 and n_chunks is None:
        chunks.append(amplicon_contig(np.random.default_rng(seed + 7919)))
    info = dict(config=config, genome=total, chunks=len(chunks), depth=spec["depth"], mean_len=spec["mean_len"],
                error=spec["err"], strains=spec["strains"], seed=seed, amplicon=1 if config == 5 else 0, links=[],
                description=f"synthetic {total / 1e6:g} Mb " + spec["what"])
    return chunks, info


def cigar_text(cigar: np.ndarray):
    """all ops of `cigar` as SAM text in one byte array, and the byte offset of every op (n_ops + 1 entries)"""
    ln = (cigar >> 4).astype(np.int64)
    op = (cigar & 15).astype(np.int64)
    nd = np.ones(ln.shape[0], dtype=np.int64)
    p = 10
    while ln.size and (ln >= p).any():
        nd += ln >= p
        p *= 10
    off = np.zeros(ln.shape[0] + 1, dtype=np.int64)
    np.cumsum(nd + 1, out=off[1:])
    buf = np.empty(int(off[-1]), dtype=np.uint8)
    letters = np.frombuffer(CIGAR_OPS.encode(), dtype=np.uint8)
    buf[off[:-1] + nd] = letters[op]
    k, p = 0, 1
    while ln.size and (nd > k).any():
        m = nd > k
        buf[off[:-1][m] + nd[m] - 1 - k] = 48 + (ln[m] // p) % 10
        k += 1
        p *= 10
    return buf, off


def chunk_text(c: ContigBatch, fastq: bool = False):
    """(GFA S line, SAM @SQ line, reads file records, SAM records) of one contig chunk, as bytes"""
    seq_all = BASES[c.read_bases].tobytes()
    cig_all, cig_byte = cigar_text(c.cigar)
    cig_all = cig_all.tobytes()
    rl = c.read_len()
    fa, sam = [], []
    cname = c.name.encode()
    for i in range(c.n_reads):
        rname = b"%s_r%d" % (cname, i)
        seq = seq_all[c.read_off[i]:c.read_off[i + 1]]
        if fastq:  # quality lines that start with '@' exercise the record detection of parse_reads
            fa.append(b"@%s some comment\n%s\n+\n%s%s\n" % (rname, seq, b"@" if i % 3 == 0 else b"I", b"I" * (len(seq) - 1)))
        else:
            fa.append(b">%s\n%s\n" % (rname, seq))
        cg = cig_all[cig_byte[c.cigar_off[i]]:cig_byte[c.cigar_off[i + 1]]] or b"*"
        sam.append(b"%s\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t*\t*\tNM:i:0\tLN:i:%d\n" %
                   (rname, 0 if c.strand[i] else 16, cname, int(c.start[i]) + 1, cg, int(rl[i])))
    return (b"S\t%s\t%s\n" % (cname, c.contig_str().encode()), b"@SQ\tSN:%s\tLN:%d\n" % (cname, c.length),
            b"".join(fa), b"".join(sam))


def write_files(chunks, prefix: str, fastq: bool = False, links=()):
    """GFA + FASTA (or FASTQ) + SAM in the form the reference executables read (LN:i / NM:i tags appended)."""
    ext = ".fastq" if fastq else ".fasta"
    with open(prefix + ".gfa", "wb") as gfa, open(prefix + ext, "wb") as fa, open(prefix + ".sam", "wb") as sam:
        texts = [chunk_text(c, fastq) for c in chunks]
        for t in texts:
            gfa.write(t[0])
            sam.write(t[1])
        for l in links:
            gfa.write(l.encode() + b"\n")
        for t in texts:
            fa.write(t[2])
            sam.write(t[3])
    return prefix + ".gfa", prefix + ext, prefix + ".sam"
