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
