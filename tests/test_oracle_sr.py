"""CPU tests of the separate_reads stages: (1) the oracle's restatement of create_read_graph_matrix and
chinese_whispers_high_memory (oracle/hs_oracle_sr.cpp) against the compiled reference (oracle/_ref/libhsref_sr.so,
std::random_device pinned by oracle/ref_pin_rng.cpp); (2) the host logic of the HS_separate_reads drop-in
(hairsplitter_b200/host/hs_sepreads*.cpp: .col parser, window walk, finalize_clustering chain, low-memory path, ploidy
limit) driven by oracle/sr_hostcheck -- the product pipeline with the three GPU stages replaced by the oracle --
against oracle/_ref/HS_separate_reads_pinned: the .gro files must be byte-identical."""
import os
import subprocess

import numpy as np
import pytest

import cases
from hairsplitter_b200 import synth
from oracle.pyoracle import PIN_SEED, Oracle, RefSR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CV = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")
REF_SR = os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_pinned")
HOSTCHECK = os.path.join(ROOT, "oracle", "sr_hostcheck")

needs_ref = pytest.mark.skipif(not (RefSR.available() and os.path.exists(REF_SR)), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def _check_graph_and_whispers(oracle, col, masked, snps_for_runs, error_rate):
    sim, diff = oracle.read_pair_counts(*col)
    adj_off, adj = oracle.read_graph(sim, diff, masked, error_rate)
    r_off, r_adj = RefSR.read_graph(col, masked, error_rate)
    assert np.array_equal(adj_off, r_off) and np.array_equal(adj, r_adj)
    for s in snps_for_runs:
        init = cases.start_labels(col, s, masked)
        got = oracle.chinese_whispers(col[0], masked, adj_off, adj, init, PIN_SEED)
        want = RefSR.chinese_whispers(col[0], masked, adj_off, adj, masked[init])
        assert np.array_equal(masked[got], want)
    return adj.size


@needs_ref
@pytest.mark.parametrize("case", ["ont", "hifi"])
def test_oracle_graph_and_whispers_match_reference(oracle, case):
    cb, err = {"ont": (cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06), 0.06),
               "hifi": (cases.hifi_case(), 0.01)}[case]
    col, pos = cases.snp_columns(oracle, cb, err)
    wins = cases.windows_of(col, pos)
    assert len(wins) >= 3
    links = 0
    for masked, inside in wins[:6]:
        links += _check_graph_and_whispers(oracle, col, masked, inside[:4], err)
    assert links > 0


@needs_ref
def test_oracle_graph_with_tied_distances(oracle):
    """few SNPs -> coarse distances: the order std::sort leaves equal keys in decides the neighbours"""
    rng = np.random.default_rng(17)
    for n_reads, n_snps, err in [(40, 6, 0.1), (90, 9, 0.15), (33, 4, 0.3), (64, 12, 0.6)]:
        col = cases.coarse_columns(rng, n_reads, n_snps)
        masked = np.sort(rng.choice(n_reads, size=n_reads - 5, replace=False)).astype(np.int32)
        _check_graph_and_whispers(oracle, col, masked, range(n_snps), err)


def _gro_pair(tmp, chunks, err, low="0", rare="0", amp="0", ploidy=None, threads="1"):
    for i, c in enumerate(chunks):
        c.name = f"ctg{i}"
    gfa, reads, sam = synth.write_files(chunks, os.path.join(tmp, "in"))
    col = os.path.join(tmp, "a.col")
    subprocess.run([REF_CV, gfa, reads, sam, "4", tmp, os.path.join(tmp, "err"), amp, "0", col, os.path.join(tmp, "a.vcf"), "0.33"],
                   check=True, stdout=subprocess.DEVNULL)
    pl = os.path.join(tmp, "ploidy.txt")
    if ploidy:
        with open(pl, "w") as f:
            f.write("".join(f"{c.name}\t{ploidy}\n" for c in chunks))
    ref, ours = os.path.join(tmp, "ref.gro"), os.path.join(tmp, "ours.gro")
    subprocess.run([REF_SR, col, "1", err, pl, low, rare, amp, ref, "0"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([HOSTCHECK, col, threads, err, pl, low, rare, amp, ours, "0"], check=True, stdout=subprocess.DEVNULL)
    return open(ref, "rb").read(), open(ours, "rb").read()


@needs_ref
@pytest.mark.parametrize("case", ["ont_multi", "hifi", "low_memory", "rarest", "ploidy", "coverage_over_1000", "amplicon"])
def test_host_pipeline_gro_identical_to_pinned_reference(tmp_path, case):
    kw, threads = {}, "1"
    if case == "ont_multi":
        chunks = [cases.small_case(seed=102, length=40000, depth=70, mean_len=7000, error=0.06),
                  cases.small_case(seed=5, length=3000, depth=12, mean_len=900, hard=0.4),
                  cases.small_case(seed=6, length=90, depth=5, mean_len=60)]
        err, threads = "0.06", "4"
    elif case == "hifi":
        chunks, err = [cases.hifi_case()], "0.01"
    elif case == "low_memory":
        chunks, err, kw = [cases.small_case(seed=103, length=12000, depth=40, mean_len=4000, error=0.06)], "0.06", dict(low="1")
    elif case == "rarest":
        chunks, err, kw = [cases.small_case(seed=104, length=30000, depth=50, mean_len=5000, error=0.06)], "0.06", dict(rare="0.2")
    elif case == "ploidy":
        chunks, err, kw = [cases.small_case(seed=105, length=30000, depth=60, mean_len=5000, error=0.05)], "0.05", dict(ploidy=2)
    elif case == "coverage_over_1000":
        # low_memory_now without -l: neighbour lists for the runs, but finalize_clustering on the empty matrix (:1708)
        chunks, err = [cases.small_case(seed=107, length=1500, depth=1500, mean_len=1400, error=0.05)], "0.05"
    else:
        chunks = [cases.small_case(seed=108, length=2500, depth=300, mean_len=2400, error=0.05),
                  cases.small_case(seed=109, length=1200, depth=30, mean_len=800, error=0.05, n_strains=2)]
        err, kw = "0.05", dict(amp="1")
    ref, ours = _gro_pair(str(tmp_path), chunks, err, threads=threads, **kw)
    assert ref == ours
    assert ref.count(b"GROUP") >= 1


def _golden(tmp, name):
    """a committed .col of the reference's HS_call_variants and the pinned reference's .gro for it
    (tests/golden/make_golden_sr.py); returns (path of the unpacked .col, .gro bytes, arguments)"""
    import gzip
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_sr
    _, err, low, rare, amp = make_golden_sr.CASES[name]
    g = os.path.join(ROOT, "tests", "golden")
    col = os.path.join(tmp, name + ".col")
    with open(col, "wb") as f:
        f.write(gzip.open(os.path.join(g, f"sr_{name}.col.gz")).read())
    return col, gzip.open(os.path.join(g, f"sr_{name}.gro.gz")).read(), (err, low, rare, amp)


@pytest.mark.parametrize("case", ["ont", "lowmem", "amplicon"])
def test_host_pipeline_matches_golden_gro(tmp_path, case):
    """the same comparison against committed fixtures: needs neither /root/reference nor oracle/_ref at run time"""
    col, want, (err, low, rare, amp) = _golden(str(tmp_path), case)
    out = os.path.join(str(tmp_path), "ours.gro")
    subprocess.run([HOSTCHECK, col, "2", err, os.path.join(str(tmp_path), "no_ploidy"), low, rare, amp, out, "0"], check=True,
                   stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == want
    assert want.count(b"GROUP") >= 1


def test_usage_and_help_status(tmp_path):
    """hairsplitter.py probes the executable with --help and expects status 0 (hairsplitter.py:241-252)"""
    r = subprocess.run([HOSTCHECK, "--help"], stdout=subprocess.PIPE)
    assert r.returncode == 0 and b"Usage" in r.stdout
    r = subprocess.run([HOSTCHECK, "a", "b"], stdout=subprocess.PIPE)
    assert r.returncode == 1 and b"Usage" in r.stdout
