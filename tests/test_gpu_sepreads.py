"""GPU parity of the separate_reads stages through the C ABI: the read graph of every window (hsgpu_graph_build =
create_read_graph_matrix, reference src/separate_reads.cpp:706-828) and the chinese-whispers runs
(hsgpu_graph_whispers = chinese_whispers_high_memory, src/cluster_graph.cpp:240-310) against the oracle
(oracle/hs_oracle_sr.cpp, itself pinned against the compiled reference by tests/test_oracle_sr.py), and the
HS_separate_reads drop-in executable against the RNG-pinned reference executable: .gro byte-identical.
Integer / index outputs: tolerance 0."""
import os
import subprocess

import numpy as np
import pytest

import cases
from hairsplitter_b200 import api, synth
from oracle.pyoracle import PIN_SEED, Oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_separate_reads")
REF_CV = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")
REF_SR = os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_pinned")


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def _check_batch(ctx, oracle, cols, windows, error_rate, seeds):
    """cols: contigs; windows: [(contig, masked, snp indices to start runs from)]"""
    pairs = api.Pairs(ctx, cols)
    pairs.compute()
    g = api.Graph(pairs, [(c, m) for c, m, _ in windows], error_rate)
    g.build()
    adj_off, adj = g.adjacency()
    counts = [oracle.read_pair_counts(*col) for col in cols]
    base, want_graphs = 0, []
    for c, masked, _ in windows:
        o_off, o_adj = oracle.read_graph(counts[c][0], counts[c][1], masked, error_rate)
        m = masked.size
        got_off = adj_off[base:base + m + 1] - adj_off[base]
        assert np.array_equal(got_off, o_off), "degrees differ"
        assert np.array_equal(adj[adj_off[base]:adj_off[base + m]], o_adj), "neighbours differ"
        want_graphs.append((o_off, o_adj))
        base += m
    # clustering runs
    run_window, init, want = [], [], []
    for w, (c, masked, starts) in enumerate(windows):
        for s in starts:
            lab = cases.start_labels(cols[c], s, masked)
            run_window.append(w)
            init.append(lab)
            want.append(oracle.chinese_whispers(cols[c][0], masked, want_graphs[w][0], want_graphs[w][1], lab, seeds))
    rank = []
    for col in cols:
        for sd in seeds:
            o = oracle.shuffled_order(col[0], sd)
            r = np.zeros(col[0], np.int32)
            r[o] = np.arange(col[0], dtype=np.int32)
            rank.append(r)
    if run_window:
        got = g.whispers(run_window, np.concatenate(init), np.concatenate(rank), n_orders=len(seeds))
        assert np.array_equal(got, np.concatenate(want))
    n_links, replayed = adj.size, g.replayed
    g.close()
    pairs.close()
    return n_links, replayed, len(run_window)


@pytest.mark.parametrize("case", ["ont", "hifi"])
def test_graph_and_whispers_match_oracle(ctx, oracle, case):
    cbs, err = {"ont": ([cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06), cases.medium_case()], 0.06),
                "hifi": ([cases.hifi_case()], 0.01)}[case]
    cols, windows = [], []
    for c, cb in enumerate(cbs):
        col, pos = cases.snp_columns(oracle, cb, err)
        cols.append(col)
        windows += [(c, masked, inside[:5]) for masked, inside in cases.windows_of(col, pos)]
    assert len(windows) >= 4
    links, _, runs = _check_batch(ctx, oracle, cols, windows, err, [PIN_SEED])
    assert links > 0 and runs > 0


def test_several_sweep_orders(ctx, oracle):
    """unpinned mode: a different shuffled order per sweep (n_orders > 1)"""
    cb = cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06)
    col, pos = cases.snp_columns(oracle, cb, 0.06)
    windows = [(0, masked, inside[:6]) for masked, inside in cases.windows_of(col, pos)]
    _check_batch(ctx, oracle, [col], windows, 0.06, [11, 22, 33, 44])


def test_tied_distances_are_replayed_exactly(ctx, oracle):
    """few SNPs -> coarse distances -> equal keys straddle the "first five neighbours" cut; nodes with more than
    32 neighbours take the shared-memory counting path of the clustering kernel; error rates above 0.5 make the
    zero distances linkable"""
    rng = np.random.default_rng(17)
    total_replayed = 0
    for n_reads, n_snps, err in [(40, 6, 0.1), (90, 9, 0.15), (33, 4, 0.3), (64, 12, 0.6), (200, 5, 0.2), (1, 3, 0.1), (2, 3, 0.1)]:
        col = cases.coarse_columns(rng, n_reads, n_snps)
        keep = max(1, n_reads - 5)
        masked = np.sort(rng.choice(n_reads, size=keep, replace=False)).astype(np.int32)
        windows = [(0, masked, list(range(n_snps))), (0, masked[: max(1, keep // 2)], [0, 1])]
        _, replayed, _ = _check_batch(ctx, oracle, [col], windows, err, [PIN_SEED])
        total_replayed += replayed
    assert total_replayed > 0


def test_empty_batches(ctx, oracle):
    rng = np.random.default_rng(3)
    col = cases.coarse_columns(rng, 20, 4)
    pairs = api.Pairs(ctx, [col])
    pairs.compute()
    g = api.Graph(pairs, [], 0.1)
    g.build()
    adj_off, adj = g.adjacency()
    assert adj_off.tolist() == [0] and adj.size == 0
    g.close()
    g = api.Graph(pairs, [(0, np.zeros(0, np.int32)), (0, np.array([3], np.int32))], 0.1)
    g.build()
    adj_off, adj = g.adjacency()
    assert adj_off.tolist() == [0, 0] and adj.size == 0
    out = g.whispers([1], np.array([0], np.int32), np.arange(20, dtype=np.int32))
    assert out.tolist() == [0]
    g.close()
    pairs.close()


def _gro_pair(tmp, chunks, err, low="0", rare="0", amp="0", ploidy=None, threads="4", env=None):
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
    e = dict(os.environ, HS_PIN_SEED=str(PIN_SEED), **(env or {}))
    subprocess.run([OURS, col, threads, err, pl, low, rare, amp, ours, "0"], check=True, stdout=subprocess.DEVNULL, env=e)
    return open(ref, "rb").read(), open(ours, "rb").read()


@pytest.mark.skipif(not os.path.exists(REF_SR), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["ont_multi", "hifi", "low_memory", "rarest", "ploidy", "coverage_over_1000", "amplicon"])
def test_gro_identical_to_pinned_reference(tmp_path, case):
    assert os.path.exists(OURS), "build hairsplitter_b200/host first (python -c 'import __graft_entry__ as g; g.build()')"
    kw = {}
    if case == "ont_multi":
        chunks = [cases.small_case(seed=102, length=40000, depth=70, mean_len=7000, error=0.06),
                  cases.small_case(seed=101, length=60000, depth=60, mean_len=9000, error=0.10, n_strains=2),
                  cases.small_case(seed=5, length=3000, depth=12, mean_len=900, hard=0.4),
                  cases.small_case(seed=6, length=90, depth=5, mean_len=60)]
        err = "0.06"
    elif case == "hifi":
        chunks, err = [cases.hifi_case()], "0.01"
    elif case == "low_memory":
        chunks, err, kw = [cases.small_case(seed=103, length=12000, depth=40, mean_len=4000, error=0.06)], "0.06", dict(low="1")
    elif case == "rarest":
        chunks, err, kw = [cases.small_case(seed=104, length=30000, depth=50, mean_len=5000, error=0.06)], "0.06", dict(rare="0.2")
    elif case == "ploidy":
        chunks, err, kw = [cases.small_case(seed=105, length=30000, depth=60, mean_len=5000, error=0.05)], "0.05", dict(ploidy=2)
    elif case == "coverage_over_1000":
        chunks, err = [cases.small_case(seed=107, length=1500, depth=1500, mean_len=1400, error=0.05)], "0.05"
    else:
        chunks = [cases.small_case(seed=108, length=2500, depth=300, mean_len=2400, error=0.05),
                  cases.small_case(seed=109, length=1200, depth=30, mean_len=800, error=0.05, n_strains=2)]
        err, kw = "0.05", dict(amp="1")
    ref, ours = _gro_pair(str(tmp_path), chunks, err, **kw)
    assert ref == ours
    assert ref.count(b"GROUP") >= 1


@pytest.mark.skipif(not os.path.exists(REF_SR), reason="oracle/_ref not built")
def test_two_gpus_give_the_same_gro(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    chunks = [cases.small_case(seed=102, length=40000, depth=70, mean_len=7000, error=0.06),
              cases.small_case(seed=101, length=60000, depth=60, mean_len=9000, error=0.10, n_strains=2), cases.hifi_case()]
    ref, ours = _gro_pair(str(tmp_path), chunks, "0.06", env={"HSGPU_NGPUS": "2"})
    assert ref == ours


@pytest.mark.parametrize("case", ["ont", "lowmem", "amplicon"])
def test_gro_matches_golden(tmp_path, case):
    """committed fixtures (tests/golden/make_golden_sr.py: the reference's .col and the RNG-pinned reference's .gro):
    the GPU executable reproduces the .gro byte for byte without oracle/_ref at run time"""
    import gzip
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_sr
    _, err, low, rare, amp = make_golden_sr.CASES[case]
    g = os.path.join(ROOT, "tests", "golden")
    tmp = str(tmp_path)
    col, out = os.path.join(tmp, case + ".col"), os.path.join(tmp, "ours.gro")
    with open(col, "wb") as f:
        f.write(gzip.open(os.path.join(g, f"sr_{case}.col.gz")).read())
    want = gzip.open(os.path.join(g, f"sr_{case}.gro.gz")).read()
    subprocess.run([OURS, col, "4", err, os.path.join(tmp, "no_ploidy"), low, rare, amp, out, "0"], check=True,
                   stdout=subprocess.DEVNULL, env=dict(os.environ, HS_PIN_SEED=str(PIN_SEED)))
    assert open(out, "rb").read() == want


def test_unpinned_run_is_a_valid_clustering(tmp_path):
    """without HS_PIN_SEED the sweep orders are random (as in the reference): same windows and reads, labels may differ"""
    chunks = [cases.small_case(seed=102, length=40000, depth=70, mean_len=7000, error=0.06)]
    chunks[0].name = "ctg0"
    tmp = str(tmp_path)
    gfa, reads, sam = synth.write_files(chunks, os.path.join(tmp, "in"))
    col = os.path.join(tmp, "a.col")
    subprocess.run([REF_CV, gfa, reads, sam, "4", tmp, os.path.join(tmp, "err"), "0", "0", col, os.path.join(tmp, "a.vcf"), "0.33"],
                   check=True, stdout=subprocess.DEVNULL)
    env = {k: v for k, v in os.environ.items() if k != "HS_PIN_SEED"}
    out = {}
    for tag, extra in (("free", {}), ("pin", {"HS_PIN_SEED": str(PIN_SEED)})):
        path = os.path.join(tmp, tag + ".gro")
        subprocess.run([OURS, col, "4", "0.06", "none", "0", "0", "0", path, "0"], check=True, stdout=subprocess.DEVNULL,
                       env=dict(env, **extra))
        out[tag] = [l.split("\t") for l in open(path) if l.startswith("GROUP")]
    assert len(out["free"]) == len(out["pin"]) > 5
    for a, b in zip(out["free"], out["pin"]):
        assert a[:4] == b[:4]  # window bounds and the reads present


def test_usage_and_help_status():
    r = subprocess.run([OURS, "--help"], stdout=subprocess.PIPE)
    assert r.returncode == 0 and b"Usage" in r.stdout
    r = subprocess.run([OURS], stdout=subprocess.PIPE)
    assert r.returncode == 1


def _lowmem_check(ctx, col, masked_list, err):
    """device neighbour lists (hsgpu_graph_create_ex, low-memory windows) against the host path's
    create_read_graph_low_memory (hairsplitter_b200/host) and, when it travelled, the reference's own function"""
    from oracle import pyoracle
    pairs = api.Pairs(ctx, [col])
    pairs.compute()
    g = api.Graph(pairs, [(0, m) for m in masked_list], err, low_memory=[1] * len(masked_list))
    g.build()
    adj_off, adj = g.adjacency()
    base = 0
    for masked in masked_list:
        h_off, h_adj, consecutive = api.HostLogic.read_graph_low_memory(col, masked, err)
        assert consecutive
        m = masked.size
        assert np.array_equal(adj_off[base:base + m + 1] - adj_off[base], h_off)
        assert np.array_equal(adj[adj_off[base]:adj_off[base + m]], h_adj)
        if pyoracle.RefSR.available():
            r_off, r_adj = pyoracle.RefSR.read_graph_low_memory(col, masked, err)
            assert np.array_equal(h_off, r_off) and np.array_equal(h_adj, r_adj)
        base += m
    n_links, replayed = adj.size, g.replayed
    g.close()
    pairs.close()
    return n_links, replayed


def test_low_memory_read_graph_on_the_device(ctx, oracle):
    """a15: create_read_graph_low_memory (src/separate_reads.cpp:538-693) from the tensor-core pair counts: windows of
    an ONT contig, an amplicon-like column set (every read on every SNP, coarse distances: ties and NaN distances of
    pairs without a common SNP are replayed with the reference's sort), reads without any SNP cell"""
    cb = cases.small_case(seed=291, length=20000, depth=50, mean_len=5000, error=0.06)
    col, pos = cases.snp_columns(oracle, cb, 0.06)
    wins = [m for m, _ in cases.windows_of(col, pos)]
    links, _ = _lowmem_check(ctx, col, wins, 0.06)
    assert links > 0
    rng = np.random.default_rng(23)
    total_replayed = 0
    for n_reads, n_snps, err in [(60, 8, 0.1), (200, 5, 0.2), (33, 4, 0.3), (300, 8, 0.1)]:
        c = cases.coarse_columns(rng, n_reads, n_snps)
        masked = np.sort(rng.choice(n_reads, size=n_reads - 3, replace=False)).astype(np.int32)
        _, rep = _lowmem_check(ctx, c, [masked, masked[: n_reads // 2]], err)
        total_replayed += rep
    # reads 0..4 appear in no SNP column (mask_extend drops them), reads 5..9 only on the first SNP
    n_reads, n_snps = 50, 6
    c = list(cases.coarse_columns(rng, n_reads, n_snps))
    off, idx, code = c[1], c[2], c[3]
    keep = np.ones(idx.size, bool)
    for s in range(n_snps):
        seg = slice(int(off[s]), int(off[s + 1]))
        keep[seg] &= idx[seg] >= 5
        if s > 0:
            keep[seg] &= idx[seg] >= 10
    new_off = np.zeros(n_snps + 1, np.int64)
    for s in range(n_snps):
        new_off[s + 1] = new_off[s] + int(keep[int(off[s]):int(off[s + 1])].sum())
    c = (n_reads, new_off, idx[keep], code[keep], c[4], c[5])
    _lowmem_check(ctx, c, [np.arange(n_reads, dtype=np.int32), np.arange(0, n_reads, 2, dtype=np.int32)], 0.1)


def test_low_memory_executable_device_and_host_lists_agree(tmp_path):
    """HS_separate_reads -l on the low-memory golden: device neighbour lists (default) and HS_LOWMEM_HOST=1 give the
    same .gro, which is the RNG-pinned reference's"""
    import gzip
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_sr
    _, err, low, rare, amp = make_golden_sr.CASES["lowmem"]
    assert low == "1"
    gold = os.path.join(ROOT, "tests", "golden")
    col = os.path.join(str(tmp_path), "in.col")
    with gzip.open(os.path.join(gold, "sr_lowmem.col.gz"), "rb") as f, open(col, "wb") as o:
        o.write(f.read())
    want = gzip.open(os.path.join(gold, "sr_lowmem.gro.gz"), "rb").read()
    for tag, extra in (("dev", {}), ("host", {"HS_LOWMEM_HOST": "1"})):
        gro = os.path.join(str(tmp_path), tag + ".gro")
        env = dict(os.environ, HS_PIN_SEED=str(PIN_SEED), HS_TIMING="1", **extra)
        r = subprocess.run([OURS, col, "4", err, os.path.join(str(tmp_path), "nop"), low, rare, amp, gro, "0"], check=True,
                           env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        assert open(gro, "rb").read() == want, tag
        assert ("low-memory read graphs" in r.stderr.decode()) == (tag == "dev")


@pytest.mark.skipif(not os.path.exists(REF_SR), reason="oracle/_ref not built")
def test_both_executables_chained_through_the_sidecar(tmp_path):
    """f2: our HS_call_variants leaves <col>.hsb next to the .col; our HS_separate_reads reads it instead of the text.
    The .gro must be the pinned reference's on the same .col, with the sidecar, without it (HS_SIDECAR=0), and with the
    rarest-strain filter of parse_column_file (src/separate_reads.cpp:167) applied to the arrays."""
    ours_cv = os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_call_variants")
    chunks = [cases.small_case(seed=104, length=30000, depth=50, mean_len=5000, error=0.06),
              cases.small_case(seed=101, length=60000, depth=60, mean_len=9000, error=0.10, n_strains=2)]
    for i, c in enumerate(chunks):
        c.name = f"ctg{i}"
    tmp = str(tmp_path)
    gfa, reads, sam = synth.write_files(chunks, os.path.join(tmp, "in"))
    col = os.path.join(tmp, "ours.col")
    subprocess.run([ours_cv, gfa, reads, sam, "4", tmp, os.path.join(tmp, "err"), "0", "0", col, os.path.join(tmp, "ours.vcf"), "0.33"],
                   check=True, stdout=subprocess.DEVNULL)
    assert os.path.getsize(col + ".hsb") > 1000
    for rare in ("0", "0.2"):
        ref = os.path.join(tmp, f"ref_{rare}.gro")
        subprocess.run([REF_SR, col, "1", "0.06", "none", "0", rare, "0", ref, "0"], check=True, stdout=subprocess.DEVNULL)
        want = open(ref, "rb").read()
        assert want.count(b"GROUP") > 5
        for tag, extra in (("sidecar", {}), ("text", {"HS_SIDECAR": "0"})):
            gro = os.path.join(tmp, f"{tag}_{rare}.gro")
            env = dict(os.environ, HS_PIN_SEED=str(PIN_SEED), HS_TIMING="1", **extra)
            r = subprocess.run([OURS, col, "4", "0.06", "none", "0", rare, "0", gro, "0"], check=True, env=env,
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
            assert ("binary sidecar" in r.stderr.decode()) == (tag == "sidecar")
            assert open(gro, "rb").read() == want, (tag, rare)


GLUED_SR = os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_glued")


@pytest.mark.skipif(not os.path.exists(GLUED_SR), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["ont", "amplicon"])
def test_reference_main_on_libhsgpu_gives_the_reference_gro(tmp_path, case):
    """INTEGRATION.md section 4 compiled and run: the reference's unmodified main() of HS_separate_reads (from
    libhsref_sr_open.so, random_device pinned) with list_similarities_and_differences_between_reads3 bound to
    integration/glue_separate_reads.cpp, i.e. to hsgpu_read_pair_counts -- the tcgen05 kernel -- through the C ABI.
    The .gro is the committed golden one of the pinned reference."""
    import gzip
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_sr
    _, err, low, rare, amp = make_golden_sr.CASES[case]
    g = os.path.join(ROOT, "tests", "golden")
    tmp = str(tmp_path)
    col, out = os.path.join(tmp, case + ".col"), os.path.join(tmp, "glued.gro")
    with open(col, "wb") as f:
        f.write(gzip.open(os.path.join(g, f"sr_{case}.col.gz")).read())
    want = gzip.open(os.path.join(g, f"sr_{case}.gro.gz")).read()
    subprocess.run([GLUED_SR, col, "1", err, os.path.join(tmp, "no_ploidy"), low, rare, amp, out, "0"], check=True,
                   stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == want
