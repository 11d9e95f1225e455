"""CPU tests of the host side of libhsgpu: the library loads without a GPU, exports every symbol the
header declares, fails loudly instead of falling back, and its ranking code (rank.cuh, the same source
the kernels compile) agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hairsplitter_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = api.load()
    header = open(os.path.join(ROOT, "include", "hsgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(hsgpu_[A-Za-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(api.EXPORTS)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.HsgpuError):
        api.Context(0)


def test_pack_bases_and_parse_cigar():
    lib = api.load()
    seq = b"ACGTNACGTTTGACCAGTNNA"
    out = np.zeros(2, np.uint32)
    lib.hsgpu_pack_bases_ascii(seq, len(seq), out.ctypes.data)
    codes = np.array([{65: 0, 67: 1, 71: 2}.get(c, 3) for c in seq], dtype=np.uint8)
    assert np.array_equal(out, api.pack_codes(codes))
    out2 = np.zeros(2, np.uint32)
    lib.hsgpu_pack_bases_codes(codes.ctypes.data, len(seq), out2.ctypes.data)
    assert np.array_equal(out, out2)
    unpacked = [(int(out[j // 16]) >> (2 * (j % 16))) & 3 for j in range(len(seq))]
    assert unpacked == list(codes)
    ops = np.zeros(16, np.uint32)
    n = lib.hsgpu_parse_cigar(b"12S30M2I1D5=1X7H", ops.ctypes.data, 16)
    assert n == 7
    assert [(int(o) >> 4, "MIDNSHP=X"[int(o) & 15]) for o in ops[:n]] == [(12, "S"), (30, "M"), (2, "I"), (1, "D"),
                                                                           (5, "="), (1, "X"), (7, "H")]
    assert lib.hsgpu_parse_cigar(b"*", ops.ctypes.data, 16) == 0
    assert lib.hsgpu_parse_cigar(b"5M", ops.ctypes.data, 0) == -4
    assert lib.hsgpu_parse_cigar(b"5Q", ops.ctypes.data, 16) == -3


def test_pack_cigar8():
    """the 8-bit CIGAR form: four op classes, ops longer than 63 split, N/P refused"""
    lib = api.load()
    ops = np.zeros(16, np.uint32)
    n = lib.hsgpu_parse_cigar(b"12S130M2I1D5=63X0M7H", ops.ctypes.data, 16)
    assert n == 8
    out = np.zeros(32, np.uint8)
    m = lib.hsgpu_pack_cigar8(ops.ctypes.data, n, out.ctypes.data, 32)
    want = [(12, 3), (63, 0), (63, 0), (4, 0), (2, 1), (1, 2), (5, 0), (63, 0), (0, 0), (7, 3)]
    assert m == len(want)
    assert [(int(b) >> 2, int(b) & 3) for b in out[:m]] == want
    assert lib.hsgpu_pack_cigar8(ops.ctypes.data, n, None, 0) == m          # counting only
    assert lib.hsgpu_pack_cigar8(ops.ctypes.data, n, out.ctypes.data, 3) == -4  # capacity
    n = lib.hsgpu_parse_cigar(b"5M3N5M", ops.ctypes.data, 16)
    assert lib.hsgpu_pack_cigar8(ops.ctypes.data, n, out.ctypes.data, 32) == -3  # N has no 8-bit form
    # numpy-side wrapper: offsets follow the splits
    cig = np.array([(200 << 4) | 0, (3 << 4) | 1, (64 << 4) | 2, (10 << 4) | 0], np.uint32)
    c8, off8 = api.cigar8(cig, np.array([0, 2, 2, 4], np.int64))
    assert off8.tolist() == [0, 5, 5, 8]
    assert [(int(b) >> 2, int(b) & 3) for b in c8] == [(63, 0), (63, 0), (63, 0), (11, 0), (3, 1), (63, 2), (1, 2), (10, 0)]


def test_mean_distance_float_semantics(oracle):
    lib = api.load()
    for d, a in [(0, 0), (5, 100), (1234567, 17000000), (16777216, 2 ** 27), (20000000, 2 ** 28)]:
        assert np.float32(lib.hsgpu_mean_distance(d, a)).tobytes() == np.float32(oracle.mean_distance(d, a)).tobytes()


def test_ranking_source_matches_oracle(oracle):
    """rank.cuh compiled for the host (the kernels compile the same header)"""
    lib = api.load()
    rng = np.random.default_rng(0)
    for it in range(3000):
        n = int(rng.integers(1, 40 if it % 20 else 126))
        keys = (rng.permutation(125)[:n] + 33).astype(np.uint8)
        out = np.zeros(140, np.uint8)
        k = lib.hsgpu_debug_rh_order(keys.ctypes.data, n, out.ctypes.data)
        assert np.array_equal(out[:k], oracle.rh_order(keys))
    for it in range(3000):
        n = int(rng.integers(1, 129))
        keys = rng.permutation(256)[:n].astype(np.uint8)
        counts = rng.integers(0, rng.integers(1, 8), n).astype(np.int32)
        k2, c2 = keys.copy(), counts.copy()
        lib.hsgpu_debug_sort_desc(k2.ctypes.data, c2.ctypes.data, n)
        a = oracle.sort_desc(keys, counts)
        assert np.array_equal(a[0], k2) and np.array_equal(a[1], c2)
    n_fast = n_slot = 0
    for it in range(6000):
        depth = int(rng.integers(0, 90))
        ncodes = int(rng.integers(1, 30 if it % 10 else 125))
        alphabet = (rng.permutation(125)[:ncodes] + 33).astype(np.uint8)
        w = rng.random(ncodes) ** 3
        col = rng.choice(alphabet, size=depth, p=w / w.sum()).astype(np.uint8)
        want = oracle.column_rank(col)
        # bucket-table fast path with fallbacks (what the kernels run), pure replay, slot-order path alone
        for mode in (0, 1, 2):
            out = np.zeros(6, np.int32)
            lib.hsgpu_debug_rank_column(col.ctypes.data, depth, out.ctypes.data, mode)
            assert np.array_equal(out[:5], want), (mode, col)
            n_fast += (mode == 0 and out[5] == 0)
            n_slot += (mode == 2 and out[5] == 0)
    assert n_fast > 2000  # the table decides the large majority of columns
    assert n_slot > 3000
    # deep, noisy columns (more than 13 codes: the sorted vector leaves libstdc++'s insertion-sort regime) with
    # small tied counts: the regime the slot-order path exists for
    n_slot = 0
    for it in range(6000):
        ncodes = int(rng.integers(14, 49))
        alphabet = (rng.permutation(125)[:ncodes] + 33).astype(np.uint8)
        reps = rng.integers(1, 4, ncodes)
        reps[0] += int(rng.integers(0, 40))
        col = rng.permutation(np.repeat(alphabet, reps)).astype(np.uint8)
        want = oracle.column_rank(col)
        for mode in (0, 2):
            out = np.zeros(6, np.int32)
            lib.hsgpu_debug_rank_column(col.ctypes.data, col.shape[0], out.ctypes.data, mode)
            assert np.array_equal(out[:5], want), (mode, col)
            n_slot += (mode == 2 and out[5] == 0)
    assert n_slot > 3000
    # tie-heavy columns: every code once or twice, so the rank is decided by the iteration order of the
    # reference's table alone (home bucket, then the hash bits kept in the info byte, then history)
    n_table = 0
    for it in range(8000):
        ncodes = int(rng.integers(2, 14))
        alphabet = (rng.permutation(125)[:ncodes] + 33).astype(np.uint8)
        col = rng.permutation(np.repeat(alphabet, rng.integers(1, 3, ncodes))).astype(np.uint8)
        want = oracle.column_rank(col)
        out = np.zeros(6, np.int32)
        lib.hsgpu_debug_rank_column(col.ctypes.data, col.shape[0], out.ctypes.data, 0)
        assert np.array_equal(out[:5], want), col
        n_table += out[5] == 0
    assert n_table > 6000


def test_synthetic_generator_is_self_consistent():
    """every CIGAR consumes exactly its read and stays inside the contig"""
    import cases
    for cb in (cases.small_case(), cases.hifi_case(), cases.small_case(seed=12, eqx=True)):
        rl = cb.read_len()
        for i in range(cb.n_reads):
            ops = cb.cigar[cb.cigar_off[i]:cb.cigar_off[i + 1]]
            ln, op = ops >> 4, ops & 15
            assert int(ln[np.isin(op, [0, 1, 4, 5, 7, 8])].sum()) == int(rl[i])
            assert cb.start[i] + int(ln[np.isin(op, [0, 2, 7, 8])].sum()) <= cb.length


def test_edlib_shim_structs_are_layout_compatible_with_edlib(tmp_path):
    """hsgpu_EdlibAlignConfig / hsgpu_EdlibAlignResult claim the layout of edlib's structs (a maintainer may cast):
    checked by the C compiler against the reference's own header, where the reference tree is present"""
    import subprocess
    ref_inc = "/root/reference/src/edlib/include"
    if not os.path.exists(os.path.join(ref_inc, "edlib.h")):
        pytest.skip("reference tree not present (GPU box)")
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stddef.h>
#include "edlib.h"
#include "hsgpu.h"
#define SAME(T, U, f) _Static_assert(offsetof(T, f) == offsetof(U, f) && sizeof(((T*)0)->f) == sizeof(((U*)0)->f), #f)
_Static_assert(sizeof(EdlibAlignConfig) == sizeof(hsgpu_EdlibAlignConfig), "config size");
_Static_assert(sizeof(EdlibAlignResult) == sizeof(hsgpu_EdlibAlignResult), "result size");
SAME(EdlibAlignConfig, hsgpu_EdlibAlignConfig, k);
SAME(EdlibAlignConfig, hsgpu_EdlibAlignConfig, mode);
SAME(EdlibAlignConfig, hsgpu_EdlibAlignConfig, task);
SAME(EdlibAlignConfig, hsgpu_EdlibAlignConfig, additionalEqualities);
SAME(EdlibAlignConfig, hsgpu_EdlibAlignConfig, additionalEqualitiesLength);
SAME(EdlibAlignResult, hsgpu_EdlibAlignResult, status);
SAME(EdlibAlignResult, hsgpu_EdlibAlignResult, editDistance);
SAME(EdlibAlignResult, hsgpu_EdlibAlignResult, endLocations);
SAME(EdlibAlignResult, hsgpu_EdlibAlignResult, startLocations);
SAME(EdlibAlignResult, hsgpu_EdlibAlignResult, numLocations);
SAME(EdlibAlignResult, hsgpu_EdlibAlignResult, alignment);
SAME(EdlibAlignResult, hsgpu_EdlibAlignResult, alignmentLength);
SAME(EdlibAlignResult, hsgpu_EdlibAlignResult, alphabetLength);
_Static_assert(EDLIB_MODE_NW == 0 && EDLIB_MODE_SHW == 1 && EDLIB_MODE_HW == 2, "modes");
_Static_assert(EDLIB_TASK_DISTANCE == 0 && EDLIB_TASK_LOC == 1 && EDLIB_TASK_PATH == 2, "tasks");
_Static_assert(EDLIB_STATUS_OK == 0 && EDLIB_STATUS_ERROR == 1, "status");
int main(void) { return 0; }
''')
    r = subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", ref_inc, "-I", os.path.join(ROOT, "include"), "-c", str(src),
                        "-o", str(tmp_path / "layout.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


GLUED = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants_glued")


@pytest.mark.skipif(not os.path.exists(GLUED), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_main_binds_to_the_glue(tmp_path):
    """integration/glue_call_variants.cpp compiled against the reference's headers: the reference's main() (in
    libhsref_cv.so) must reach the glue's generate_msa through the PLT. Without a GPU that shows as the library's loud
    failure right after the reference's own parsers have run -- and as the usage text with status 0 when called the
    way hairsplitter.py probes its dependencies (hairsplitter.py:229-239)."""
    import subprocess
    import sys
    import torch
    r = subprocess.run([GLUED, "--version"], stdout=subprocess.PIPE)
    assert r.returncode == 0 and b"Usage" in r.stdout
    syms = subprocess.run(["nm", "-D", "--defined-only", GLUED], stdout=subprocess.PIPE, text=True).stdout
    for name in ("generate_msa", "13call_variants", "keep_only_robust_variants"):
        assert any(name in line and " T " in line for line in syms.splitlines()), name
    if torch.cuda.is_available():
        pytest.skip("GPU present: tests/test_gpu_callvariants.py runs the glued executable for real")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from hairsplitter_b200 import synth
    chunk = cases.small_case(seed=91, length=4000, depth=10, mean_len=1500, error=0.06)
    chunk.name = "ctg0"
    tmp = str(tmp_path)
    gfa, reads, sam = synth.write_files([chunk], os.path.join(tmp, "in"))
    r = subprocess.run([GLUED, gfa, reads, sam, "1", tmp, tmp + "/err", "0", "0", tmp + "/a.col", tmp + "/a.vcf", "0.33"],
                       stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1
    assert "Calling variants on each contig" in r.stdout           # the reference's main() got that far by itself
    assert "hsgpu_ctx_create failed" in r.stdout and "no CPU fallback" in r.stdout  # ... and called into the glue


MOCK_DIR = os.path.join(ROOT, "oracle", "_ref", "mock_for_glue_test")
REF_CV = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")


@pytest.mark.skipif(not (os.path.exists(GLUED) and os.path.exists(os.path.join(MOCK_DIR, "libhsgpu.so")) and os.path.exists(REF_CV)),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_glue_logic_against_the_reference_with_the_oracle_behind_the_c_abi(tmp_path):
    """The glue itself, checked where there is no GPU: the glued executable runs with oracle/mock_hsgpu.c (the few entry
    points the glue calls, computed by the oracle with the semantics include/hsgpu.h documents) in front of the real
    library -- for this subprocess only. What is under test is the reference-side code: the conversions between
    Read / Overlap / Column / Partition and the flat arrays of the C ABI, the order of the calls, what the reference's
    main() gets back. The files must be the reference executable's, byte for byte (one thread), and block for block
    with two OpenMP threads. The same comparison with the real library is the GPU test
    tests/test_gpu_callvariants.py::test_reference_main_on_libhsgpu_gives_the_reference_files."""
    import filecmp
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from hairsplitter_b200 import synth
    chunks = [cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06), cases.hifi_case(),
              cases.small_case(seed=5, length=3000, depth=12, mean_len=900, hard=0.4)]
    for i, c in enumerate(chunks):
        c.name = f"ctg{i}"
    tmp = str(tmp_path)
    files = synth.write_files(chunks, os.path.join(tmp, "in"))

    def run(exe, tag, threads, env=None):
        out = [os.path.join(tmp, f"{tag}.{e}") for e in ("col", "vcf", "err")]
        subprocess.run([exe, *files, str(threads), tmp, out[2], "0", "0", out[0], out[1], "0.33"], check=True,
                       stdout=subprocess.DEVNULL, env=env)
        return out

    ref = run(REF_CV, "ref", 1)
    mock_env = dict(os.environ, LD_LIBRARY_PATH=MOCK_DIR)
    one = run(GLUED, "glued1", 1, mock_env)
    for a, b in zip(ref, one):
        assert filecmp.cmp(a, b, shallow=False), (a, b)
    assert open(ref[0], "rb").read().count(b"SNPS\t") > 20
    two = run(GLUED, "glued2", 2, mock_env)
    assert sorted(open(ref[0], "rb").read().split(b"CONTIG\t")) == sorted(open(two[0], "rb").read().split(b"CONTIG\t"))


GLUED_SR = os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_glued")


@pytest.mark.skipif(not (os.path.exists(GLUED_SR) and os.path.exists(os.path.join(MOCK_DIR, "libhsgpu.so"))),
                    reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", ["ont", "amplicon"])
def test_separate_reads_glue_logic_with_the_oracle_behind_the_c_abi(tmp_path, case):
    """integration/glue_separate_reads.cpp under the reference's own main() of HS_separate_reads (random_device pinned):
    list_similarities_and_differences_between_reads3 bound to hsgpu_read_pair_counts, here answered by the oracle
    (oracle/mock_hsgpu.c, this subprocess only). The .gro must be the committed golden one of the pinned reference.
    Without the stand-in the same executable stops in the glue: no GPU, no fallback. The run with the real library is
    tests/test_gpu_sepreads.py::test_reference_main_on_libhsgpu_gives_the_reference_gro."""
    import gzip
    import subprocess
    import sys
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_sr
    _, err, low, rare, amp = make_golden_sr.CASES[case]
    gold = os.path.join(ROOT, "tests", "golden")
    tmp = str(tmp_path)
    col, out = os.path.join(tmp, case + ".col"), os.path.join(tmp, case + ".gro")
    with open(col, "wb") as f:
        f.write(gzip.open(os.path.join(gold, f"sr_{case}.col.gz")).read())
    want = gzip.open(os.path.join(gold, f"sr_{case}.gro.gz")).read()
    cmd = [GLUED_SR, col, "1", err, os.path.join(tmp, "no_ploidy"), low, rare, amp, out, "0"]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, env=dict(os.environ, LD_LIBRARY_PATH=MOCK_DIR))
    assert open(out, "rb").read() == want
    if not torch.cuda.is_available():
        r = subprocess.run(cmd, stdout=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stdout


OURS_CV = os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_call_variants")
HOSTCHECK = os.path.join(ROOT, "oracle", "sr_hostcheck")
REF_SR_PINNED = os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_pinned")


@pytest.mark.skipif(not (os.path.exists(OURS_CV) and os.path.exists(os.path.join(MOCK_DIR, "libhsgpu.so")) and os.path.exists(REF_CV)),
                    reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", ["ont_multi", "hifi_fastq", "edges"])
def test_drop_in_executable_host_side_against_the_reference(tmp_path, case):
    """The HOST side of bin/HS_call_variants end to end where there is no GPU: parsers with parse_SAM's filters, 2-bit
    packing, 16-bit CIGAR, batching of the contigs, partition building (loops 1-2), the merge with the automatic SNPs,
    the .col / .vcf writers and the sidecar -- with oracle/mock_hsgpu.c answering the C-ABI calls in this one
    subprocess (LD_LIBRARY_PATH; the oracle computes what the kernels compute on a B200). Files byte-identical to the
    reference executable's; the cases are those of the GPU test of the same executable
    (tests/test_gpu_callvariants.py::test_col_vcf_error_rate_identical_to_reference). Then the chain: the .col and its
    sidecar go through the host pipeline of HS_separate_reads (oracle/sr_hostcheck: the product's pipeline with the GPU
    stages computed by the oracle) and the .gro must be the pinned reference's on the same .col."""
    import filecmp
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from hairsplitter_b200 import synth
    from oracle.pyoracle import PIN_SEED
    if case == "ont_multi":
        chunks = [cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06), cases.medium_case(),
                  cases.small_case(seed=5, length=3000, depth=12, mean_len=900, hard=0.4)]
        fastq = False
    elif case == "hifi_fastq":
        chunks = [cases.hifi_case(), cases.small_case(seed=12, eqx=True)]
        fastq = True
    else:
        chunks = cases.ragged_cases() + [cases.deep_case()]
        fastq = False
    for i, c in enumerate(chunks):
        c.name = f"ctg{i}"
    tmp = str(tmp_path)
    files = synth.write_files(chunks, os.path.join(tmp, "in"), fastq=fastq)

    def run(exe, tag, threads, env=None):
        out = [os.path.join(tmp, f"{tag}.{e}") for e in ("col", "vcf", "err")]
        subprocess.run([exe, *files, str(threads), tmp, out[2], "0", "0", out[0], out[1], "0.33"], check=True,
                       stdout=subprocess.DEVNULL, env=env)
        return out

    ref = run(REF_CV, "ref", 1)
    ours = run(OURS_CV, "ours", 4, dict(os.environ, LD_LIBRARY_PATH=MOCK_DIR))
    for a, b in zip(ref, ours):
        assert filecmp.cmp(a, b, shallow=False), (a, b)
    assert os.path.getsize(ref[0]) > 1000
    assert os.path.exists(ours[0] + ".hsb")
    if not (os.path.exists(HOSTCHECK) and os.path.exists(REF_SR_PINNED)):
        return
    err = open(ref[2]).read().split()[0]
    want, got = os.path.join(tmp, "ref.gro"), os.path.join(tmp, "ours.gro")
    subprocess.run([REF_SR_PINNED, ref[0], "1", err, "none", "0", "0", "0", want, "0"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([HOSTCHECK, ours[0], "2", err, "none", "0", "0", "0", got, "0"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, env=dict(os.environ, HS_PIN_SEED=str(PIN_SEED), HS_TIMING="1"))
    assert "binary sidecar" in r.stderr.decode()
    assert open(got, "rb").read() == open(want, "rb").read()


def test_host_packer_equals_the_library_packer():
    """hs::pack_bases_2bit (table-driven, what bin/HS_call_variants packs reads with) against hsgpu_pack_bases_ascii (the
    C-ABI function, src/sequence.cpp:13-23 semantics): identical words for arbitrary bytes and every length around the
    16-base word boundaries"""
    host = C.CDLL(os.path.join(ROOT, "hairsplitter_b200", "libhshost.so"))
    lib = api.load()
    for f in (host.hshost_pack_bases_2bit, lib.hsgpu_pack_bases_ascii):
        f.argtypes = [C.c_char_p, C.c_int64, C.c_void_p]
        f.restype = None
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 256, 4096, dtype=np.uint8).tobytes()
    acgt = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, 4096)].tobytes()
    for data in (raw, acgt):
        for n in list(range(0, 70)) + [4095, 4096]:
            a = np.full((n + 15) // 16 + 1, 0xDEADBEEF, np.uint32)
            b = a.copy()
            lib.hsgpu_pack_bases_ascii(data, n, a.ctypes.data)
            host.hshost_pack_bases_2bit(data, n, b.ctypes.data)
            assert np.array_equal(a, b), n


def test_host_cigar_parser_semantics():
    """cigar_ops of the drop-in's parsers (what HS_call_variants feeds the pileup with): lengths of any width, zero
    lengths dropped (no loop of the reference sees them), letters outside "MIDNSHP=X" as padding, "*" and "" give
    nothing; on ordinary strings the same ops as the C-ABI parser hsgpu_parse_cigar"""
    host = C.CDLL(os.path.join(ROOT, "hairsplitter_b200", "libhshost.so"))
    host.hshost_cigar_ops.argtypes = [C.c_char_p, C.c_void_p, C.c_int64]
    host.hshost_cigar_ops.restype = C.c_int64
    lib = api.load()
    lib.hsgpu_parse_cigar.argtypes = [C.c_char_p, C.c_void_p, C.c_int64]
    lib.hsgpu_parse_cigar.restype = C.c_int64

    def ops_of(text, fn=host.hshost_cigar_ops):
        out = np.zeros(len(text) + 8, np.uint32)
        k = fn(text.encode(), out.ctypes.data, out.size)
        return [(int(x) >> 4, int(x) & 15) for x in out[:k]]

    assert ops_of("5M0I3D007S2=1X9N4H2P3Z") == [(5, 0), (3, 2), (7, 4), (2, 7), (1, 8), (9, 3), (4, 5), (2, 6), (3, 6)]
    assert ops_of("*") == [] and ops_of("") == [] and ops_of("12") == []
    assert ops_of("123456789M1D") == [(123456789, 0), (1, 2)]
    assert ops_of("0000000012M") == [(12, 0)]  # ten digits: through std::stoi like the reference's own conversion
    rng = np.random.default_rng(9)
    for _ in range(20):
        text = "".join("%d%s" % (rng.integers(1, 5000), "MIDNSHP=X"[rng.integers(0, 9)]) for _ in range(int(rng.integers(1, 400))))
        assert ops_of(text) == ops_of(text, lib.hsgpu_parse_cigar)
