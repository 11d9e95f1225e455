"""CPU tests of the host side of the HS_call_variants drop-in (hairsplitter_b200/host): the partition
builder (loops 1+2 of keep_only_robust_variants) against the compiled reference, the chi-square, and
the parsers against the CONTIG/READ lines the reference executable writes. The reference objects come
from oracle/_ref (built from /root/reference by oracle/Makefile; they travel with the snapshot)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
from hairsplitter_b200 import synth
from oracle.pyoracle import RefCV

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "hairsplitter_b200", "libhshost.so")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")


@pytest.fixture(scope="module")
def hostlib():
    if not os.path.exists(HOSTLIB):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "hairsplitter_b200", "host")], check=True)
    L = C.CDLL(HOSTLIB)
    L.hshost_build_partitions.restype = C.c_void_p
    L.hshost_build_partitions.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_float]
    L.hshost_parts_count.argtypes = [C.c_void_p]
    L.hshost_part_size.argtypes = [C.c_void_p, C.c_int]
    L.hshost_part_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    L.hshost_parts_free.argtypes = [C.c_void_p]
    L.hshost_chi_square.restype = C.c_float
    L.hshost_chi_square.argtypes = [C.c_int] * 4
    L.hshost_parse_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p]
    return L


def _host_partitions(L, pile, sus, mean_error):
    pos = np.ascontiguousarray(sus["pos"], np.int32)
    off = np.zeros(pos.size + 1, np.int64)
    idx, code = [], []
    for i, q in enumerate(pos):
        a, b = pile["col_off"][q], pile["col_off"][q + 1]
        idx.append(pile["read_idx"][a:b])
        code.append(pile["code"][a:b])
        off[i + 1] = off[i] + (b - a)
    idx = np.ascontiguousarray(np.concatenate(idx) if idx else np.zeros(0), np.uint32)
    code = np.ascontiguousarray(np.concatenate(code) if code else np.zeros(0), np.uint8)
    rb = np.ascontiguousarray(sus["ref_base"], np.uint8)
    sb = np.ascontiguousarray(sus["second_base"], np.uint8)
    h = L.hshost_build_partitions(int(pos.size), off.ctypes.data, idx.ctypes.data, code.ctypes.data, pos.ctypes.data,
                                  rb.ctypes.data, sb.ctypes.data, float(mean_error))
    parts = []
    for p in range(L.hshost_parts_count(h)):
        n = L.hshost_part_size(h, p)
        a = [np.zeros(n, np.int32), np.zeros(n, np.int16), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(2, np.int32)]
        L.hshost_part_get(h, p, *[x.ctypes.data for x in a])
        parts.append(dict(read_idx=a[0], state=a[1], more=a[2], less=a[3], left=int(a[4][0]), right=int(a[4][1])))
    L.hshost_parts_free(h)
    return parts


@pytest.mark.skipif(not RefCV.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["small", "medium", "hifi", "three_strains"])
def test_partition_builder_matches_reference(hostlib, case):
    cb = {"small": cases.small_case, "medium": cases.medium_case, "hifi": cases.hifi_case,
          "three_strains": lambda: cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06)}[case]()
    ref = RefCV(cb)
    pile = ref.pileup()
    md = ref.mean_distance()
    cv = ref.call_variants(md)
    want, _, _ = ref.robust(md)
    got = _host_partitions(hostlib, pile, cv["suspects"], md)
    assert len(got) == len(want)
    assert len(want) > 0
    for g, w in zip(got, want):
        for k in ("read_idx", "state", "more", "less"):
            assert np.array_equal(g[k], w[k]), k
        assert (g["left"], g["right"]) == (w["left"], w["right"])
    ref.close()


@pytest.mark.skipif(not RefCV.available(), reason="oracle/_ref not built")
def test_host_chi_square_is_bit_exact(hostlib):
    rng = np.random.default_rng(3)
    tables = [(a, b, c, d) for a in range(4) for b in range(4) for c in range(4) for d in range(4)]
    tables += [tuple(int(x) for x in rng.integers(0, 80, 4)) for _ in range(5000)]
    for t in tables:
        assert np.float32(hostlib.hshost_chi_square(*t)).tobytes() == np.float32(RefCV.chi_square(*t)).tobytes(), t


def _write_inputs(tmp, chunks, fastq=False, clips=True):
    return synth.write_files(chunks, os.path.join(tmp, "in"), fastq=fastq)


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref not built")
@pytest.mark.parametrize("fastq", [False, True])
def test_parsers_match_reference_read_lines(hostlib, tmp_path, fastq):
    chunks = [cases.small_case(seed=5, length=3000, depth=12, mean_len=900, hard=0.4),
              cases.small_case(seed=6, length=1500, depth=8, mean_len=600, hard=0.4)]
    gfa, reads, sam = _write_inputs(str(tmp_path), chunks, fastq=fastq)
    col, vcf, err = [os.path.join(str(tmp_path), n) for n in ("ref.col", "ref.vcf", "ref.err")]
    subprocess.run([REF_EXE, gfa, reads, sam, "1", str(tmp_path), err, "0", "0", col, vcf, "0.33"], check=True,
                   stdout=subprocess.DEVNULL)
    dump = os.path.join(str(tmp_path), "ours.dump")
    assert hostlib.hshost_parse_dump(gfa.encode(), reads.encode(), sam.encode(), 0, dump.encode()) == 0
    want = {}
    for line in open(col):
        f = line.rstrip("\n").split("\t")
        if f[0] == "CONTIG":
            cur = want.setdefault((f[1], f[2]), [])
        elif f[0] == "READ":
            cur.append(tuple(f[1:]))
    got = {}
    for line in open(dump):
        f = line.rstrip("\n").split("\t")
        if f[0] == "CONTIG":
            cur = got.setdefault((f[1], f[2]), [])
        elif f[0] == "READ":
            cur.append(tuple(f[1:7]))
    assert got == want
    assert sum(len(v) for v in want.values()) > 10


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref not built")
@pytest.mark.parametrize("amplicon", [0, 1])
def test_sam_filters_and_index_quirks_match_reference(hostlib, tmp_path, amplicon):
    """unmapped / secondary / supplementary flags, hard-clip and NM limits, short lines, '*' CIGARs, unknown
    read names (the second record of an unknown read lands on read 0 in the reference) and blank lines"""
    chunks = [cases.small_case(seed=7, length=2500, depth=10, mean_len=700, hard=0.5),
              cases.small_case(seed=8, length=1200, depth=6, mean_len=500)]
    for i, c in enumerate(chunks):
        c.name = f"ctg{i}"
    gfa, reads, sam = synth.write_files(chunks, os.path.join(str(tmp_path), "in"))
    r0, r1 = "ctg0_r0", "ctg1_r1"
    extra = [
        f"{r0}\t4\tctg0\t10\t60\t50M\t*\t0\t0\t*\t*\tNM:i:0\tLN:i:700",            # unmapped
        f"{r0}\t256\tctg0\t10\t60\t50M\t*\t0\t0\t*\t*\tNM:i:0\tLN:i:700",          # secondary
        f"{r0}\t2048\tctg1\t5\t60\t400H30M2D10M300H\t*\t0\t0\t*\t*\tNM:i:2\tLN:i:740",  # supplementary keeps its H
        f"{r0}\t0\tctg1\t5\t60\t400H30M300H\t*\t0\t0\t*\t*\tNM:i:2\tLN:i:730",     # too much H for a primary record
        f"{r1}\t16\tctg0\t100\t60\t5S40M3I20M7S\t*\t0\t0\t*\t*\tNM:i:30\tLN:i:75",  # NM > 20 % (amplicon only)
        f"{r1}\t16\tctg0\t300\t60\t3H5S40M7S2H\t*\t0\t0\t*\t*\tNM:i:1\tLN:i:57",   # S hidden behind H at the end
        f"{r1}\t0\tctg0\t200\t60\t*\t*\t0\t0\t*\t*\tNM:i:0\tLN:i:500",              # no CIGAR
        f"{r1}\t0\tctg0\t200\t60\t20M\t*\t0\t0",                                     # too few fields
        "ghost_read\t0\tctg0\t50\t60\t30M\t*\t0\t0\t*\t*\tNM:i:0\tLN:i:30",          # unknown: ignored, registered
        "ghost_read\t0\tctg0\t60\t60\t30M\t*\t0\t0\t*\t*\tNM:i:0\tLN:i:30",          # ... and now it is read 0
        "",
        f"{r0}\t0\tctg0\t900\t60\t25M1I25M\t*\t0\t0\t*\t*\tNM:i:1\tLN:i:700\t",     # trailing tab
    ]
    with open(sam, "a") as f:
        f.write("\n".join(extra) + "\n")
    col, vcf, err = [os.path.join(str(tmp_path), n) for n in ("ref.col", "ref.vcf", "ref.err")]
    subprocess.run([REF_EXE, gfa, reads, sam, "1", str(tmp_path), err, str(amplicon), "0", col, vcf, "0.33"], check=True,
                   stdout=subprocess.DEVNULL)
    dump = os.path.join(str(tmp_path), "ours.dump")
    assert hostlib.hshost_parse_dump(gfa.encode(), reads.encode(), sam.encode(), amplicon, dump.encode()) == 0

    def blocks(path, ncol):
        out, cur = {}, None
        for line in open(path):
            f = line.rstrip("\n").split("\t")
            if f[0] == "CONTIG":
                cur = out.setdefault(f[1], [])
            elif f[0] == "READ":
                cur.append(tuple(f[1:ncol]))
        return out
    want, got = blocks(col, 7), blocks(dump, 7)
    assert got == want
    n_extra = sum(len(v) for v in want.values()) - sum(c.n_reads for c in chunks)
    assert n_extra == (5 if amplicon else 6)
