"""The binary sidecar of the .col file (hairsplitter_b200/host/hs_colbin.h) on the CPU.

HS_separate_reads may read "<col>.hsb" instead of tokenising the text. It must leave exactly the structures
parse_column_file (reference src/separate_reads.cpp:46-190) builds from the text, for every value of the two parse
filters (max_coverage :157, rarest strain abundance :167), and it must never be used for a .col it does not belong to.
The .col files are the committed golden ones, written by the reference's own HS_call_variants
(tests/golden/make_golden_sr.py).
"""
import ctypes as C
import gzip
import os
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "hairsplitter_b200", "libhshost.so")
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = ["ont", "lowmem", "amplicon"]
NO_LIMIT = 2147483647


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(HOSTLIB):
        import __graft_entry__

        __graft_entry__.build()
    L = C.CDLL(HOSTLIB)
    L.hshost_col_digest.restype = C.c_double
    L.hshost_col_digest.argtypes = [C.c_char_p, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_uint64)]
    L.hshost_rewrite_col.restype = C.c_int
    L.hshost_rewrite_col.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    L.hshost_col_sidecar_enabled.restype = C.c_int
    return L


def digest(lib, path, route, max_coverage=NO_LIMIT, rarest=0.0):
    out = (C.c_uint64 * 5)()
    dt = lib.hshost_col_digest(path.encode(), max_coverage, rarest, route, out)
    return dt, tuple(out)


def golden_col(tmp_path, case):
    p = str(tmp_path / f"{case}.ref.col")
    with gzip.open(os.path.join(GOLD, f"sr_{case}.col.gz"), "rb") as f, open(p, "wb") as o:
        o.write(f.read())
    return p


def blocks(path):
    """the CONTIG blocks of a .col file as a sorted list (the writers order them by their own containers)"""
    text = open(path, "rb").read()
    parts = text.split(b"CONTIG\t")
    return sorted(b"CONTIG\t" + p for p in parts[1:])


def rewritten(lib, tmp_path, case):
    ref = golden_col(tmp_path, case)
    col = str(tmp_path / f"{case}.col")
    assert lib.hshost_rewrite_col(ref.encode(), col.encode(), str(tmp_path / f"{case}.vcf").encode()) == 0
    return ref, col


@pytest.mark.parametrize("case", CASES)
def test_writer_text_unchanged_and_sidecar_written(lib, tmp_path, case):
    ref, col = rewritten(lib, tmp_path, case)
    assert blocks(col) == blocks(ref)  # write_outputs still prints what the reference printed
    assert os.path.exists(col + ".hsb")
    assert not os.path.exists(col + ".hsb.tmp")
    assert not os.path.exists(ref + ".hsb")  # a .col from the reference's executable has none ...
    dt, d = digest(lib, ref, 1)
    assert dt < 0  # ... so the sidecar route refuses
    _, d2 = digest(lib, ref, 2)
    assert d2[4] == 0  # and HS_separate_reads takes the text


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("max_coverage,rarest", [(NO_LIMIT, 0.0), (1000000000, 0.05), (NO_LIMIT, 0.2), (25, 0.0), (3, 0.4), (0, 0.0)])
def test_sidecar_parse_equals_text_parse(lib, tmp_path, case, max_coverage, rarest):
    _, col = rewritten(lib, tmp_path, case)
    _, text = digest(lib, col, 0, max_coverage, rarest)
    dt, side = digest(lib, col, 1, max_coverage, rarest)
    assert dt >= 0 and side[4] == 1
    assert side[:4] == text[:4]
    assert text[1] > 0 and (text[2] > 0 or rarest > 0 or max_coverage == 0)


def test_sidecar_same_structures_as_reference_text(lib, tmp_path):
    # the rewritten text parses like the reference's own file (blocks may come in another order: compare block sets
    # above, counts here)
    ref, col = rewritten(lib, tmp_path, "ont")
    _, a = digest(lib, ref, 0)
    _, b = digest(lib, col, 1)
    assert a[1:4] == b[1:4]


def test_stale_or_damaged_sidecar_is_ignored(lib, tmp_path):
    _, col = rewritten(lib, tmp_path, "ont")
    _, text = digest(lib, col, 0)
    side = col + ".hsb"
    good = open(side, "rb").read()

    # the .col is replaced by another file of another size
    other = golden_col(tmp_path, "lowmem")
    moved = str(tmp_path / "moved.col")
    shutil.copyfile(other, moved)
    shutil.copyfile(side, moved + ".hsb")
    dt, _ = digest(lib, moved, 1)
    assert dt < 0
    _, d = digest(lib, moved, 2)
    _, want = digest(lib, other, 0)
    assert d[4] == 0 and d[:4] == want[:4]

    # same bytes, new modification time (a copy): not trusted either
    copy = str(tmp_path / "copy.col")
    shutil.copyfile(col, copy)
    os.utime(copy, ns=(1, 1))
    shutil.copyfile(side, copy + ".hsb")
    assert digest(lib, copy, 1)[0] < 0
    _, d = digest(lib, copy, 2)
    assert d[4] == 0 and d[:4] == text[:4]

    # an edited byte in the first 64 KiB with size and time kept
    st = os.stat(col)
    raw = bytearray(open(col, "rb").read())
    i = raw.index(b"SNPS\t") + 5
    raw[i] = ord("1") if raw[i] != ord("1") else ord("2")
    open(col, "wb").write(bytes(raw))
    os.utime(col, ns=(st.st_atime_ns, st.st_mtime_ns))
    assert digest(lib, col, 1)[0] < 0

    # truncated and corrupted sidecars
    for damage in ("truncate", "magic", "counts", "offsets"):
        fresh = str(tmp_path / f"{damage}.col")
        shutil.copyfile(copy, fresh)
        st = os.stat(fresh)
        b = bytearray(good)
        # re-stamp the header with this file's time so that only the damage decides (size and hash already agree)
        b[16:24] = int(st.st_mtime_ns).to_bytes(8, "little")
        ok_side = bytes(b)
        open(fresh + ".hsb", "wb").write(ok_side)
        dt, d = digest(lib, fresh, 1)
        assert dt >= 0 and d[:4] == text[:4], "the re-stamped sidecar must be accepted before it is damaged"
        if damage == "truncate":
            b = b[: len(b) // 2]
        elif damage == "magic":
            b[0] ^= 0xFF
        elif damage == "counts":
            first_block = int.from_bytes(b[40:48], "little")
            b[first_block + 8 : first_block + 16] = (int.from_bytes(b[first_block + 8 : first_block + 16], "little") + 1).to_bytes(8, "little")
        elif damage == "offsets":
            first_block = int.from_bytes(b[40:48], "little")
            head, n_snps = (int.from_bytes(b[first_block + 8 * k : first_block + 8 * k + 8], "little") for k in (0, 1))
            pad8 = lambda n: (n + 7) & ~7
            off_at = first_block + 24 + pad8(head) + pad8(4 * n_snps) + 2 * pad8(n_snps)
            b[off_at + 8 : off_at + 16] = (1 << 40).to_bytes(8, "little")
        open(fresh + ".hsb", "wb").write(bytes(b))
        assert digest(lib, fresh, 1)[0] < 0, damage
        _, d = digest(lib, fresh, 2)
        assert d[4] == 0 and d[:4] == text[:4], damage


def test_switch_off(lib, tmp_path, monkeypatch):
    monkeypatch.setenv("HS_SIDECAR", "0")
    assert lib.hshost_col_sidecar_enabled() == 0
    ref, col = rewritten(lib, tmp_path, "ont")
    assert not os.path.exists(col + ".hsb")
    assert blocks(col) == blocks(ref)
    monkeypatch.delenv("HS_SIDECAR")
    assert lib.hshost_col_sidecar_enabled() == 1


def test_writer_prints_every_number_width(lib, tmp_path):
    """write_outputs counts the bytes of a block before it formats it: numbers of every width (1 to 10 digits for read
    indices and positions, 1 to 3 for codes) must come out as the decimal text the reference's stream would print"""
    idx = [0, 9, 10, 99, 100, 999, 1000, 9999, 10000, 99999, 100000, 999999, 1000000, 9999999, 10000000, 99999999,
           100000000, 999999999, 1000000000, 2147483647]
    codes = [33, 99, 100, 157, 255, 0, 1, 9, 10, 200] * 2
    lines = ["CONTIG\tctgA\t2000\t12.5", "READ\tr0\t0\t10\t5\t1999999999\t1", "READ\tr1\t3\t2147483647\t0\t7\t0"]
    for pos in (0, 7, 10, 123456, 1999999999):
        lines.append("SNPS\t%d\t%d\t%d\t%s\t%s" % (pos, 33 + pos % 100, 255 - pos % 100, "".join("%d," % i for i in idx),
                                                 "".join("%d," % c for c in codes)))
    text = ("\n".join(lines) + "\n\n").encode()
    src, out = str(tmp_path / "in.col"), str(tmp_path / "out.col")
    open(src, "wb").write(text)
    assert lib.hshost_rewrite_col(src.encode(), out.encode(), str(tmp_path / "out.vcf").encode()) == 0
    assert open(out, "rb").read() == text
    _, a = digest(lib, out, 0)
    dt, b = digest(lib, out, 1)
    assert dt >= 0 and a[:4] == b[:4] and a[3] == 5 * len(idx)


@pytest.mark.parametrize("seed", range(12))
def test_random_ragged_files_parse_the_same_from_text_and_sidecar(lib, tmp_path, seed):
    """seeded random .col files with the shapes the golden ones do not have: contigs without reads or without SNPS
    lines, SNPS lines without cells, single-cell columns, codes over the whole byte range (a code 32 is a blank for
    parse_column_file, :150-157, and is dropped), large read indices, several filter settings per file"""
    import numpy as np
    rng = np.random.default_rng(1000 + seed)
    lines = []
    for c in range(int(rng.integers(1, 6))):
        n_reads = int(rng.integers(0, 40))
        lines.append("CONTIG\tctg%d_%d\t%d\t%.4g" % (seed, c, int(rng.integers(50, 5000)), float(rng.uniform(0, 300))))
        for r in range(n_reads):
            a, b = sorted(int(x) for x in rng.integers(0, 5000, 2))
            lines.append("READ\tread_%d\t%d\t%d\t%d\t%d\t%d" % (r, int(rng.integers(0, 100)), int(rng.integers(100, 9000)), a, b, int(rng.integers(0, 2))))
        pos = 0
        for s in range(int(rng.integers(0, 30)) if n_reads else 0):
            pos += int(rng.integers(1, 200))
            k = int(rng.choice([0, 1, 2, n_reads])) if rng.random() < 0.3 else int(rng.integers(0, n_reads + 1))
            idx = np.sort(rng.choice(n_reads, size=k, replace=False)) if k else np.zeros(0, int)
            if k and rng.random() < 0.1:
                idx = idx.astype(np.int64) + int(rng.integers(0, 2**31 - 1 - n_reads))
            palette = rng.choice(np.arange(0, 256), size=int(rng.integers(1, 5)), replace=False)
            if rng.random() < 0.2:
                palette[0] = 32
            codes = rng.choice(palette, size=k)
            ref, sec = (int(x) for x in rng.choice(palette, 2))
            lines.append("SNPS\t%d\t%d\t%d\t%s\t%s" % (pos, ref, sec, "".join("%d," % i for i in idx), "".join("%d," % x for x in codes)))
        lines.append("")
    src, out = str(tmp_path / "in.col"), str(tmp_path / "out.col")
    open(src, "wb").write(("\n".join(lines) + "\n").encode())
    assert lib.hshost_rewrite_col(src.encode(), out.encode(), str(tmp_path / "out.vcf").encode()) == 0
    assert os.path.exists(out + ".hsb")
    for max_coverage, rarest in [(NO_LIMIT, 0.0), (5, 0.0), (NO_LIMIT, 0.3), (1, 0.5)]:
        _, want = digest(lib, src, 0, max_coverage, rarest)      # the file as generated, text route
        _, text = digest(lib, out, 0, max_coverage, rarest)      # the drop-in writer's text of the same content
        dt, side = digest(lib, out, 1, max_coverage, rarest)     # and its sidecar
        assert dt >= 0 and side[4] == 1
        assert side[:4] == text[:4]
        assert text[1:4] == want[1:4]


@pytest.mark.timeout(60)
def test_col_written_to_a_pipe(lib, tmp_path):
    """a <col_out> that is not a regular file (the reference writes through a stream, which does not care): the blocks
    go out one after the other instead of at file offsets, and no sidecar is attempted (its identity check would wait
    on the pipe for ever)"""
    import threading
    ref = golden_col(tmp_path, "ont")
    fifo = str(tmp_path / "out.col")
    os.mkfifo(fifo)
    got = []
    reader = threading.Thread(target=lambda: got.append(open(fifo, "rb").read()), daemon=True)
    reader.start()
    assert lib.hshost_rewrite_col(ref.encode(), fifo.encode(), str(tmp_path / "out.vcf").encode()) == 0
    reader.join(30)
    assert got and sorted(got[0].split(b"CONTIG\t")) == sorted(open(ref, "rb").read().split(b"CONTIG\t"))
    assert not os.path.exists(fifo + ".hsb")
