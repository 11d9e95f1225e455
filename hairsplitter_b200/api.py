"""ctypes binding of libhsgpu.so (include/hsgpu.h) for the Python test and benchmark harness.

The product is the C-ABI library; this module only marshals numpy arrays to it. It fails loudly when
the library is missing or when no sm_100 GPU is present -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhsgpu.so")

_lib = None


class HsgpuError(RuntimeError):
    pass


class PileupInput(C.Structure):
    _fields_ = [
        ("n_contigs", C.c_int32),
        ("contig_len", C.c_void_p),
        ("contig_bases", C.c_void_p),
        ("contig_word_off", C.c_void_p),
        ("contig_read_off", C.c_void_p),
        ("n_reads", C.c_int64),
        ("read_bases", C.c_void_p),
        ("read_word_off", C.c_void_p),
        ("read_len", C.c_void_p),
        ("cigar", C.c_void_p),
        ("cigar_off", C.c_void_p),
        ("read_start", C.c_void_p),
        ("read_strand", C.c_void_p),
        ("cigar16", C.c_void_p),
        ("cigar8", C.c_void_p),
    ]


class EdlibAlignConfig(C.Structure):
    """layout of EdlibAlignConfig (reference src/edlib/include/edlib.h:100-106)"""
    _fields_ = [("k", C.c_int), ("mode", C.c_int), ("task", C.c_int), ("additionalEqualities", C.c_void_p),
                ("additionalEqualitiesLength", C.c_int)]


class EdlibAlignResult(C.Structure):
    """layout of EdlibAlignResult (reference src/edlib/include/edlib.h:213-262)"""
    _fields_ = [("status", C.c_int), ("editDistance", C.c_int), ("endLocations", C.POINTER(C.c_int)),
                ("startLocations", C.POINTER(C.c_int)), ("numLocations", C.c_int),
                ("alignment", C.POINTER(C.c_ubyte)), ("alignmentLength", C.c_int), ("alphabetLength", C.c_int)]


class Partitions(C.Structure):
    _fields_ = [
        ("n_parts", C.c_int32),
        ("part_off", C.c_void_p),
        ("read_idx", C.c_void_p),
        ("state", C.c_void_p),
        ("more", C.c_void_p),
        ("less", C.c_void_p),
    ]


DISTANCE_DTYPE = np.dtype(
    [("n00", "<i4"), ("n01", "<i4"), ("n10", "<i4"), ("n11", "<i4"), ("solid00", "<i4"), ("solid01", "<i4"),
     ("solid10", "<i4"), ("solid11", "<i4"), ("second_base", "u1"), ("augmented", "u1"), ("pad", "u1", (2,)),
     ("chi_square", "<f4")]
)
EDLIB_RESULT_DTYPE = np.dtype(
    [("status", "<i4"), ("edit_distance", "<i4"), ("n_locations", "<i4"), ("alignment_length", "<i4"),
     ("alphabet_length", "<i4"), ("has_start_locations", "<i4"), ("loc_off", "<i8"), ("aln_off", "<i8")]
)

# every symbol include/hsgpu.h declares (tests check that the library exports all of them)
EXPORTS = [
    "hsgpu_ctx_create", "hsgpu_ctx_destroy", "hsgpu_last_error", "hsgpu_sync", "hsgpu_launch_count", "hsgpu_stream",
    "hsgpu_profile_enable", "hsgpu_profile_report", "hsgpu_host_alloc", "hsgpu_host_free", "hsgpu_pack_bases_ascii", "hsgpu_pack_bases_codes", "hsgpu_parse_cigar", "hsgpu_pack_cigar8",
    "hsgpu_pileup_create", "hsgpu_pileup_destroy", "hsgpu_pileup_build", "hsgpu_pileup_stats", "hsgpu_mean_distance",
    "hsgpu_pileup_read_ends", "hsgpu_pileup_export", "hsgpu_pileup_extract_columns", "hsgpu_column_rank",
    "hsgpu_column_counts", "hsgpu_suspects", "hsgpu_suspects_all", "hsgpu_column_summary", "hsgpu_partition_tables", "hsgpu_robust_filter",
    "hsgpu_partitions_set", "hsgpu_robust_filter_all", "hsgpu_pileup_info",
    "hsgpu_read_pair_counts", "hsgpu_pairs_create", "hsgpu_pairs_compute", "hsgpu_pairs_fetch", "hsgpu_pairs_info",
    "hsgpu_pairs_destroy", "hsgpu_graph_create", "hsgpu_graph_create_ex", "hsgpu_graph_build", "hsgpu_graph_adjacency", "hsgpu_graph_whispers",
    "hsgpu_graph_destroy", "hsgpu_edlib_align_batch", "hsgpu_edlibAlign", "hsgpu_edlibFreeAlignResult", "hsgpu_clip_reads",
]
PAIRS_DENSE, PAIRS_KEEP_ORDER, PAIRS_SIMT = 1, 2, 4
CLIP_DTYPE = np.dtype([(k, np.int32) for k in ("status", "read_start", "read_end", "cigar_start", "cigar_end", "op_first",
                                                "op_first_skip", "op_last", "op_last_take")])


def clipped_cigar(ops, clip) -> str:
    """the clipped CIGAR string of one hsgpu_clip result (what convert_cigar2 gives the reference, create_new_contigs.cpp:461)"""
    runs = []
    for k in range(int(clip["op_first"]), min(int(clip["op_last"]) + 1, len(ops))):
        n, ty = int(ops[k]) >> 4, int(ops[k]) & 15
        if k == int(clip["op_last"]):
            n = int(clip["op_last_take"])
        if k == int(clip["op_first"]):
            n -= int(clip["op_first_skip"])
        if n <= 0:
            continue
        if runs and runs[-1][1] == ty:
            runs[-1][0] += n
        else:
            runs.append([n, ty])
    return "".join(f"{n}{'MIDNSHP=X'[ty]}" for n, ty in runs)


def load():
    """Loads libhsgpu.so (building it is __graft_entry__.build()'s job) and declares prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HsgpuError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(make -C hairsplitter_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.hsgpu_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.hsgpu_ctx_destroy.argtypes = [vp]
    L.hsgpu_ctx_destroy.restype = None
    L.hsgpu_last_error.argtypes = [vp]
    L.hsgpu_last_error.restype = C.c_char_p
    L.hsgpu_sync.argtypes = [vp]
    L.hsgpu_launch_count.argtypes = [vp]
    L.hsgpu_launch_count.restype = i64
    L.hsgpu_stream.argtypes = [vp]
    L.hsgpu_stream.restype = vp
    L.hsgpu_profile_enable.argtypes = [vp, C.c_int]
    L.hsgpu_profile_report.argtypes = [vp]
    L.hsgpu_profile_report.restype = C.c_char_p
    L.hsgpu_host_alloc.argtypes = [C.POINTER(vp), i64]
    L.hsgpu_host_free.argtypes = [vp]
    L.hsgpu_host_free.restype = None
    L.hsgpu_pack_bases_ascii.argtypes = [C.c_char_p, i64, vp]
    L.hsgpu_pack_bases_ascii.restype = None
    L.hsgpu_pack_bases_codes.argtypes = [vp, i64, vp]
    L.hsgpu_pack_bases_codes.restype = None
    L.hsgpu_parse_cigar.argtypes = [C.c_char_p, vp, i64]
    L.hsgpu_parse_cigar.restype = i64
    L.hsgpu_pack_cigar8.argtypes = [vp, i64, vp, i64]
    L.hsgpu_pack_cigar8.restype = i64
    L.hsgpu_suspects_all.argtypes = [vp, i64, vp, vp, vp, vp]
    L.hsgpu_pileup_create.argtypes = [vp, C.POINTER(PileupInput), C.POINTER(vp)]
    L.hsgpu_pileup_destroy.argtypes = [vp]
    L.hsgpu_pileup_destroy.restype = None
    L.hsgpu_pileup_build.argtypes = [vp]
    L.hsgpu_pileup_stats.argtypes = [vp, vp, vp, vp]
    L.hsgpu_mean_distance.argtypes = [i64, i64]
    L.hsgpu_mean_distance.restype = f32
    L.hsgpu_pileup_read_ends.argtypes = [vp, vp]
    L.hsgpu_pileup_export.argtypes = [vp, i32, i64, vp, vp, vp]
    L.hsgpu_pileup_extract_columns.argtypes = [vp, i32, i32, vp, i64, vp, vp, vp]
    L.hsgpu_column_rank.argtypes = [vp, vp, f32]
    L.hsgpu_column_counts.argtypes = [vp, vp, vp]
    L.hsgpu_suspects.argtypes = [vp, i32, i32, vp, vp]
    L.hsgpu_column_summary.argtypes = [vp, i32, vp, vp, vp, vp]
    L.hsgpu_partition_tables.argtypes = [vp, i32, C.POINTER(Partitions), i32, vp, vp]
    L.hsgpu_robust_filter.argtypes = [vp, i32, C.POINTER(Partitions), i32, vp, i32, vp, vp]
    L.hsgpu_pileup_info.argtypes = [vp, vp]
    L.hsgpu_partitions_set.argtypes = [vp, vp]
    L.hsgpu_robust_filter_all.argtypes = [vp, i64, vp, vp]
    L.hsgpu_read_pair_counts.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.hsgpu_pairs_create.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(vp)]
    L.hsgpu_pairs_compute.argtypes = [vp]
    L.hsgpu_pairs_fetch.argtypes = [vp, i32, vp, vp]
    L.hsgpu_pairs_info.argtypes = [vp, vp]
    L.hsgpu_pairs_destroy.argtypes = [vp]
    L.hsgpu_pairs_destroy.restype = None
    L.hsgpu_graph_create.argtypes = [vp, i32, vp, vp, vp, f32, C.POINTER(vp)]
    L.hsgpu_graph_create_ex.argtypes = [vp, i32, vp, vp, vp, vp, f32, C.POINTER(vp)]
    L.hsgpu_graph_build.argtypes = [vp, vp]
    L.hsgpu_graph_adjacency.argtypes = [vp, vp, i64, vp, vp]
    L.hsgpu_graph_whispers.argtypes = [vp, i64, vp, vp, i32, vp, vp]
    L.hsgpu_graph_destroy.argtypes = [vp]
    L.hsgpu_graph_destroy.restype = None
    L.hsgpu_edlibAlign.argtypes = [vp, C.c_char_p, C.c_int, C.c_char_p, C.c_int, EdlibAlignConfig]
    L.hsgpu_edlibAlign.restype = EdlibAlignResult
    L.hsgpu_edlibFreeAlignResult.argtypes = [EdlibAlignResult]
    L.hsgpu_edlibFreeAlignResult.restype = None
    L.hsgpu_edlib_align_batch.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i64, vp, i64]
    for name in ("hsgpu_debug_rank_column", "hsgpu_debug_rh_order", "hsgpu_debug_sort_desc"):
        getattr(L, name).restype = C.c_int if name == "hsgpu_debug_rh_order" else None
    L.hsgpu_debug_rank_column.argtypes = [vp, C.c_int, vp, C.c_int]
    L.hsgpu_debug_rh_order.argtypes = [vp, C.c_int, vp]
    L.hsgpu_debug_sort_desc.argtypes = [vp, vp, C.c_int]
    _lib = L
    return L


def _a(x, dt):
    return np.ascontiguousarray(x, dtype=dt)


def pack_codes(codes: np.ndarray) -> np.ndarray:
    """u8 base codes -> 2-bit packed u32 words (vectorised equivalent of hsgpu_pack_bases_codes)."""
    codes = _a(codes, np.uint8)
    n = codes.shape[0]
    nw = (n + 15) // 16
    pad = np.zeros(nw * 16, dtype=np.uint32)
    pad[:n] = codes & 3
    sh = (2 * np.arange(16, dtype=np.uint32))[None, :]
    return np.bitwise_or.reduce(pad.reshape(nw, 16) << sh, axis=1).astype(np.uint32)


def compact_cigar(cigar: np.ndarray, cigar_off: np.ndarray):
    """u32 BAM ops -> the optional u16 form of hsgpu_pileup_input.cigar16 (ops longer than 4095 are split)."""
    ln = (cigar >> 4).astype(np.int64)
    op = (cigar & 15).astype(np.uint16)
    reps = np.maximum((ln + 4094) // 4095, 1)
    if int(reps.max(initial=1)) == 1:
        return ((ln.astype(np.uint16) << 4) | op), cigar_off
    first = np.cumsum(reps) - reps                      # index of the first piece of every op
    total = int(reps.sum())
    piece_len = np.full(total, 4095, dtype=np.int64)
    last = first + reps - 1
    piece_len[last] = ln - 4095 * (reps - 1)
    out = (piece_len.astype(np.uint16) << 4) | np.repeat(op, reps)
    new_off = np.concatenate([first, [total]])[cigar_off]
    return out, new_off.astype(np.int64)


def cigar8(cigar: np.ndarray, cigar_off: np.ndarray):
    """u32 BAM ops -> the 8-bit form of hsgpu_pileup_input.cigar8 through hsgpu_pack_cigar8, read by read
    (ops longer than 63 are split, so the offsets change). Raises when an op has no 8-bit form (N, P)."""
    L = load()
    cigar = np.ascontiguousarray(cigar, np.uint32)
    n = int(cigar_off[-1])
    total = int(L.hsgpu_pack_cigar8(cigar.ctypes.data, n, None, 0))
    if total < 0:
        raise HsgpuError("hsgpu_pack_cigar8: the CIGAR has ops without an 8-bit form")
    out = np.zeros(max(total, 1), np.uint8)
    new_off = np.zeros(cigar_off.size, np.int64)
    if total == n:  # nothing was split: one call, same offsets
        L.hsgpu_pack_cigar8(cigar.ctypes.data, n, out.ctypes.data, total)
        new_off[:] = cigar_off
        return out, new_off
    at = 0
    for r in range(cigar_off.size - 1):
        a, b = int(cigar_off[r]), int(cigar_off[r + 1])
        got = int(L.hsgpu_pack_cigar8(cigar.ctypes.data + 4 * a, b - a, out.ctypes.data + at, total - at))
        assert got >= 0
        at += got
        new_off[r + 1] = at
    assert at == total
    return out, new_off


class PackedBatch:
    """Host-side packing of a list of synth.ContigBatch into the flat arrays of hsgpu_pileup_input."""

    def __init__(self, chunks):
        L = load()
        nc = len(chunks)
        self.n_contigs = nc
        self.contig_len = np.array([c.length for c in chunks], dtype=np.int32)
        cw = [(c.length + 15) // 16 for c in chunks]
        self.contig_word_off = np.zeros(nc + 1, dtype=np.int64)
        np.cumsum(cw, out=self.contig_word_off[1:])
        self.contig_bases = np.zeros(max(1, int(self.contig_word_off[-1])), dtype=np.uint32)
        nreads = [c.n_reads for c in chunks]
        self.contig_read_off = np.zeros(nc + 1, dtype=np.int64)
        np.cumsum(nreads, out=self.contig_read_off[1:])
        nr = int(self.contig_read_off[-1])
        self.n_reads = nr
        self.read_len = np.concatenate([c.read_len() for c in chunks]).astype(np.int32) if nr else np.zeros(0, np.int32)
        rw = (self.read_len.astype(np.int64) + 15) // 16
        self.read_word_off = np.zeros(nr + 1, dtype=np.int64)
        np.cumsum(rw, out=self.read_word_off[1:])
        self.read_bases = np.zeros(max(1, int(self.read_word_off[-1])), dtype=np.uint32)
        self.read_start = np.concatenate([c.start for c in chunks]).astype(np.int32) if nr else np.zeros(0, np.int32)
        self.read_strand = np.concatenate([c.strand for c in chunks]).astype(np.uint8) if nr else np.zeros(0, np.uint8)
        self.cigar = np.concatenate([c.cigar for c in chunks]).astype(np.uint32) if nr else np.zeros(1, np.uint32)
        cn = np.concatenate([np.diff(c.cigar_off) for c in chunks]) if nr else np.zeros(0, np.int64)
        self.cigar_off = np.zeros(nr + 1, dtype=np.int64)
        np.cumsum(cn, out=self.cigar_off[1:])
        r = 0
        for ci, c in enumerate(chunks):
            w0 = int(self.contig_word_off[ci])
            contig = _a(c.contig, np.uint8)
            L.hsgpu_pack_bases_codes(contig.ctypes.data, c.length, self.contig_bases[w0:].ctypes.data)
            rb = _a(c.read_bases, np.uint8)
            # pack read by read so that every read starts on a word boundary
            ro = c.read_off
            for i in range(c.n_reads):
                L.hsgpu_pack_bases_codes(rb[ro[i]:].ctypes.data, int(ro[i + 1] - ro[i]),
                                         self.read_bases[int(self.read_word_off[r]):].ctypes.data)
                r += 1
        self.input_bytes = sum(int(a.nbytes) for a in (self.contig_len, self.contig_bases, self.contig_word_off,
                                                       self.contig_read_off, self.read_bases, self.read_word_off,
                                                       self.read_len, self.cigar, self.cigar_off, self.read_start,
                                                       self.read_strand))

    def use_compact_cigar(self):
        """switches the batch to the 16-bit CIGAR form (half the CIGAR bytes over PCIe)"""
        self.cigar16, self.cigar16_off = compact_cigar(self.cigar, self.cigar_off)
        self.input_bytes += int(self.cigar16.nbytes) - int(self.cigar.nbytes)
        return self

    def use_cigar8(self):
        """switches the batch to the 8-bit CIGAR form (a quarter of the CIGAR bytes over PCIe)"""
        before = int(self.cigar16.nbytes) if getattr(self, "cigar16", None) is not None else int(self.cigar.nbytes)
        self.cigar8, self.cigar8_off = cigar8(self.cigar, self.cigar_off)
        self.cigar16 = None
        self.input_bytes += int(self.cigar8.nbytes) - before
        return self

    def struct(self) -> PileupInput:
        s = PileupInput()
        s.n_contigs = self.n_contigs
        s.contig_len = self.contig_len.ctypes.data
        s.contig_bases = self.contig_bases.ctypes.data
        s.contig_word_off = self.contig_word_off.ctypes.data
        s.contig_read_off = self.contig_read_off.ctypes.data
        s.n_reads = self.n_reads
        s.read_bases = self.read_bases.ctypes.data
        s.read_word_off = self.read_word_off.ctypes.data
        s.read_len = self.read_len.ctypes.data
        s.cigar8 = None
        if getattr(self, "cigar8", None) is not None:
            s.cigar = None
            s.cigar16 = None
            s.cigar8 = self.cigar8.ctypes.data
            s.cigar_off = self.cigar8_off.ctypes.data
        elif getattr(self, "cigar16", None) is not None:
            s.cigar = None
            s.cigar16 = self.cigar16.ctypes.data
            s.cigar_off = self.cigar16_off.ctypes.data
        else:
            s.cigar = self.cigar.ctypes.data
            s.cigar16 = None
            s.cigar_off = self.cigar_off.ctypes.data
        s.read_start = self.read_start.ctypes.data
        s.read_strand = self.read_strand.ctypes.data
        return s


def make_partitions(parts):
    """parts: list of dicts(read_idx, state, more, less) -> (Partitions struct, keepalive arrays)."""
    n = len(parts)
    off = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum([len(p["read_idx"]) for p in parts], out=off[1:])
    cat = lambda k, dt: (np.concatenate([_a(p[k], dt) for p in parts]) if n and off[-1] else np.zeros(1, dt))
    idx, st, mo, le = cat("read_idx", np.int32), cat("state", np.int16), cat("more", np.int32), cat("less", np.int32)
    s = Partitions()
    s.n_parts = n
    s.part_off = off.ctypes.data
    s.read_idx = idx.ctypes.data
    s.state = st.ctypes.data
    s.more = mo.ctypes.data
    s.less = le.ctypes.data
    return s, (off, idx, st, mo, le)


class HostLogic:
    """libhshost.so: the sequential host side of keep_only_robust_variants (loops 1-2: greedy partition building
    and merging, reference src/call_variants.cpp:590-708 + src/Partition.cpp), the C++ code HS_call_variants runs
    between hsgpu_column_rank and hsgpu_robust_filter_all. Product code (hairsplitter_b200/host), not the oracle."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(_HERE, "libhshost.so"))
            L.hshost_build_partitions.restype = C.c_void_p
            L.hshost_build_partitions.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_float]
            L.hshost_parts_count.argtypes = [C.c_void_p]
            L.hshost_part_size.argtypes = [C.c_void_p, C.c_int]
            L.hshost_part_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
            L.hshost_parts_free.argtypes = [C.c_void_p]
            cls._lib = L
        return cls._lib

    @classmethod
    def read_graph_low_memory(cls, col, masked, error_rate):
        """create_read_graph_low_memory of the host path (hs_sepreads.cpp) -> (adj_off, adj, reads cover consecutive SNPs)"""
        L = cls.lib()
        n_reads, snp_off, idx, code, rb, sb = col
        snp_off, idx, code = _a(snp_off, np.int64), _a(idx, np.uint32), _a(code, np.uint8)
        rb, sb, masked = _a(rb, np.uint8), _a(sb, np.uint8), _a(masked, np.int32)
        f = L.hshost_read_graph_low_memory
        f.restype = C.c_int64
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        adj_off = np.zeros(masked.size + 1, np.int64)
        cons = C.c_int(0)
        n = f(n_reads, snp_off.size - 1, snp_off.ctypes.data, idx.ctypes.data, code.ctypes.data, rb.ctypes.data, sb.ctypes.data,
              masked.size, masked.ctypes.data, float(error_rate), adj_off.ctypes.data, None, C.byref(cons))
        adj = np.zeros(max(int(n), 1), np.int32)
        f(n_reads, snp_off.size - 1, snp_off.ctypes.data, idx.ctypes.data, code.ctypes.data, rb.ctypes.data, sb.ctypes.data,
          masked.size, masked.ctypes.data, float(error_rate), adj_off.ctypes.data, adj.ctypes.data, C.byref(cons))
        return adj_off, adj[:int(n)], bool(cons.value)

    @classmethod
    def build_partitions(cls, pos, off, read_idx, code, ref_base, second_base, mean_error):
        """suspect columns (CSR over columns) -> the final partitions as a list of dicts"""
        L = cls.lib()
        pos, off = _a(pos, np.int32), _a(off, np.int64)
        read_idx, code = _a(read_idx, np.uint32), _a(code, np.uint8)
        rb, sb = _a(ref_base, np.uint8), _a(second_base, np.uint8)
        h = L.hshost_build_partitions(int(pos.size), off.ctypes.data, read_idx.ctypes.data, code.ctypes.data,
                                      pos.ctypes.data, rb.ctypes.data, sb.ctypes.data, float(mean_error))
        parts = []
        for p in range(L.hshost_parts_count(h)):
            n = L.hshost_part_size(h, p)
            a = [np.zeros(n, np.int32), np.zeros(n, np.int16), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(2, np.int32)]
            L.hshost_part_get(h, p, *[x.ctypes.data for x in a])
            parts.append(dict(read_idx=a[0], state=a[1], more=a[2], less=a[3], left=int(a[4][0]), right=int(a[4][1])))
        L.hshost_parts_free(h)
        return parts


class Context:
    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.hsgpu_ctx_create(device, C.byref(h))
        if rc != 0:
            raise HsgpuError(f"hsgpu_ctx_create failed ({rc}): {self.lib.hsgpu_last_error(None).decode()}")
        self.h = h

    def check(self, rc, what=""):
        if rc != 0:
            raise HsgpuError(f"{what} failed ({rc}): {self.lib.hsgpu_last_error(self.h).decode()}")

    def sync(self):
        self.check(self.lib.hsgpu_sync(self.h), "hsgpu_sync")

    def launches(self) -> int:
        return int(self.lib.hsgpu_launch_count(self.h))

    def stream(self) -> int:
        return int(self.lib.hsgpu_stream(self.h) or 0)

    def profile(self, on: bool):
        self.check(self.lib.hsgpu_profile_enable(self.h, 1 if on else 0), "hsgpu_profile_enable")

    def profile_report(self):
        """{kernel name: (launches, total_ms)} since the last report"""
        out = {}
        for line in self.lib.hsgpu_profile_report(self.h).decode().splitlines():
            name, n, ms = line.split("\t")
            out[name] = (int(n), float(ms))
        return out

    def close(self):
        if self.h:
            self.lib.hsgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- read x read counts ------------------------------------------------------------------
    def read_pair_counts(self, n_reads, snp_off, read_idx, code, ref_base, second_base):
        snp_off = _a(snp_off, np.int64)
        read_idx = _a(read_idx, np.uint32)
        code = _a(code, np.uint8)
        ref_base = _a(ref_base, np.uint8)
        second_base = _a(second_base, np.uint8)
        sim = np.zeros((n_reads, n_reads), dtype=np.int32)
        diff = np.zeros((n_reads, n_reads), dtype=np.int32)
        self.check(self.lib.hsgpu_read_pair_counts(self.h, n_reads, snp_off.shape[0] - 1, snp_off.ctypes.data,
                                                   read_idx.ctypes.data, code.ctypes.data, ref_base.ctypes.data,
                                                   second_base.ctypes.data, sim.ctypes.data, diff.ctypes.data),
                   "hsgpu_read_pair_counts")
        return sim, diff

    # -- read clipping (modify_GFA) ------------------------------------------------------------
    def clip_reads(self, cigar, cigar_off, pos_2_1, item_read, left, right):
        """hsgpu_clip_reads -> structured array (status, read_start, read_end, cigar_start, cigar_end, op_first,
        op_first_skip, op_last, op_last_take) per item"""
        cigar, cigar_off = _a(cigar, np.uint32), _a(cigar_off, np.int64)
        pos_2_1, item_read = _a(pos_2_1, np.int32), _a(item_read, np.int64)
        left, right = _a(left, np.int32), _a(right, np.int32)
        out = np.zeros(item_read.size, dtype=CLIP_DTYPE)
        f = self.lib.hsgpu_clip_reads
        f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.check(f(self.h, pos_2_1.size, cigar.ctypes.data, cigar_off.ctypes.data, pos_2_1.ctypes.data, item_read.size,
                     item_read.ctypes.data, left.ctypes.data, right.ctypes.data, out.ctypes.data), "hsgpu_clip_reads")
        return out

    # -- edlib ---------------------------------------------------------------------------------
    def edlib_align_batch(self, queries, targets, k=-1, mode=2, task=2):
        """queries/targets: lists of bytes. Returns (results structured array, ends, starts, alignment)."""
        n = len(queries)
        qo = np.zeros(n + 1, dtype=np.int64)
        to = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(q) for q in queries], out=qo[1:])
        np.cumsum([len(t) for t in targets], out=to[1:])
        qb = np.frombuffer(b"".join(queries) or b"\0", dtype=np.uint8)
        tb = np.frombuffer(b"".join(targets) or b"\0", dtype=np.uint8)
        return self.edlib_align_batch_arrays(qb, qo, tb, to, k=k, mode=mode, task=task)

    def edlib_align_batch_arrays(self, qb, qo, tb, to, k=-1, mode=2, task=2, out=None):
        """hsgpu_edlib_align_batch on concatenated sequences (uint8) and their int64 offsets -- the C ABI's own
        arguments; `out` = (results, ends, starts, alignment) buffers to fill (numpy arrays, e.g. views of pinned
        memory), allocated here when None."""
        n = len(qo) - 1
        if out is None:
            res = np.zeros(n, dtype=EDLIB_RESULT_DTYPE)
            ends = np.zeros(int(to[-1]) + n + 8, dtype=np.int32)
            starts = np.zeros(int(to[-1]) + n + 8, dtype=np.int32)
            aln = np.zeros(int(qo[-1] + to[-1]) + 8, dtype=np.uint8)
        else:
            res, ends, starts, aln = out
        self.check(self.lib.hsgpu_edlib_align_batch(self.h, n, qb.ctypes.data, qo.ctypes.data, tb.ctypes.data,
                                                    to.ctypes.data, k, mode, task, res.ctypes.data, ends.ctypes.data,
                                                    starts.ctypes.data, len(ends), aln.ctypes.data, len(aln)),
                   "hsgpu_edlib_align_batch")
        return res, ends, starts, aln


    def edlib_align(self, query: bytes, target: bytes, k=-1, mode=2, task=2):
        """hsgpu_edlibAlign: one pair, edlib's own result struct; returned as the dict the oracle uses"""
        cfg = EdlibAlignConfig(k, mode, task, None, 0)
        r = self.lib.hsgpu_edlibAlign(self.h, query, len(query), target, len(target), cfg)
        n = r.numLocations
        out = dict(status=r.status, edit_distance=r.editDistance, alphabet_length=r.alphabetLength,
                   end_locations=(np.array(r.endLocations[:n], np.int32) if r.endLocations else None),
                   start_locations=(np.array(r.startLocations[:n], np.int32) if r.startLocations else None),
                   alignment=(np.array(r.alignment[:r.alignmentLength], np.uint8) if r.alignment else None))
        self.lib.hsgpu_edlibFreeAlignResult(r)
        return out


class Pairs:
    """hsgpu_pairs: read x read SNP agreement counts of a batch of contigs, operands and results on the device.
    contigs: list of (n_reads, snp_off, read_idx, code, ref_base, second_base) with contig-local read indices."""

    def __init__(self, ctx: Context, contigs, flags=0):
        self.ctx, self.lib = ctx, ctx.lib
        self.n_reads = _a([c[0] for c in contigs], np.int32)
        n_snps = [int(np.asarray(c[1]).shape[0]) - 1 for c in contigs]
        self.snp_base = np.zeros(len(contigs) + 1, np.int64)
        self.snp_base[1:] = np.cumsum(n_snps)
        offs, cell0 = [np.zeros(1, np.int64)], 0
        for c in contigs:
            so = _a(c[1], np.int64)
            offs.append(so[1:] - so[0] + cell0)
            cell0 += int(so[-1] - so[0])
        self.snp_off = np.concatenate(offs)
        cat = lambda k, dt: (np.concatenate([_a(c[k], dt) for c in contigs]) if contigs else np.zeros(0, dt))
        self.read_idx, self.code = cat(2, np.uint32), cat(3, np.uint8)
        self.ref_base, self.second_base = cat(4, np.uint8), cat(5, np.uint8)
        h = C.c_void_p()
        ctx.check(self.lib.hsgpu_pairs_create(ctx.h, len(contigs), self.n_reads.ctypes.data, self.snp_base.ctypes.data,
                                              self.snp_off.ctypes.data, self.read_idx.ctypes.data, self.code.ctypes.data,
                                              self.ref_base.ctypes.data, self.second_base.ctypes.data, flags, C.byref(h)),
                  "hsgpu_pairs_create")
        self.h = h

    def compute(self):
        self.ctx.check(self.lib.hsgpu_pairs_compute(self.h), "hsgpu_pairs_compute")

    def fetch(self, contig):
        n = int(self.n_reads[contig])
        sim, diff = np.zeros((n, n), np.int32), np.zeros((n, n), np.int32)
        self.ctx.check(self.lib.hsgpu_pairs_fetch(self.h, contig, sim.ctypes.data, diff.ctypes.data), "hsgpu_pairs_fetch")
        return sim, diff

    def info(self):
        v = np.zeros(8, np.int64)
        self.ctx.check(self.lib.hsgpu_pairs_info(self.h, v.ctypes.data), "hsgpu_pairs_info")
        return dict(zip(("tile_pairs", "tile_pairs_dense", "kblocks", "kblocks_dense", "rows", "k_ld", "out_elems",
                         "identity"), (int(x) for x in v)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.hsgpu_pairs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Graph:
    """hsgpu_graph: read graphs (create_read_graph_matrix) and chinese-whispers runs of a batch of windows over the
    device-resident counts of a Pairs object. windows: list of (contig, ascending read indices spanning the window)."""

    def __init__(self, pairs: Pairs, windows, error_rate, low_memory=None):
        """low_memory: optional per-window flags -- the neighbour choice of create_read_graph_low_memory"""
        self.ctx, self.lib, self.pairs = pairs.ctx, pairs.lib, pairs
        self.win_contig = _a([w[0] for w in windows], np.int32)
        self.win_off = np.zeros(len(windows) + 1, np.int64)
        self.win_off[1:] = np.cumsum([len(w[1]) for w in windows])
        self.win_reads = (np.concatenate([_a(w[1], np.int32) for w in windows]) if windows else np.zeros(0, np.int32))
        h = C.c_void_p()
        self.win_low = None if low_memory is None else _a(low_memory, np.uint8)
        self.ctx.check(self.lib.hsgpu_graph_create_ex(pairs.h, len(windows), self.win_contig.ctypes.data, self.win_off.ctypes.data,
                                                      self.win_reads.ctypes.data,
                                                      None if self.win_low is None else self.win_low.ctypes.data,
                                                      float(error_rate), C.byref(h)), "hsgpu_graph_create_ex")
        self.h = h
        self.replayed = 0

    def build(self):
        n = C.c_int64(0)
        self.ctx.check(self.lib.hsgpu_graph_build(self.h, C.byref(n)), "hsgpu_graph_build")
        self.replayed = int(n.value)

    def adjacency(self):
        """(adj_off, adj): CSR over all masked reads of all windows, neighbours as local indices"""
        total = int(self.win_off[-1])
        adj_off = np.zeros(total + 1, np.int64)
        n = C.c_int64(0)
        self.ctx.check(self.lib.hsgpu_graph_adjacency(self.h, adj_off.ctypes.data, 0, None, C.byref(n)), "hsgpu_graph_adjacency")
        adj = np.zeros(max(int(n.value), 1), np.int32)
        self.ctx.check(self.lib.hsgpu_graph_adjacency(self.h, None, int(n.value), adj.ctypes.data, C.byref(n)),
                       "hsgpu_graph_adjacency")
        return adj_off, adj[:int(n.value)]

    def whispers(self, run_window, init_labels, order_rank, n_orders=1):
        """run_window[i] = window of run i; init_labels = the runs' local label vectors concatenated;
        order_rank = per contig n_orders arrays of n_reads positions, concatenated"""
        run_window = _a(run_window, np.int32)
        init = _a(init_labels, np.int32)
        rank = _a(order_rank, np.int32)
        out = np.zeros(init.size, np.int32)
        self.ctx.check(self.lib.hsgpu_graph_whispers(self.h, run_window.size, run_window.ctypes.data, init.ctypes.data,
                                                     int(n_orders), rank.ctypes.data, out.ctypes.data), "hsgpu_graph_whispers")
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.hsgpu_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pileup:
    """hsgpu_pileup: a batch of contig chunks resident on the device."""

    def __init__(self, ctx: Context, packed: PackedBatch):
        self.ctx = ctx
        self.lib = ctx.lib
        self.packed = packed
        h = C.c_void_p()
        s = packed.struct()
        ctx.check(self.lib.hsgpu_pileup_create(ctx.h, C.byref(s), C.byref(h)), "hsgpu_pileup_create")
        self.h = h

    def close(self):
        if self.h:
            self.lib.hsgpu_pileup_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self):
        self.ctx.check(self.lib.hsgpu_pileup_build(self.h), "hsgpu_pileup_build")

    def stats(self):
        nc = self.packed.n_contigs
        cells = np.zeros(nc, np.int64)
        dist = np.zeros(nc, np.int64)
        alen = np.zeros(nc, np.int64)
        self.ctx.check(self.lib.hsgpu_pileup_stats(self.h, cells.ctypes.data, dist.ctypes.data, alen.ctypes.data),
                       "hsgpu_pileup_stats")
        return cells, dist, alen

    def mean_distance(self, dist, alen) -> float:
        return float(np.float32(self.lib.hsgpu_mean_distance(int(dist), int(alen))))

    def read_ends(self):
        out = np.zeros(self.packed.n_reads, np.int32)
        self.ctx.check(self.lib.hsgpu_pileup_read_ends(self.h, out.ctypes.data), "hsgpu_pileup_read_ends")
        return out

    def column_rank(self, mean_error=None, auto_threshold=0.33):
        me = None
        if mean_error is not None:
            me = _a(mean_error, np.float32)
        self.ctx.check(self.lib.hsgpu_column_rank(self.h, me.ctypes.data if me is not None else None,
                                                  float(auto_threshold)), "hsgpu_column_rank")

    def column_counts(self):
        nc = self.packed.n_contigs
        ns = np.zeros(nc, np.int32)
        ds = np.zeros(nc, np.int64)
        self.ctx.check(self.lib.hsgpu_column_counts(self.h, ns.ctypes.data, ds.ctypes.data), "hsgpu_column_counts")
        return ns, ds

    def suspects(self, contig):
        cap = int(self.packed.contig_len[contig]) // 6 + 2
        pos = np.zeros(cap, np.int32)
        au = np.zeros(cap, np.uint8)
        ns, _ = self.column_counts()
        self.ctx.check(self.lib.hsgpu_suspects(self.h, contig, cap, pos.ctypes.data, au.ctypes.data), "hsgpu_suspects")
        n = int(ns[contig])
        return pos[:n], au[:n]

    def suspects_all(self, want_depth=True, out=None):
        """every contig's suspect list in one call: (pos, is_automatic, off[n_contigs+1], depth_sum); `out` = buffers
        of a caller that keeps them from call to call (pos int32, is_automatic uint8, off int64, depth_sum int64)"""
        nc = self.packed.n_contigs
        if out is not None:
            pos, au, off, ds = out
            cap = min(len(pos), len(au))
        else:
            cap = int(self.packed.contig_len.astype(np.int64).sum()) // 6 + 2 * nc
            pos = np.empty(cap, np.int32)
            au = np.empty(cap, np.uint8)
            off = np.zeros(nc + 1, np.int64)
            ds = np.zeros(nc, np.int64)
        self.ctx.check(self.lib.hsgpu_suspects_all(self.h, cap, pos.ctypes.data, au.ctypes.data, off.ctypes.data,
                                                   ds.ctypes.data if want_depth else None), "hsgpu_suspects_all")
        n = int(off[nc])
        return pos[:n], au[:n], off, ds

    def column_summary(self, contig):
        Lc = int(self.packed.contig_len[contig])
        rb = np.zeros(Lc, np.uint8)
        sb = np.zeros(Lc, np.uint8)
        cnt = np.zeros(3 * Lc, np.uint32)
        dep = np.zeros(Lc, np.uint32)
        self.ctx.check(self.lib.hsgpu_column_summary(self.h, contig, rb.ctypes.data, sb.ctypes.data, cnt.ctypes.data,
                                                     dep.ctypes.data), "hsgpu_column_summary")
        return dict(ref_base=rb, second_base=sb, counts=cnt.reshape(Lc, 3), depth=dep)

    def export(self, contig):
        Lc = int(self.packed.contig_len[contig])
        col_off = np.zeros(Lc + 1, np.int64)
        rc = self.lib.hsgpu_pileup_export(self.h, contig, 0, col_off.ctypes.data, None, None)
        if rc not in (0, -4):
            self.ctx.check(rc, "hsgpu_pileup_export")
        n = int(col_off[-1])
        idx = np.zeros(max(n, 1), np.uint32)
        code = np.zeros(max(n, 1), np.uint8)
        self.ctx.check(self.lib.hsgpu_pileup_export(self.h, contig, n, col_off.ctypes.data, idx.ctypes.data,
                                                    code.ctypes.data), "hsgpu_pileup_export")
        return dict(col_off=col_off, read_idx=idx[:n], code=code[:n])

    def extract_columns(self, contig, pos):
        pos = _a(pos, np.int32)
        off = np.zeros(pos.shape[0] + 1, np.int64)
        rc = self.lib.hsgpu_pileup_extract_columns(self.h, contig, pos.shape[0], pos.ctypes.data, 0, off.ctypes.data,
                                                   None, None)
        if rc not in (0, -4):
            self.ctx.check(rc, "hsgpu_pileup_extract_columns")
        n = int(off[-1])
        idx = np.zeros(max(n, 1), np.uint32)
        code = np.zeros(max(n, 1), np.uint8)
        self.ctx.check(self.lib.hsgpu_pileup_extract_columns(self.h, contig, pos.shape[0], pos.ctypes.data, n,
                                                             off.ctypes.data, idx.ctypes.data, code.ctypes.data),
                       "hsgpu_pileup_extract_columns")
        return off, idx[:n], code[:n]

    def partition_tables(self, contig, parts, pos):
        pos = _a(pos, np.int32)
        s, keep = make_partitions(parts)
        out = np.zeros((pos.shape[0], len(parts)), dtype=DISTANCE_DTYPE)
        self.ctx.check(self.lib.hsgpu_partition_tables(self.h, contig, C.byref(s), pos.shape[0], pos.ctypes.data,
                                                       out.ctypes.data), "hsgpu_partition_tables")
        del keep
        return out

    def info(self):
        v = np.zeros(8, np.int64)
        self.ctx.check(self.lib.hsgpu_pileup_info(self.h, v.ctypes.data), "hsgpu_pileup_info")
        return v

    def cigar_bytes(self):
        return int(self.info()[0])

    def filter_info(self):
        v = self.info()
        return {"active_columns": int(v[4]), "active_cells": int(v[5]), "state_bytes": int(v[6]), "kept": int(v[7]),
                "parts_per_cell": float(v[6]) / max(float(v[5]), 1.0)}

    @staticmethod
    def prepare_partitions(parts_per_contig):
        """the hsgpu_partitions array of a batch (parts_per_contig[c] = list of partition dicts of contig c) as
        (ctypes array, keep-alive arrays, bytes the arrays hold): build once, hand to partitions_set as often as needed"""
        n = len(parts_per_contig)
        arr = (Partitions * n)()
        keep, nbytes = [], 0
        for c, parts in enumerate(parts_per_contig):
            s, k = make_partitions(parts)
            arr[c] = s
            keep.append(k)
            nbytes += sum(int(x.nbytes) for x in k) if len(parts) else 0
        return arr, keep, nbytes

    def partitions_set(self, parts_per_contig=None, prepared=None):
        """uploads the final partitions of every contig; they stay on the device with the pileup"""
        if prepared is None:
            prepared = self.prepare_partitions(parts_per_contig)
        assert len(prepared[0]) == int(self.packed.contig_len.shape[0])
        self.ctx.check(self.lib.hsgpu_partitions_set(self.h, C.cast(prepared[0], C.c_void_p)), "hsgpu_partitions_set")

    def robust_filter_all(self, capacity=None, out=None):
        """loops 3+4 of keep_only_robust_variants for every contig in one launch -> (kept positions, off[n_contigs+1]).
        Without a capacity the call is made twice (sizes first). `out` = (kept int32, off int64) buffers of a caller
        that keeps them from call to call."""
        n = int(self.packed.contig_len.shape[0])
        if out is not None:
            kept, off = out
            self.ctx.check(self.lib.hsgpu_robust_filter_all(self.h, len(kept), kept.ctypes.data, off.ctypes.data),
                           "hsgpu_robust_filter_all")
            return kept[: int(off[n])], off
        off = np.zeros(n + 1, np.int64)
        if capacity is None:
            rc = self.lib.hsgpu_robust_filter_all(self.h, 0, None, off.ctypes.data)
            if rc not in (0, -4):
                self.ctx.check(rc, "hsgpu_robust_filter_all")
            capacity = int(off[-1])
            if capacity == 0:
                return np.zeros(0, np.int32), off
        kept = np.zeros(max(int(capacity), 1), np.int32)
        self.ctx.check(self.lib.hsgpu_robust_filter_all(self.h, int(capacity), kept.ctypes.data, off.ctypes.data),
                       "hsgpu_robust_filter_all")
        return kept[: int(off[-1])], off

    def host_partitions(self, contig):
        """loops 1-2 of keep_only_robust_variants on the host (libhshost.so) for one contig of a ranked pileup: what
        HS_call_variants does between hsgpu_column_rank and hsgpu_robust_filter_all"""
        pos, _ = self.suspects(contig)
        off, idx, code = self.extract_columns(contig, pos)
        summ = self.column_summary(contig)
        _, dist, alen = self.stats()
        md = self.mean_distance(dist[contig], alen[contig])
        return HostLogic.build_partitions(pos, off, idx, code, summ["ref_base"][pos], summ["second_base"][pos], md)

    def robust_filter(self, contig, parts, suspect_pos):
        suspect_pos = _a(suspect_pos, np.int32)
        s, keep = make_partitions(parts)
        cap = int(self.packed.contig_len[contig])
        kept = np.zeros(cap + 1, np.int32)
        n = np.zeros(1, np.int32)
        self.ctx.check(self.lib.hsgpu_robust_filter(self.h, contig, C.byref(s), suspect_pos.shape[0],
                                                    suspect_pos.ctypes.data, cap, kept.ctypes.data, n.ctypes.data),
                       "hsgpu_robust_filter")
        del keep
        return kept[: int(n[0])]
