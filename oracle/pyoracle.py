"""ctypes bindings for the parity oracle (TEST INFRASTRUCTURE ONLY).

Two libraries are bound here:
  * ``oracle/liboracle.so``      -- our plain-C restatement of the reference's hot path (hs_oracle.c)
  * ``oracle/_ref/libhsref_*.so`` -- the UNMODIFIED reference compiled from /root/reference by
                                    oracle/Makefile (present wherever it was built; travels to the
                                    GPU box as a prebuilt file)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; nothing under hairsplitter_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_P = np.ctypeslib.ndpointer


def build(ref: bool = True) -> None:
    """Compile liboracle.so (and _ref/ when the reference sources are present)."""
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so", "sr_hostcheck"], check=True)
    if ref and os.path.exists("/root/reference/src/call_variants.cpp"):
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref"], check=True)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Oracle:
    """liboracle.so"""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.hso_pileup.restype = C.c_int64
        L.hso_pileup.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        L.hso_mean_distance.restype = C.c_float
        L.hso_mean_distance.argtypes = [C.c_int64, C.c_int64]
        L.hso_ref_codes.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.hso_rh_order.restype = C.c_int
        L.hso_rh_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hso_sort_desc.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.hso_column_rank.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hso_call_variants.restype = C.c_int32
        L.hso_call_variants.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.hso_distance.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_void_p]
        L.hso_chi_square.restype = C.c_float
        L.hso_chi_square.argtypes = [C.c_int32] * 4
        L.hso_rescue_prefilter.restype = C.c_int
        L.hso_rescue_prefilter.argtypes = [C.c_int32, C.c_int32]
        L.hso_robust_filter.restype = C.c_int32
        L.hso_robust_filter.argtypes = [C.c_int32] + [C.c_void_p] * 5 + [C.c_int32] + [C.c_void_p] * 5 + [C.c_int32,
                                                                                                       C.c_void_p,
                                                                                                       C.c_void_p]
        L.hso_read_pair_counts.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 7
        L.hso_edlib_align.restype = C.c_int32
        L.hso_edlib_align.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [
            C.c_void_p] * 7

    # -- pileup ---------------------------------------------------------------------------------
    def pileup(self, cb):
        """cb: synth.ContigBatch. Returns dict(col_off, read_idx, code, stats, read_end)."""
        Lc = cb.length
        contig = _arr(cb.contig, np.uint8)
        rb = _arr(cb.read_bases, np.uint8)
        ro = _arr(cb.read_off, np.int64)
        cg = _arr(cb.cigar, np.uint32)
        co = _arr(cb.cigar_off, np.int64)
        st = _arr(cb.start, np.int32)
        sd = _arr(cb.strand, np.uint8)
        col_off = np.zeros(Lc + 1, dtype=np.int64)
        stats = np.zeros(2, dtype=np.int64)
        read_end = np.zeros(cb.n_reads, dtype=np.int32)
        args = [contig.ctypes.data, Lc, cb.n_reads, rb.ctypes.data, ro.ctypes.data, cg.ctypes.data, co.ctypes.data,
                st.ctypes.data, sd.ctypes.data]
        n = self.lib.hso_pileup(*args, 0, col_off.ctypes.data, None, None, stats.ctypes.data, read_end.ctypes.data)
        read_idx = np.zeros(max(n, 1), dtype=np.uint32)
        code = np.zeros(max(n, 1), dtype=np.uint8)
        self.lib.hso_pileup(*args, n, col_off.ctypes.data, read_idx.ctypes.data, code.ctypes.data, stats.ctypes.data,
                            read_end.ctypes.data)
        return dict(col_off=col_off, read_idx=read_idx[:n], code=code[:n], stats=stats, read_end=read_end)

    def mean_distance(self, dist, alen):
        return float(np.float32(self.lib.hso_mean_distance(int(dist), int(alen))))

    def ref_codes(self, contig):
        contig = _arr(contig, np.uint8)
        out = np.zeros(contig.shape[0], dtype=np.uint8)
        self.lib.hso_ref_codes(contig.ctypes.data, contig.shape[0], out.ctypes.data)
        return out

    def rh_order(self, keys):
        keys = _arr(keys, np.uint8)
        out = np.zeros(keys.shape[0] + 8, dtype=np.uint8)
        n = self.lib.hso_rh_order(keys.ctypes.data, keys.shape[0], out.ctypes.data)
        return out[:n]

    def sort_desc(self, keys, counts):
        keys = _arr(keys, np.uint8).copy()
        counts = _arr(counts, np.int32).copy()
        self.lib.hso_sort_desc(keys.ctypes.data, counts.ctypes.data, keys.shape[0])
        return keys, counts

    def column_rank(self, codes):
        codes = _arr(codes, np.uint8)
        out = np.zeros(5, dtype=np.int32)
        self.lib.hso_column_rank(codes.ctypes.data, codes.shape[0], out.ctypes.data)
        return out

    def call_variants(self, col_off, code, mean_error, auto_threshold=0.33):
        col_off = _arr(col_off, np.int64)
        code = _arr(code, np.uint8)
        Lc = col_off.shape[0] - 1
        ref_base = np.zeros(Lc, dtype=np.uint8)
        second_base = np.zeros(Lc, dtype=np.uint8)
        cap = Lc // 6 + 8
        pos = np.zeros(cap, dtype=np.int32)
        is_auto = np.zeros(cap, dtype=np.uint8)
        depth = np.zeros(1, dtype=np.int64)
        n = self.lib.hso_call_variants(col_off.ctypes.data, code.ctypes.data, Lc, float(mean_error),
                                       float(auto_threshold), ref_base.ctypes.data, second_base.ctypes.data,
                                       pos.ctypes.data, is_auto.ctypes.data, cap, depth.ctypes.data)
        return dict(ref_base=ref_base, second_base=second_base, suspect_pos=pos[:n], suspect_is_auto=is_auto[:n],
                    depth_sum=int(depth[0]))

    def distance(self, p_idx, p_state, p_more, p_less, c_idx, c_code, ref_base):
        p_idx = _arr(p_idx, np.int32)
        p_state = _arr(p_state, np.int16)
        p_more = _arr(p_more, np.int32)
        p_less = _arr(p_less, np.int32)
        c_idx = _arr(c_idx, np.uint32)
        c_code = _arr(c_code, np.uint8)
        out = np.zeros(10, dtype=np.int32)
        self.lib.hso_distance(p_idx.shape[0], p_idx.ctypes.data, p_state.ctypes.data, p_more.ctypes.data,
                              p_less.ctypes.data, c_idx.shape[0], c_idx.ctypes.data, c_code.ctypes.data,
                              int(ref_base), out.ctypes.data)
        return out

    def chi_square(self, n00, n01, n10, n11):
        return float(np.float32(self.lib.hso_chi_square(int(n00), int(n01), int(n10), int(n11))))

    def rescue_prefilter(self, ref_base, second_base):
        return bool(self.lib.hso_rescue_prefilter(int(ref_base), int(second_base)))

    def robust_filter(self, col_off, read_idx, code, ref_base, second_base, parts, suspect_pos):
        col_off = _arr(col_off, np.int64)
        read_idx = _arr(read_idx, np.uint32)
        code = _arr(code, np.uint8)
        ref_base = _arr(ref_base, np.uint8)
        second_base = _arr(second_base, np.uint8)
        suspect_pos = _arr(suspect_pos, np.int32)
        n = len(parts)
        off = np.zeros(n + 1, dtype=np.int64)
        if n:
            np.cumsum([len(q["read_idx"]) for q in parts], out=off[1:])
        cat = lambda k, dt: (np.concatenate([_arr(q[k], dt) for q in parts]) if n and off[-1] else np.zeros(1, dt))
        idx, st, mo, le = cat("read_idx", np.int32), cat("state", np.int16), cat("more", np.int32), cat("less", np.int32)
        Lc = col_off.shape[0] - 1
        kept = np.zeros(Lc + 1, dtype=np.int32)
        nk = self.lib.hso_robust_filter(Lc, col_off.ctypes.data, read_idx.ctypes.data, code.ctypes.data,
                                        ref_base.ctypes.data, second_base.ctypes.data, n, off.ctypes.data,
                                        idx.ctypes.data, st.ctypes.data, mo.ctypes.data, le.ctypes.data,
                                        suspect_pos.shape[0], suspect_pos.ctypes.data, kept.ctypes.data)
        return kept[:nk]

    def read_pair_counts(self, n_reads, snp_off, read_idx, code, ref_base, second_base):
        snp_off = _arr(snp_off, np.int64)
        read_idx = _arr(read_idx, np.uint32)
        code = _arr(code, np.uint8)
        ref_base = _arr(ref_base, np.uint8)
        second_base = _arr(second_base, np.uint8)
        sim = np.zeros((n_reads, n_reads), dtype=np.int32)
        diff = np.zeros((n_reads, n_reads), dtype=np.int32)
        self.lib.hso_read_pair_counts(n_reads, snp_off.shape[0] - 1, snp_off.ctypes.data, read_idx.ctypes.data,
                                      code.ctypes.data, ref_base.ctypes.data, second_base.ctypes.data,
                                      sim.ctypes.data, diff.ctypes.data)
        return sim, diff


    def read_graph(self, sim, diff, masked, error_rate):
        """create_read_graph_matrix restated (oracle/hs_oracle_sr.cpp) on dense counts -> (adj_off, adj), local indices"""
        sim, diff = _arr(sim, np.int32), _arr(diff, np.int32)
        masked = _arr(masked, np.int32)
        n = sim.shape[0]
        adj_off = np.zeros(masked.size + 1, np.int64)
        f = self.lib.hso_read_graph
        f.restype = C.c_int64
        f.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        cnt = f(n, sim.ctypes.data, diff.ctypes.data, masked.size, masked.ctypes.data, float(error_rate), adj_off.ctypes.data, None)
        adj = np.zeros(max(int(cnt), 1), np.int32)
        f(n, sim.ctypes.data, diff.ctypes.data, masked.size, masked.ctypes.data, float(error_rate), adj_off.ctypes.data, adj.ctypes.data)
        return adj_off, adj[:int(cnt)]

    def shuffled_order(self, n, seed):
        """0..n-1 after std::shuffle with std::mt19937(seed)"""
        out = np.zeros(n, np.int32)
        f = self.lib.hso_shuffled_order
        f.restype = None
        f.argtypes = [C.c_int32, C.c_uint32, C.c_void_p]
        f(n, int(seed), out.ctypes.data)
        return out

    def chinese_whispers(self, n_reads, masked, adj_off, adj, init_local, seed):
        """chinese_whispers_high_memory restated, labels as local indices; sweep s is shuffled with seed (an int) or
        seeds[min(s, len-1)] (a sequence)"""
        masked = _arr(masked, np.int32)
        adj_off, adj, init = _arr(adj_off, np.int64), _arr(adj, np.int32), _arr(init_local, np.int32)
        out = np.zeros(masked.size, np.int32)
        f = self.lib.hso_chinese_whispers
        f.restype = None
        seeds = _arr([seed] if np.isscalar(seed) else seed, np.uint32)
        f.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        f(n_reads, masked.size, masked.ctypes.data, adj_off.ctypes.data, adj.ctypes.data, init.ctypes.data, seeds.size,
          seeds.ctypes.data, out.ctypes.data)
        return out

    def clip_read(self, ops, pos_2_1, left, right):
        """the CIGAR walk of modify_GFA's read clipping -> (status, [posOnReadStart, posOnReadEnd, posOnCIGARStart, posOnCIGAREnd])"""
        ops = _arr(ops, np.uint32)
        out = np.zeros(4, np.int32)
        f = self.lib.hso_clip_read
        f.restype = C.c_int32
        f.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        st = f(ops.ctypes.data, ops.size, int(pos_2_1), int(left), int(right), out.ctypes.data)
        return int(st), out

    def edlib_align(self, query: bytes, target: bytes, k=-1, mode=2, task=2):
        m, n = len(query), len(target)
        ed = np.zeros(1, np.int32)
        al = np.zeros(1, np.int32)
        nl = np.zeros(1, np.int32)
        ends = np.zeros(n + 2, np.int32)
        starts = np.zeros(n + 2, np.int32)
        alen = np.zeros(1, np.int32)
        aln = np.zeros(m + n + 2, np.uint8)
        st = self.lib.hso_edlib_align(query, m, target, n, k, mode, task, ed.ctypes.data, al.ctypes.data,
                                      nl.ctypes.data, ends.ctypes.data, starts.ctypes.data, alen.ctypes.data,
                                      aln.ctypes.data)
        nloc = int(nl[0])
        has_start = task >= 1 and m > 0 and n > 0 and int(ed[0]) >= 0  # NULL in edlib's early-return cases (:162-180)
        return dict(status=int(st), edit_distance=int(ed[0]), alphabet_length=int(al[0]),
                    end_locations=ends[:nloc].copy(), start_locations=starts[:nloc].copy() if has_start else None,
                    alignment=aln[:int(alen[0])].copy() if task == 2 else None)


class EdlibAlignConfig(C.Structure):
    _fields_ = [("k", C.c_int), ("mode", C.c_int), ("task", C.c_int), ("additionalEqualities", C.c_void_p),
                ("additionalEqualitiesLength", C.c_int)]


class EdlibAlignResult(C.Structure):
    _fields_ = [("status", C.c_int), ("editDistance", C.c_int), ("endLocations", C.POINTER(C.c_int)),
                ("startLocations", C.POINTER(C.c_int)), ("numLocations", C.c_int),
                ("alignment", C.POINTER(C.c_ubyte)), ("alignmentLength", C.c_int), ("alphabetLength", C.c_int)]


class RefEdlib:
    """The vendored edlib itself (reference src/edlib), compiled into oracle/_ref/libhsref_edlib.so."""

    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(HERE, "_ref", "libhsref_edlib.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(HERE, "_ref", "libhsref_edlib.so"))
            L.edlibAlign.restype = EdlibAlignResult
            L.edlibAlign.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, EdlibAlignConfig]
            L.edlibFreeAlignResult.argtypes = [EdlibAlignResult]
            L.edlibFreeAlignResult.restype = None
            cls._lib = L
        return cls._lib

    @classmethod
    def align(cls, query: bytes, target: bytes, k=-1, mode=2, task=2):
        L = cls.lib()
        cfg = EdlibAlignConfig(k, mode, task, None, 0)
        r = L.edlibAlign(query, len(query), target, len(target), cfg)
        n = r.numLocations
        out = dict(status=r.status, edit_distance=r.editDistance, alphabet_length=r.alphabetLength,
                   end_locations=np.array([r.endLocations[i] for i in range(n)], np.int32) if r.endLocations else np.zeros(0, np.int32),
                   start_locations=(np.array([r.startLocations[i] for i in range(n)], np.int32)
                                    if r.startLocations else None),
                   alignment=(np.ctypeslib.as_array(r.alignment, shape=(r.alignmentLength,)).copy()
                              if r.alignment and r.alignmentLength else (np.zeros(0, np.uint8) if task == 2 else None)))
        L.edlibFreeAlignResult(r)
        return out


class RefSR:
    """The reference's own separate_reads.cpp functions through oracle/ref_shim_sr.cpp."""

    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(HERE, "_ref", "libhsref_sr.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(HERE, "_ref", "libhsref_sr.so"))
            L.hsref_read_pair_counts.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 7
            L.hsref_read_graph.restype = C.c_int64
            L.hsref_read_graph.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
            L.hsref_chinese_whispers.restype = None
            L.hsref_chinese_whispers.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5
            cls._lib = L
        return cls._lib

    @classmethod
    def read_pair_counts(cls, n_reads, snp_off, read_idx, code, ref_base, second_base, want_output=True):
        """list_similarities_and_differences_between_reads3 (src/separate_reads.cpp:374-433), densified"""
        snp_off = np.ascontiguousarray(snp_off, np.int64)
        read_idx = np.ascontiguousarray(read_idx, np.uint32)
        code = np.ascontiguousarray(code, np.uint8)
        ref_base = np.ascontiguousarray(ref_base, np.uint8)
        second_base = np.ascontiguousarray(second_base, np.uint8)
        sim = np.zeros((n_reads, n_reads), np.int32) if want_output else None
        diff = np.zeros((n_reads, n_reads), np.int32) if want_output else None
        cls.lib().hsref_read_pair_counts(n_reads, snp_off.shape[0] - 1, snp_off.ctypes.data, read_idx.ctypes.data,
                                         code.ctypes.data, ref_base.ctypes.data, second_base.ctypes.data,
                                         sim.ctypes.data if want_output else None, diff.ctypes.data if want_output else None)
        return sim, diff


    @classmethod
    def read_graph(cls, col, masked, error_rate):
        """create_read_graph_matrix (src/separate_reads.cpp:706-828) -> (adj_off, adj) over the masked reads, local indices"""
        n_reads, snp_off, read_idx, code, rb, sb = col
        snp_off, read_idx = np.ascontiguousarray(snp_off, np.int64), np.ascontiguousarray(read_idx, np.uint32)
        code, rb, sb = (np.ascontiguousarray(x, np.uint8) for x in (code, rb, sb))
        masked = np.ascontiguousarray(masked, np.int32)
        adj_off = np.zeros(masked.size + 1, np.int64)
        args = [n_reads, snp_off.shape[0] - 1, snp_off.ctypes.data, read_idx.ctypes.data, code.ctypes.data, rb.ctypes.data,
                sb.ctypes.data, masked.size, masked.ctypes.data, float(error_rate), adj_off.ctypes.data]
        n = cls.lib().hsref_read_graph(*args, None)
        adj = np.zeros(max(int(n), 1), np.int32)
        cls.lib().hsref_read_graph(*args, adj.ctypes.data)
        return adj_off, adj[:int(n)]

    @classmethod
    def read_graph_low_memory(cls, col, masked, error_rate):
        """the reference's create_read_graph_low_memory on SNP columns -> (adj_off, adj) over the masked reads"""
        n_reads, snp_off, idx, code, rb, sb = col
        snp_off, idx, code = _arr(snp_off, np.int64), _arr(idx, np.uint32), _arr(code, np.uint8)
        rb, sb, masked = _arr(rb, np.uint8), _arr(sb, np.uint8), _arr(masked, np.int32)
        f = cls.lib().hsref_read_graph_low_memory
        f.restype = C.c_int64
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        adj_off = np.zeros(masked.size + 1, np.int64)
        n = f(n_reads, snp_off.size - 1, snp_off.ctypes.data, idx.ctypes.data, code.ctypes.data, rb.ctypes.data, sb.ctypes.data,
              masked.size, masked.ctypes.data, float(error_rate), adj_off.ctypes.data, None)
        adj = np.zeros(max(int(n), 1), np.int32)
        f(n_reads, snp_off.size - 1, snp_off.ctypes.data, idx.ctypes.data, code.ctypes.data, rb.ctypes.data, sb.ctypes.data,
          masked.size, masked.ctypes.data, float(error_rate), adj_off.ctypes.data, adj.ctypes.data)
        return adj_off, adj[:int(n)]

    @classmethod
    def chinese_whispers(cls, n_reads, masked, adj_off, adj, init_reads):
        """chinese_whispers_high_memory (src/cluster_graph.cpp:240-310), random_device pinned; labels as read indices"""
        masked = np.ascontiguousarray(masked, np.int32)
        adj_off, adj = np.ascontiguousarray(adj_off, np.int64), np.ascontiguousarray(adj, np.int32)
        init = np.ascontiguousarray(init_reads, np.int32)
        out = np.zeros(masked.size, np.int32)
        cls.lib().hsref_chinese_whispers(n_reads, masked.size, masked.ctypes.data, adj_off.ctypes.data, adj.ctypes.data,
                                         init.ctypes.data, out.ctypes.data)
        return out


PIN_SEED = 20260117  # oracle/ref_pin_rng.cpp


class RefClip:
    """oracle/_ref/libhsref_clip.so: the reference's own read-clipping loop body (cut out of its source at build time)"""
    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(HERE, "_ref", "libhsref_clip.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(HERE, "_ref", "libhsref_clip.so"))
            L.hsref_clip_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
            L.hsref_clip_cigar.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
            cls._lib = L
        return cls._lib

    @classmethod
    def clip_read(cls, cigar: str, pos_2_1, left, right):
        out = np.zeros(4, np.int32)
        st = cls.lib().hsref_clip_read(cigar.encode(), int(pos_2_1), int(left), int(right), out.ctypes.data)
        return int(st), out

    @classmethod
    def clip_cigar(cls, cigar: str, a, b) -> str:
        buf = C.create_string_buffer(2 * len(cigar) + 64)
        n = cls.lib().hsref_clip_cigar(cigar.encode(), int(a), int(b), buf, len(buf))
        assert n >= 0
        return buf.value.decode()


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libhsref_cv.so"))


class RefCV:
    """The reference's own call_variants.cpp functions through oracle/ref_shim_cv.cpp."""

    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(HERE, "_ref", "libhsref_cv.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(HERE, "_ref", "libhsref_cv.so"))
            L.hsref_cv_create.restype = C.c_void_p
            L.hsref_cv_create.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.hsref_cv_destroy.argtypes = [C.c_void_p]
            L.hsref_cv_mean_distance.restype = C.c_float
            L.hsref_cv_mean_distance.argtypes = [C.c_void_p]
            L.hsref_cv_n_columns.restype = C.c_long
            L.hsref_cv_n_columns.argtypes = [C.c_void_p]
            L.hsref_cv_n_cells.restype = C.c_long
            L.hsref_cv_n_cells.argtypes = [C.c_void_p]
            L.hsref_cv_get_pileup.argtypes = [C.c_void_p] * 4
            L.hsref_cv_get_newref.argtypes = [C.c_void_p] * 2
            L.hsref_cv_get_read_ends.argtypes = [C.c_void_p] * 3
            L.hsref_cv_call_variants.restype = C.c_int
            L.hsref_cv_call_variants.argtypes = [C.c_void_p, C.c_float, C.c_float]
            L.hsref_cv_depth.restype = C.c_float
            L.hsref_cv_depth.argtypes = [C.c_void_p]
            L.hsref_cv_get_column_bases.argtypes = [C.c_void_p] * 3
            L.hsref_cv_list_size.restype = C.c_int
            L.hsref_cv_list_size.argtypes = [C.c_void_p, C.c_int]
            L.hsref_cv_list_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            L.hsref_cv_robust.restype = C.c_int
            L.hsref_cv_robust.argtypes = [C.c_void_p, C.c_float]
            L.hsref_cv_part_size.restype = C.c_int
            L.hsref_cv_part_size.argtypes = [C.c_void_p, C.c_int]
            L.hsref_cv_part_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
            L.hsref_cv_distance.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
            L.hsref_cv_distance_custom.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            L.hsref_chi_square.restype = C.c_float
            L.hsref_chi_square.argtypes = [C.c_int] * 4
            L.hsref_rh_order.restype = C.c_int
            L.hsref_rh_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            L.hsref_sort_desc.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, cb):
        L = self.lib()
        n = cb.n_reads
        reads = (C.c_char_p * n)(*[cb.read_str(i).encode() for i in range(n)])
        cigars = (C.c_char_p * n)(*[cb.cigar_str(i).encode() for i in range(n)])
        st = _arr(cb.start, np.int32)
        sd = _arr(cb.strand, np.uint8)
        self.h = L.hsref_cv_create(cb.contig_str().encode(), n, reads, cigars, st.ctypes.data, sd.ctypes.data)
        self.n_reads = n
        self.L = L.hsref_cv_n_columns(self.h)

    def close(self):
        if self.h:
            self.lib().hsref_cv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mean_distance(self):
        return float(np.float32(self.lib().hsref_cv_mean_distance(self.h)))

    def pileup(self):
        L = self.lib()
        n = L.hsref_cv_n_cells(self.h)
        col_off = np.zeros(self.L + 1, dtype=np.int64)
        read_idx = np.zeros(max(n, 1), dtype=np.uint32)
        code = np.zeros(max(n, 1), dtype=np.uint8)
        L.hsref_cv_get_pileup(self.h, col_off.ctypes.data, read_idx.ctypes.data, code.ctypes.data)
        return dict(col_off=col_off, read_idx=read_idx[:n], code=code[:n])

    def newref(self):
        out = np.zeros(self.L, dtype=np.uint8)
        self.lib().hsref_cv_get_newref(self.h, out.ctypes.data)
        return out

    def read_limits(self):
        s = np.zeros(self.n_reads, dtype=np.int32)
        e = np.zeros(self.n_reads, dtype=np.int32)
        self.lib().hsref_cv_get_read_ends(self.h, s.ctypes.data, e.ctypes.data)
        return s, e

    def call_variants(self, mean_error=-1.0, auto_threshold=0.33):
        L = self.lib()
        L.hsref_cv_call_variants(self.h, float(mean_error), float(auto_threshold))
        ref_base = np.zeros(self.L, dtype=np.uint8)
        second_base = np.zeros(self.L, dtype=np.uint8)
        L.hsref_cv_get_column_bases(self.h, ref_base.ctypes.data, second_base.ctypes.data)
        return dict(ref_base=ref_base, second_base=second_base, suspects=self.get_list(0), automatic=self.get_list(1),
                    depth=float(np.float32(L.hsref_cv_depth(self.h))))

    def get_list(self, which):
        L = self.lib()
        n = L.hsref_cv_list_size(self.h, which)
        pos = np.zeros(n, dtype=np.int32)
        rb = np.zeros(n, dtype=np.uint8)
        sb = np.zeros(n, dtype=np.uint8)
        L.hsref_cv_list_get(self.h, which, pos.ctypes.data, rb.ctypes.data, sb.ctypes.data)
        return dict(pos=pos, ref_base=rb, second_base=sb)

    def robust(self, mean_error=-1.0):
        """keep_only_robust_variants + merge; returns (partitions, filtered, merged)."""
        L = self.lib()
        npart = L.hsref_cv_robust(self.h, float(mean_error))
        parts = []
        for p in range(npart):
            n = L.hsref_cv_part_size(self.h, p)
            idx = np.zeros(n, dtype=np.int32)
            state = np.zeros(n, dtype=np.int16)
            more = np.zeros(n, dtype=np.int32)
            less = np.zeros(n, dtype=np.int32)
            lr = np.zeros(2, dtype=np.int32)
            L.hsref_cv_part_get(self.h, p, idx.ctypes.data, state.ctypes.data, more.ctypes.data, less.ctypes.data,
                                lr.ctypes.data)
            parts.append(dict(read_idx=idx, state=state, more=more, less=less, left=int(lr[0]), right=int(lr[1])))
        return parts, self.get_list(2), self.get_list(3)

    def distance(self, p, col, ref_base):
        out = np.zeros(10, dtype=np.int32)
        self.lib().hsref_cv_distance(self.h, p, col, int(ref_base), out.ctypes.data)
        return out

    def distance_custom(self, idx, state, more, less, col, ref_base):
        idx = _arr(idx, np.int32)
        state = _arr(state, np.int16)
        more = _arr(more, np.int32)
        less = _arr(less, np.int32)
        out = np.zeros(10, dtype=np.int32)
        self.lib().hsref_cv_distance_custom(self.h, idx.shape[0], idx.ctypes.data, state.ctypes.data,
                                            more.ctypes.data, less.ctypes.data, col, int(ref_base), out.ctypes.data)
        return out

    @classmethod
    def chi_square(cls, n00, n01, n10, n11):
        return float(np.float32(cls.lib().hsref_chi_square(int(n00), int(n01), int(n10), int(n11))))

    @classmethod
    def rh_order(cls, keys):
        keys = _arr(keys, np.uint8)
        out = np.zeros(keys.shape[0] + 8, dtype=np.uint8)
        n = cls.lib().hsref_rh_order(keys.ctypes.data, keys.shape[0], out.ctypes.data)
        return out[:n]

    @classmethod
    def sort_desc(cls, keys, counts):
        keys = _arr(keys, np.uint8).copy()
        counts = _arr(counts, np.int32).copy()
        cls.lib().hsref_sort_desc(keys.ctypes.data, counts.ctypes.data, keys.shape[0])
        return keys, counts
