"""ctypes binding of include/krepp_b200.h, shaped like the reference's seam:

    reference (C++)                                     here
    ------------------------------------------------    ---------------------------------------------
    Index(dir); load_partial_*; make_rho_partial        Index(dir, device)
    IBatch(index, qs, hdist_th, chisq, dist_max, tau,   IBatch(index, reads, hdist_th=4, chisq=2.706, dist_max=nan,
           no_filter, multi, summarize)                        tau=2, no_filter=True, multi=True, summarize=False)
    IBatch::estimate_distances(stream)                  IBatch.estimate_distances() -> TSV text (src/query.cpp:141-196)
    IBatch::node_to_minfo / Minfo fields                IBatch.results()  (numpy views of the C result structs)

Everything is computed by libkrepp_b200.so on the GPU; if the library or a device is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIB = os.path.join(_HERE, "_build", "libkrepp_b200.so")


class KreppError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[krepp_b200 error {code}] {msg}")
        self.code = code


class CapacityError(KreppError):
    """A caller-owned buffer was too small; `demand` is the number of items the call needs room for."""

    def __init__(self, demand: int, msg: str):
        super().__init__(4, msg)
        self.demand = demand


class Params(C.Structure):
    _fields_ = [("hdist_th", C.c_uint32), ("chisq", C.c_double), ("dist_max", C.c_double), ("tau", C.c_uint32),
                ("no_filter", C.c_int32), ("multi", C.c_int32), ("summarize", C.c_int32), ("place", C.c_int32)]


class IndexInfo(C.Structure):
    _fields_ = [("k", C.c_uint32), ("w", C.c_uint32), ("h", C.c_uint32), ("m", C.c_uint32), ("r", C.c_uint32),
                ("frac", C.c_uint32), ("nrows", C.c_uint32), ("nkmers", C.c_uint64), ("nnodes", C.c_uint32),
                ("nleaves", C.c_uint32), ("nsubsets", C.c_uint32), ("root_se", C.c_uint32), ("mask_hash_bp", C.c_uint64),
                ("mask_drop_lr", C.c_uint64), ("device_bytes", C.c_uint64), ("mean_bucket", C.c_double),
                ("size_biased_bucket", C.c_double)]


class ShardInfo(C.Structure):
    _fields_ = [("shard", C.c_uint32), ("nshards", C.c_uint32), ("row0", C.c_uint32), ("row1", C.c_uint32),
                ("first_entry", C.c_uint64), ("n_entries", C.c_uint64)]


class Results(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("hist_stride", C.c_uint32), ("n_records", C.c_uint64), ("n_placements", C.c_uint64),
                ("reads", C.c_void_p), ("records", C.c_void_p), ("hist", C.c_void_p), ("placements", C.c_void_p),
                ("gpu_ms", C.c_float), ("match_ms", C.c_float), ("gpu_launches", C.c_uint32), ("brief", C.c_void_p),
                ("dist_begin", C.c_void_p), ("dist_rows", C.c_void_p), ("n_dist_rows", C.c_uint64), ("dist_row_bytes", C.c_uint32),
                ("seek_dist", C.c_void_p)]


RECORD_DTYPE = np.dtype([("read", "<u4"), ("leaf_se", "<u4"), ("strand", "<u4"), ("match_count", "<u4"), ("hdist_min", "<u4"),
                         ("flags", "<u4"), ("rho", "<f8"), ("d_llh", "<f8"), ("v_llh", "<f8"), ("chisq", "<f8")])
DEVICE_NONE = -1  # KREPP_DEVICE_NONE: parse + validate only
READ_DTYPE = np.dtype([("onmers", "<u4"), ("wn", "<u4", (2,)), ("hdist_filt", "<u4", (2,)), ("rec_begin", "<u4"),
                       ("rec_count", "<u4"), ("place_begin", "<u4"), ("place_count", "<u4"), ("closest", "<i4"), ("n_selected", "<u4")])
PLACEMENT_DTYPE = np.dtype([("read", "<u4"), ("se", "<u4"), ("pendant", "<f8"), ("distal", "<f8"), ("loglik", "<f8"),
                            ("lwr", "<f8"), ("d_llh", "<f8"), ("chisq", "<f8")])
BRIEF_DTYPE = np.dtype([("read", "<u4"), ("ref", "<u4"), ("d_llh", "<f8")])  # krepp_brief_t
REC_SOLVED, REC_SELECTED, REC_CLOSEST = 1, 2, 4


def brief_from_records(records: np.ndarray, chisq_value: float) -> np.ndarray:
    """krepp_brief_t rows equivalent to full records (what the device writes with KREPP_OUT_BRIEF)."""
    out = np.zeros(len(records), BRIEF_DTYPE)
    out["read"], out["d_llh"] = records["read"], records["d_llh"]
    with np.errstate(invalid="ignore"):
        ok = (records["chisq"] < chisq_value).astype(np.uint32)
    out["ref"] = records["leaf_se"] | records["strand"] << 27 | (records["flags"] & 7) << 28 | ok << 31
    return out


def dist_rows_from_records(index: "Index", params: "Params", reads: np.ndarray, records: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """The KREPP_OUT_DIST form (dist_begin, dist_rows) derived on the host from full results: a restatement of
    report_distances' decisions (ref src/query.cpp:158-196) used to check what the device writes."""
    is_leaf = index.tree()["is_leaf"]
    rank = np.cumsum(is_leaf) - 1
    wide = index.info.nleaves > 65536
    has_max = not math.isnan(params.dist_max)
    begin = np.zeros(len(reads) + 1, np.uint32)
    rows = []
    for r, s in enumerate(reads):
        first = len(rows)
        b, n, cl = int(s["rec_begin"]), int(s["rec_count"]), int(s["closest"])
        rr = records[b:b + n]
        na = False
        picked = []
        if not params.summarize and (cl < 0 or (has_max and records[cl]["d_llh"] > params.dist_max)):
            na = True
        elif not params.summarize and not params.multi:
            picked = [records[cl]]
        else:
            for x in sorted((x for x in rr if x["flags"] & REC_SELECTED), key=lambda x: x["leaf_se"]):
                keep = (not has_max) or x["d_llh"] < params.dist_max
                if params.summarize or not params.no_filter:
                    keep = keep and bool(x["chisq"] < params.chisq)
                if keep:
                    picked.append(x)
        for x in picked:
            units = int(("%.5f" % x["d_llh"]).replace(".", ""))
            rows.append((int(x["leaf_se"]) | units << 32) if wide else (int(rank[x["leaf_se"]]) << 16 | units))
        begin[r] = first | (0x80000000 if na else 0)
    begin[len(reads)] = len(rows)
    return begin, np.array(rows, dtype=np.uint64 if wide else np.uint32)


def library_path() -> str:
    return _LIB


def build_library(verbose: bool = False) -> str:
    """Compiles csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", _CSRC] + ([] if verbose else ["-s"]), check=True)
    return _LIB


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise KreppError(3, f"{_LIB} is missing: build it with krepp_b200.build_library() / make -C krepp_b200/csrc "
                            "(there is no CPU fallback)")
    L = C.CDLL(_LIB)
    L.krepp_last_error.restype = C.c_char_p
    L.krepp_index_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.krepp_index_close.argtypes = [C.c_void_p]
    L.krepp_index_info.argtypes = [C.c_void_p, C.POINTER(IndexInfo)]
    L.krepp_index_host_checksums.argtypes = [C.c_void_p, C.c_void_p]
    L.krepp_index_node_name.restype = C.c_char_p
    L.krepp_index_node_name.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
    L.krepp_index_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.krepp_index_jplace_tree.restype = C.c_size_t
    L.krepp_index_jplace_tree.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.krepp_params_default.argtypes = [C.POINTER(Params), C.c_int]
    L.krepp_batch_create.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]
    L.krepp_batch_destroy.argtypes = [C.c_void_p]
    L.krepp_batch_host_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.krepp_batch_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.krepp_batch_submit_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
    L.krepp_batch_wait.argtypes = [C.c_void_p, C.POINTER(Results)]
    L.krepp_batch_wait_device.argtypes = [C.c_void_p, C.POINTER(Results)]
    L.krepp_batch_set_output.argtypes = [C.c_void_p, C.c_uint32]
    L.krepp_batch_reserve.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
    L.krepp_batch_enable_tap.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    L.krepp_batch_read_tap.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.krepp_batch_algorithmic_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.krepp_batch_stage_times.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.POINTER(C.c_uint32)]
    L.krepp_index_open_shard.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
    L.krepp_index_open_tree.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(C.c_void_p)]
    L.krepp_geometry_open.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
    L.krepp_geometry_open_positions.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.krepp_sequence_rho.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.krepp_sketch_write.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.krepp_sketch_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.krepp_index_open_lineages.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(C.c_void_p)]
    L.krepp_index_shard_info.argtypes = [C.c_void_p, C.POINTER(ShardInfo), C.c_void_p, C.c_uint32]
    L.krepp_index_plan_shards.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.krepp_shard_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.krepp_shard_join.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.krepp_shard_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.krepp_extract_mers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.krepp_builder_create.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_void_p)]
    L.krepp_builder_destroy.argtypes = [C.c_void_p]
    L.krepp_builder_destroy.restype = None
    L.krepp_builder_has_leaf.argtypes = [C.c_void_p, C.c_char_p]
    L.krepp_builder_leaf_rank.argtypes = [C.c_void_p, C.c_char_p]
    L.krepp_builder_leaf_rank.restype = C.c_uint32
    L.krepp_builder_nleaves.argtypes = [C.c_void_p]
    L.krepp_builder_nleaves.restype = C.c_uint32
    L.krepp_builder_add_genome.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.krepp_builder_union.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.krepp_builder_set_union.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.krepp_builder_write.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.krepp_reader_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.krepp_reader_close.argtypes = [C.c_void_p]
    L.krepp_reader_set_threads.argtypes = [C.c_void_p, C.c_uint32]
    L.krepp_reader_next.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p,
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    for f in ("krepp_format_header", "krepp_format_dist", "krepp_format_place", "krepp_format_footer"):
        getattr(L, f).restype = C.c_size_t
    L.krepp_format_seek.restype = C.c_size_t
    L.krepp_format_seek.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
    L.krepp_format_header.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.c_char_p, C.c_void_p, C.c_size_t]
    L.krepp_format_dist.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Results), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.krepp_format_place.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Results), C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                     C.c_void_p, C.c_void_p, C.c_size_t]
    L.krepp_format_footer.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.c_void_p, C.c_uint64, C.c_char_p, C.c_void_p, C.c_size_t]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise KreppError(rc, load_library().krepp_last_error().decode())


def _view(ptr, dtype, n):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (dtype.itemsize * n)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class Index:
    """Index image resident on one GPU (replaces Index + TargetIndex::load_index, src/krepp.cpp:66-108)."""

    @classmethod
    def geometry(cls, k: int = 26, w: int | None = None, h: int | None = None, m: int = 4, r: int = 1, frac: bool = True, seed: int | None = None,
                 device: int = 0, ppos: bytes | None = None) -> "Index":
        """krepp_geometry_open: the LSH geometry of a library still to be built (what `krepp sketch` / `krepp index` set up before
        reading a genome); serves extract_mers, sequence_rho and sketch_write."""
        self = cls.__new__(cls)
        L = load_library()
        self._h = C.c_void_p()
        w = k + 6 if w is None else w
        h = k - 16 if h is None else h
        if ppos is not None:  # krepp_geometry_open_positions: the caller's hash positions instead of a draw
            assert len(ppos) == h
            _check(L.krepp_geometry_open_positions(k, w, h, m, r, int(frac), bytes(ppos), device, C.byref(self._h)))
        else:
            _check(L.krepp_geometry_open(k, w, h, m, r, int(frac), -1 if seed is None else seed, device, C.byref(self._h)))
        self.info = IndexInfo()
        _check(L.krepp_index_info(self._h, C.byref(self.info)))
        self.device = device
        return self

    def sequence_rho(self, seqs) -> tuple[float, float]:
        """krepp_sequence_rho: (HyperLogLog estimate of the distinct valid k-mers, of the distinct window minimizers), summed over
        the sequences; their ratio is the genome's rho."""
        bases, offs = pack_reads(list(seqs))
        bases = np.ascontiguousarray(bases) if len(bases) else np.zeros(1, np.uint8)
        a, b = C.c_double(), C.c_double()
        _check(load_library().krepp_sequence_rho(self._h, bases.ctypes.data, offs.ctypes.data, len(offs) - 1, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sketch_write(self, seqs, path: str) -> tuple[int, float]:
        """krepp_sketch_write: `krepp sketch` of the sequences into `path`; returns (k-mers in the sketch, rho)."""
        bases, offs = pack_reads(list(seqs))
        bases = np.ascontiguousarray(bases) if len(bases) else np.zeros(1, np.uint8)
        n, rho = C.c_uint64(), C.c_double()
        _check(load_library().krepp_sketch_write(self._h, bases.ctypes.data, offs.ctypes.data, len(offs) - 1, os.fsencode(path), C.byref(n), C.byref(rho)))
        return n.value, rho.value

    def __init__(self, index_dir: str, device: int = 0, shard: int = 0, nshards: int = 1, nwk: str | None = None, lineages: str | None = None):
        """nshards > 1: this handle holds bucket-range shard `shard` of the table only (SURVEY.md 8e mode B).
        nwk: `place -t` -- a Newick file whose tree replaces the index's backbone (krepp_index_open_tree).
        lineages: `place -l` -- a Greengenes/GTDB style lineage file whose taxonomy does (krepp_index_open_lineages; not together with nwk)."""
        L = load_library()
        self._h = C.c_void_p()
        if os.path.isfile(index_dir):  # the sketch of one genome (`krepp sketch`), queried by `krepp seek`
            _check(L.krepp_sketch_open(os.fsencode(index_dir), device, C.byref(self._h)))
        elif lineages:
            _check(L.krepp_index_open_lineages(os.fsencode(index_dir), device, shard, nshards, os.fsencode(lineages), C.byref(self._h)))
        else:
            _check(L.krepp_index_open_tree(os.fsencode(index_dir), device, shard, nshards, os.fsencode(nwk) if nwk else None, C.byref(self._h)))
        self.info = IndexInfo()
        _check(L.krepp_index_info(self._h, C.byref(self.info)))
        self.shard = ShardInfo()
        self.row_splits = np.zeros(nshards + 1, np.uint32)
        _check(L.krepp_index_shard_info(self._h, C.byref(self.shard), self.row_splits.ctypes.data, nshards + 1))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load_library().krepp_index_close(self._h)
            self._h = None

    __del__ = close

    def extract_mers(self, seqs) -> np.ndarray:
        """krepp_extract_mers: the leaf table of one genome (list of its sequences as bytes) under this index's geometry --
        sorted unique row << 32 | encoding keys (RSeq::extract_mers + DynHT::fill_table, ref src/rqseq.cpp:51-144,
        src/table.cpp:248-260), computed on the GPU."""
        bases, offs = pack_reads(list(seqs))
        bases = np.ascontiguousarray(bases) if len(bases) else np.zeros(1, np.uint8)
        n = C.c_uint64()
        _check(load_library().krepp_extract_mers(self._h, bases.ctypes.data, offs.ctypes.data, len(offs) - 1, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), np.uint64)
        _check(load_library().krepp_extract_mers(self._h, bases.ctypes.data, offs.ctypes.data, len(offs) - 1, out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value]

    def host_checksums(self) -> list:
        """krepp_index_host_checksums: [sum cmer words, sum bucket ends, sum c * |leaves(c)|, sum (c+1)(leaf rank+1)] mod 2^64."""
        out = np.zeros(4, np.uint64)
        _check(load_library().krepp_index_host_checksums(self._h, out.ctypes.data))
        return [int(x) for x in out]

    def node_name(self, se: int, return_na: bool = False) -> str:
        return load_library().krepp_index_node_name(self._h, se, int(return_na)).decode()

    def tree(self):
        n = self.info.nnodes + 1
        parent, nch = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        leaf, blen = np.zeros(n, np.uint8), np.zeros(n, np.float64)
        _check(load_library().krepp_index_tree(self._h, parent.ctypes.data, nch.ctypes.data, leaf.ctypes.data, blen.ctypes.data))
        return dict(parent=parent, nchildren=nch, is_leaf=leaf, blen=blen)

    def jplace_tree(self) -> str:
        L = load_library()
        n = L.krepp_index_jplace_tree(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        L.krepp_index_jplace_tree(self._h, buf, n + 1)
        return buf.value.decode()


class LibraryBuilder:
    """krepp_builder_*: `krepp index` (IndexMultiple::build_index / save_index, ref src/krepp.cpp:164-309).  geometry: an
    Index.geometry handle (on a GPU for add_genome / union); nwk: the guide tree text or None (the reference's generated tree
    over `names`); names: reference ids in input_map.tsv order."""

    def __init__(self, geometry: "Index", nwk: str | None, names):
        L = load_library()
        self._h = C.c_void_p()
        self._geometry = geometry  # the builder borrows the handle
        enc = [n.encode() for n in names]
        arr = (C.c_char_p * max(len(enc), 1))(*enc)
        _check(L.krepp_builder_create(geometry._h, None if nwk is None else nwk.encode(), arr, len(enc), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            load_library().krepp_builder_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def nleaves(self) -> int:
        return load_library().krepp_builder_nleaves(self._h)

    def leaf_rank(self, name: str) -> int | None:
        r = load_library().krepp_builder_leaf_rank(self._h, name.encode())
        return None if r == 0xFFFFFFFF else r

    def add_genome(self, name: str, seqs) -> tuple[int, float]:
        """GPU: the genome's leaf table (kept in HBM) and rho; returns (keys in the table, rho)."""
        bases, offs = pack_reads(list(seqs))
        bases = np.ascontiguousarray(bases) if len(bases) else np.zeros(1, np.uint8)
        n, rho = C.c_uint64(), C.c_double()
        _check(load_library().krepp_builder_add_genome(self._h, name.encode(), bases.ctypes.data, offs.ctypes.data, len(offs) - 1, C.byref(n), C.byref(rho)))
        return n.value, rho.value

    def union(self) -> tuple[int, int]:
        """GPU: the union of the leaf tables; returns (distinct k-mers, distinct reference sets)."""
        n, s = C.c_uint64(), C.c_uint64()
        _check(load_library().krepp_builder_union(self._h, C.byref(n), C.byref(s)))
        return n.value, s.value

    def set_union(self, keys: np.ndarray, set_of: np.ndarray, set_begin: np.ndarray, set_leaves: np.ndarray, leaf_rho: np.ndarray | None = None):
        """krepp_builder_set_union: a union computed by the caller (host arrays)."""
        keys = np.ascontiguousarray(keys, np.uint64); set_of = np.ascontiguousarray(set_of, np.uint32)
        set_begin = np.ascontiguousarray(set_begin, np.uint64); set_leaves = np.ascontiguousarray(set_leaves, np.uint32)
        rho = None if leaf_rho is None else np.ascontiguousarray(leaf_rho, np.float64)
        assert rho is None or len(rho) == self.nleaves
        _check(load_library().krepp_builder_set_union(self._h, len(keys), keys.ctypes.data, set_of.ctypes.data, len(set_begin) - 1, set_begin.ctypes.data,
                                                      set_leaves.ctypes.data if len(set_leaves) else None, None if rho is None else rho.ctypes.data))

    def write(self, index_dir: str, seed: int = 0) -> tuple[int, int]:
        """Host: colours along the tree and the library files; returns (k-mers, colour ids incl. the null id)."""
        n, s = C.c_uint64(), C.c_uint32()
        _check(load_library().krepp_builder_write(self._h, os.fsencode(index_dir), seed, C.byref(n), C.byref(s)))
        return n.value, s.value


def plan_shards(index_dir: str, budget_bytes: int, device: int = 0) -> dict:
    """krepp_index_plan_shards: bucket-range shards needed to keep every GPU's image within budget_bytes (0 = free memory of `device`)."""
    n, whole, shard = C.c_uint32(), C.c_uint64(), C.c_uint64()
    _check(load_library().krepp_index_plan_shards(os.fsencode(index_dir), device, budget_bytes, C.byref(n), C.byref(whole), C.byref(shard)))
    return dict(nshards=n.value, whole_bytes=whole.value, shard_bytes=shard.value)


def pack_reads(reads) -> tuple[np.ndarray, np.ndarray]:
    """reads: (n, L) uint8 matrix, or a list of bytes.  Returns (bases uint8[total], offsets uint64[n+1])."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, ln = reads.shape
        return np.ascontiguousarray(reads).reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(ln))
    lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    return np.frombuffer(b"".join(reads), dtype=np.uint8), offs


class IBatch:
    """One batch of reads on one slot; mirrors IBatch (src/query.hpp:46-97)."""

    def __init__(self, index: Index, reads, names=None, hdist_th: int = 4, chisq: float = 2.706, dist_max: float = math.nan,
                 tau: int = 2, no_filter: bool = True, multi: bool = True, summarize: bool = False, place: bool = False,
                 capacity: tuple[int, int] | None = None):
        L = load_library()
        self.index = index
        self.names = names
        self.params = Params(hdist_th, chisq, dist_max, tau, int(no_filter), int(multi), int(summarize), int(place))
        self.bases, self.offsets = pack_reads(reads)
        self.n_reads = len(self.offsets) - 1
        cap_reads, cap_bases = capacity or (max(self.n_reads, 1), max(int(self.offsets[-1]), 1))
        self._h = C.c_void_p()
        rc = L.krepp_batch_create(index._h, C.byref(self.params), cap_reads, cap_bases, C.byref(self._h))
        if rc:
            msg = L.krepp_last_error().decode()
            if self._h:
                L.krepp_batch_destroy(self._h)
                self._h = None
            raise KreppError(rc, msg)
        self._res = None

    def close(self):
        if getattr(self, "_h", None):
            load_library().krepp_batch_destroy(self._h)
            self._h = None

    __del__ = close

    # -- raw pipeline -------------------------------------------------------------------------------------------
    def enable_tap(self, capacity_items: int):
        _check(load_library().krepp_batch_enable_tap(self._h, 1, capacity_items))

    def keep_all_records(self, on: bool = True):
        """Parity tap 2: also return the (strand, leaf) pairs that fail the hdist_filt gate (unsolved records)."""
        _check(load_library().krepp_batch_enable_tap(self._h, 2, int(on)))

    def pin_inputs(self):
        """Moves this batch's reads into the slot's own pinned host buffers (krepp_batch_host_buffers) so that submit()
        copies host->device straight from pinned memory, without the staging memcpy."""
        pb, po = C.c_void_p(), C.c_void_p()
        _check(load_library().krepp_batch_host_buffers(self._h, C.byref(pb), C.byref(po)))
        nb = int(self.offsets[-1])
        hb = np.frombuffer((C.c_char * max(nb, 1)).from_address(pb.value), dtype=np.uint8, count=nb)
        ho = np.frombuffer((C.c_char * (8 * (self.n_reads + 1))).from_address(po.value), dtype=np.uint64, count=self.n_reads + 1)
        hb[:] = self.bases[:nb]
        ho[:] = self.offsets - self.offsets[0]
        self.bases, self.offsets = hb, ho

    def submit(self):
        _check(load_library().krepp_batch_submit(self._h, self.bases.ctypes.data, self.offsets.ctypes.data, self.n_reads))

    def submit_host(self, bases_ptr: int, offsets_ptr: int, n_reads: int):
        """krepp_batch_submit on raw host pointers (e.g. a slice of one large page-locked buffer)."""
        _check(load_library().krepp_batch_submit(self._h, bases_ptr, offsets_ptr, n_reads))

    def submit_device(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int):
        _check(load_library().krepp_batch_submit_device(self._h, d_bases_ptr, d_offsets_ptr, n_reads, n_bases))

    # -- mode B: the three phases of a batch on a sharded index (device pointers as ints; see include/krepp_b200.h) ----
    def shard_lookup(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int, d_tuples_ptr: int, cap_tuples: int,
                     d_row_begin_ptr: int) -> np.ndarray:
        """-> send_offsets[nshards + 1]: shard g's tuples are tuples[send_offsets[g] : send_offsets[g + 1]]."""
        so = np.zeros(self.index.shard.nshards + 1, np.uint64)
        rc = load_library().krepp_shard_lookup(self._h, d_bases_ptr, d_offsets_ptr, n_reads, n_bases, d_tuples_ptr, cap_tuples,
                                               d_row_begin_ptr, so.ctypes.data)
        if rc == 4 and int(so[-1]) > cap_tuples:
            raise CapacityError(int(so[-1]), load_library().krepp_last_error().decode())
        _check(rc)
        return so

    def shard_join(self, d_tuples_ptrs, d_row_begin_ptrs, d_hits_ptr: int, cap_hits: int) -> np.ndarray:
        """-> hit_offsets[n_sources + 1]: the hit entries owed to sender s are hits[hit_offsets[s] : hit_offsets[s + 1]]."""
        n = len(d_tuples_ptrs)
        ho = np.zeros(n + 1, np.uint64)
        tp = (C.c_void_p * max(n, 1))(*d_tuples_ptrs)
        rp = (C.c_void_p * max(n, 1))(*d_row_begin_ptrs)
        rc = load_library().krepp_shard_join(self._h, n, tp, rp, d_hits_ptr, cap_hits, ho.ctypes.data)
        if rc == 4 and int(ho[-1]) > cap_hits:
            raise CapacityError(int(ho[-1]), load_library().krepp_last_error().decode())
        _check(rc)
        return ho

    def shard_finish(self, d_hits_ptr: int, n_hits: int):
        _check(load_library().krepp_shard_finish(self._h, d_hits_ptr, n_hits))

    def wait(self) -> dict:
        r = Results()
        _check(load_library().krepp_batch_wait(self._h, C.byref(r)))
        nrec = int(r.n_records)
        self._res = dict(
            reads=_view(r.reads, READ_DTYPE, r.n_reads), records=_view(r.records, RECORD_DTYPE, nrec),
            hist=_view(r.hist, np.dtype("<u4"), nrec * r.hist_stride).reshape(-1, r.hist_stride),
            placements=_view(r.placements, PLACEMENT_DTYPE, int(r.n_placements)), brief=_view(r.brief, BRIEF_DTYPE, nrec),
            dist_begin=_view(r.dist_begin, np.dtype("<u4"), r.n_reads + 1 if r.dist_begin else 0),
            dist_rows=_view(r.dist_rows, np.dtype("<u4" if r.dist_row_bytes == 4 else "<u8"), int(r.n_dist_rows) if r.dist_begin else 0),
            seek_dist=_view(r.seek_dist, np.dtype("<f8"), r.n_reads if r.seek_dist else 0),
            n_records=nrec, gpu_ms=float(r.gpu_ms), match_ms=float(r.match_ms), gpu_launches=int(r.gpu_launches))
        return self._res

    def set_output(self, records: bool = True, hist: bool = True, placements: bool = True, brief: bool = False, dist: bool = False,
                   summaries: bool = True, seek: bool = False):
        """krepp_batch_set_output: which row arrays wait() copies to the host (the others come back empty)."""
        _check(load_library().krepp_batch_set_output(self._h, int(records) | 2 * int(hist) | 4 * int(placements) | 8 * int(brief) | 16 * int(dist)
                                                     | 32 * int(summaries) | 64 * int(seek)))

    def seek_sequences(self) -> str:
        """Rows of `krepp seek` for this batch on a sketch handle (SBatch::seek_sequences, src/seek.cpp:22-53): "<id>\t<distance>",
        "<id>\tNaN" when no k-mer matched."""
        if self._res is None:
            self.set_output(seek=True)
        d = self.results()["seek_dist"]
        names = self.names if self.names is not None else [f"r{i}" for i in range(self.n_reads)]
        return "".join(f"{n}\tNaN\n" if math.isnan(x) else f"{n}\t{x:.5f}\n" for n, x in zip(names, d))

    def reserve(self, records: int = 0, hits: int = 0, nodes: int = 0, placements: int = 0):
        """krepp_batch_reserve: pre-size the result buffers (per batch) instead of letting the first batches grow them."""
        _check(load_library().krepp_batch_reserve(self._h, records, hits, nodes, placements))

    def wait_device(self) -> dict:
        """krepp_batch_wait_device: per-read summaries and counts only; record / placement rows stay in HBM."""
        r = Results()
        _check(load_library().krepp_batch_wait_device(self._h, C.byref(r)))
        return dict(reads=_view(r.reads, READ_DTYPE, r.n_reads), n_records=int(r.n_records), n_placements=int(r.n_placements),
                    gpu_ms=float(r.gpu_ms), match_ms=float(r.match_ms), gpu_launches=int(r.gpu_launches))

    def results(self) -> dict:
        if self._res is None:
            self.submit()
            self.wait()
        return self._res

    def read_tap(self) -> np.ndarray:
        L = load_library()
        n = C.c_uint64()
        _check(L.krepp_batch_read_tap(self._h, 1, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), dtype=np.uint32)
        if n.value:
            _check(L.krepp_batch_read_tap(self._h, 1, out.ctypes.data, n.value, C.byref(n)))
        return out

    def algorithmic_bytes(self) -> dict:
        b, l, e = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(load_library().krepp_batch_algorithmic_bytes(self._h, C.byref(b), C.byref(l), C.byref(e)))
        return dict(bytes=b.value, lookups=l.value, entries=e.value)

    def stage_times(self) -> list:
        """[(stage name, ms)] of the last waited batch, in launch order (CUDA events on the slot's stream)."""
        ms, names, n = (C.c_float * 24)(), (C.c_char_p * 24)(), C.c_uint32()
        _check(load_library().krepp_batch_stage_times(self._h, 24, ms, names, C.byref(n)))
        return [(names[i].decode(), float(ms[i])) for i in range(min(n.value, 24))]

    # -- reference-shaped entry points ------------------------------------------------------------------------------
    def estimate_distances(self) -> str:
        """TSV body of `krepp dist` for this batch: report_distances (src/query.cpp:158-196), precision 5, reads in
        input order and references by ascending se (the reference's own order is unspecified)."""
        res = self.results()
        p = self.params
        has_max = not math.isnan(p.dist_max)
        out = []
        recs, reads = res["records"], res["reads"]
        for i in range(self.n_reads):
            name = self.names[i] if self.names is not None else f"r{i}"
            s = reads[i]
            rr = recs[s["rec_begin"]:s["rec_begin"] + s["rec_count"]]
            sel = rr[(rr["flags"] & REC_SELECTED) != 0]
            if len(sel) == 0 or (has_max and recs[s["closest"]]["d_llh"] > p.dist_max):
                out.append(f"{name}\tNA\tNaN\n")
                continue
            if p.multi:
                for r in sorted(sel, key=lambda r: r["leaf_se"]):
                    if not p.no_filter and not (r["chisq"] < p.chisq):
                        continue
                    if has_max and not (r["d_llh"] < p.dist_max):
                        continue
                    out.append(f"{name}\t{self.index.node_name(int(r['leaf_se']))}\t{r['d_llh']:.5f}\n")
            else:
                r = recs[s["closest"]]
                out.append(f"{name}\t{self.index.node_name(int(r['leaf_se']))}\t{r['d_llh']:.5f}\n")
        return "".join(out)

    def place_sequences(self, tabular: bool = False) -> str:
        """Placement rows of this batch the way IBatch::place_sequences / report_placement frame them
        (src/query.cpp:198-333): jplace "placements" entries (PP_JPLACE_FIELDS, src/query.hpp:202-204) joined by ",\n",
        or the tab-separated rows of --tabular (PP_TABULAR_FIELDS, :206).  Reads in input order, edges by ascending se."""
        res = self.results()
        out = []
        pl, reads = res["placements"], res["reads"]
        for i in range(self.n_reads):
            s = reads[i]
            if s["place_count"] == 0:
                continue
            name = self.names[i] if self.names is not None else f"r{i}"
            rows = pl[s["place_begin"]:s["place_begin"] + s["place_count"]]
            if tabular:
                for r in rows:
                    out.append(f"{name}\t{self.index.node_name(int(r['se']), True)}\t{int(r['se']) - 1}\t{r['lwr']:.5f}\t{r['d_llh']:.5f}\n")
                continue
            f = [f"[{int(r['se']) - 1}, {r['pendant']:.5f}, {r['distal']:.5f}, {r['loglik']:.5f}, {r['lwr']:.5f}, {r['d_llh']:.5f}]" for r in rows]
            head = f'\t\t\t{{"n" : ["{name}"], "p" : ['
            if len(res["records"][s["rec_begin"]:s["rec_begin"] + s["rec_count"]][(res["records"][s["rec_begin"]:s["rec_begin"] + s["rec_count"]]["flags"] & REC_SELECTED) != 0]) == 1:
                out.append(head + f[0] + "]}")  # single-candidate shortcut (src/query.cpp:231-241)
            else:
                out.append(head + ",".join("\n\t\t\t\t" + x for x in f) + "]\n\t\t\t}")
        return "".join(out) if tabular else ",\n".join(out)


# ---- host I/O layer (krepp_reader_* / krepp_format_*): pure host code, usable with Index(dir, device=-1) ------------------

class Reader:
    """FASTA/FASTQ batch reader with kseq framing (replaces QSeq, src/rqseq.cpp:161-197)."""

    def __init__(self, path: str, threads: int = 1):
        self._h = C.c_void_p()
        _check(load_library().krepp_reader_open(os.fsencode(path), C.byref(self._h)))
        if threads > 1:
            _check(load_library().krepp_reader_set_threads(self._h, threads))

    def close(self):
        if getattr(self, "_h", None):
            load_library().krepp_reader_close(self._h)
            self._h = None

    __del__ = close

    def next_batch(self, max_reads: int = 1 << 16, max_bases: int = 1 << 24, max_name_bytes: int | None = None):
        """Returns (names, reads, eof) for the next batch; reads are bytes objects."""
        max_name_bytes = max_name_bytes or 64 * max_reads
        bases = np.empty(max(max_bases, 1), np.uint8)
        offs = np.empty(max_reads + 1, np.uint64)
        names = np.empty(max(max_name_bytes, 1), np.uint8)
        noffs = np.empty(max(max_reads, 1), np.uint64)
        n, eof = C.c_uint32(), C.c_int()
        _check(load_library().krepp_reader_next(self._h, bases.ctypes.data, max_bases, offs.ctypes.data, max_reads, names.ctypes.data,
                                                max_name_bytes, noffs.ctypes.data, C.byref(n), C.byref(eof)))
        nb = names.tobytes()
        out_names = [nb[int(noffs[i]):nb.index(b"\0", int(noffs[i]))].decode() for i in range(n.value)]
        out_reads = [bases[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n.value)]
        return out_names, out_reads, bool(eof.value)

    def read_all(self, **kw):
        names, reads = [], []
        while True:
            a, b, eof = self.next_batch(**kw)
            names += a
            reads += b
            if eof:
                return names, reads


def pack_names(names) -> tuple[np.ndarray, np.ndarray]:
    blob = b"".join(n.encode() + b"\0" for n in names)
    offs = np.zeros(max(len(names), 1), np.uint64)
    at = 0
    for i, n in enumerate(names):
        offs[i] = at
        at += len(n.encode()) + 1
    return np.frombuffer(blob + b"\0", dtype=np.uint8).copy(), offs


def results_struct(reads: np.ndarray | None, records: np.ndarray | None, hist: np.ndarray | None, placements: np.ndarray | None = None,
                   brief: np.ndarray | None = None, dist_begin: np.ndarray | None = None, dist_rows: np.ndarray | None = None) -> Results:
    """A krepp_results_t over caller-owned numpy arrays (kept alive by the caller)."""
    r = Results()
    if dist_begin is not None:  # the device-selected `dist` rows only
        r.n_reads, r.n_dist_rows, r.dist_row_bytes = len(dist_begin) - 1, len(dist_rows), dist_rows.dtype.itemsize
        r.dist_begin, r.dist_rows = dist_begin.ctypes.data, dist_rows.ctypes.data
        return r
    if records is None:  # brief rows only
        r.n_reads, r.hist_stride, r.n_records, r.n_placements = len(reads), 0, len(brief), 0
        r.reads, r.records, r.hist, r.placements, r.brief = reads.ctypes.data, None, None, None, brief.ctypes.data
        return r
    r.n_reads, r.hist_stride = len(reads), hist.shape[1] if hist.ndim == 2 else 0
    r.n_records, r.n_placements = len(records), 0 if placements is None else len(placements)
    r.reads, r.records, r.hist = reads.ctypes.data, records.ctypes.data, hist.ctypes.data
    r.placements = placements.ctypes.data if placements is not None and len(placements) else None
    return r


def _format(call) -> str:
    cap = 1 << 16
    while True:
        buf = C.create_string_buffer(cap)
        n = call(buf, cap)
        if n <= cap:
            return buf.raw[:n].decode()
        cap = n + 16


def format_header(index: Index, params: Params, tabular: bool = False, invocation: str = "") -> str:
    L = load_library()
    return _format(lambda b, c: L.krepp_format_header(index._h, C.byref(params), int(tabular), invocation.encode(), b, c))


def format_dist(index: Index, params: Params, res: Results, names, wcount: np.ndarray | None = None) -> str:
    L = load_library()
    nb, no = pack_names(names)
    w = wcount.ctypes.data if wcount is not None else None
    if wcount is not None:  # accumulate exactly once
        L.krepp_format_dist(index._h, C.byref(params), C.byref(res), nb.ctypes.data, no.ctypes.data, w, None, 0)
        return ""
    return _format(lambda b, c: L.krepp_format_dist(index._h, C.byref(params), C.byref(res), nb.ctypes.data, no.ctypes.data, None, b, c))


def format_seek(seek_dist: np.ndarray, names) -> str:
    """krepp_format_seek over a caller-owned array of distances (NaN = no match)."""
    L = load_library()
    nb, no = pack_names(names)
    d = np.ascontiguousarray(seek_dist, dtype=np.float64)
    r = Results()
    r.n_reads, r.seek_dist = len(d), d.ctypes.data
    return _format(lambda b, c: L.krepp_format_seek(C.byref(r), nb.ctypes.data, no.ctypes.data, b, c))


def format_place(index: Index, params: Params, res: Results, names, tabular: bool = False, wcount: np.ndarray | None = None) -> str:
    L = load_library()
    nb, no = pack_names(names)
    if wcount is not None:
        prev = C.c_int(0)
        L.krepp_format_place(index._h, C.byref(params), C.byref(res), nb.ctypes.data, no.ctypes.data, int(tabular), C.byref(prev),
                             wcount.ctypes.data, None, 0)
        return ""

    def call(b, c):
        prev = C.c_int(0)
        return L.krepp_format_place(index._h, C.byref(params), C.byref(res), nb.ctypes.data, no.ctypes.data, int(tabular), C.byref(prev), None, b, c)
    return _format(call)


def format_footer(index: Index, params: Params, tabular: bool = False, wcount: np.ndarray | None = None, total_queries: int = 0,
                  invocation: str = "") -> str:
    L = load_library()
    w = wcount.ctypes.data if wcount is not None else None
    return _format(lambda b, c: L.krepp_format_footer(index._h, C.byref(params), int(tabular), w, total_queries, invocation.encode(), b, c))
