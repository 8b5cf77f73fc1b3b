"""ctypes binding of oracle/_build/libkrepp_oracle.so -- the CPU checker (tests only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libkrepp_oracle.so")
KO_MAX_TH = 16


class Params(C.Structure):
    _fields_ = [("hdist_th", C.c_uint32), ("chisq", C.c_double), ("dist_max", C.c_double), ("tau", C.c_uint32),
                ("no_filter", C.c_int), ("multi", C.c_int), ("want_lookups", C.c_int), ("want_place", C.c_int)]


class Lookup(C.Structure):
    _fields_ = [("strand", C.c_uint32), ("pos", C.c_uint32), ("rix", C.c_uint32), ("enc", C.c_uint32)]


class Minfo(C.Structure):
    _fields_ = [("strand", C.c_uint32), ("leaf_se", C.c_uint32), ("hdist_min", C.c_uint32), ("solved", C.c_uint32),
                ("match_count", C.c_double), ("mismatch_count", C.c_double), ("rho", C.c_double), ("nmers", C.c_double),
                ("hist", C.c_double * (KO_MAX_TH + 1)), ("d_llh", C.c_double), ("v_llh", C.c_double)]


class Sel(C.Structure):
    _fields_ = [("leaf_se", C.c_uint32), ("strand", C.c_uint32), ("is_closest", C.c_uint32), ("minfo_ix", C.c_uint32),
                ("d_llh", C.c_double), ("v_llh", C.c_double), ("chisq", C.c_double)]


class Place(C.Structure):
    _fields_ = [("se", C.c_uint32), ("edge", C.c_uint32), ("d_llh", C.c_double), ("v_llh", C.c_double),
                ("chisq", C.c_double), ("lwr", C.c_double), ("pendant", C.c_double), ("distal", C.c_double)]


class Read(C.Structure):
    _fields_ = [("len", C.c_uint64), ("onmers", C.c_uint32), ("wn", C.c_uint32 * 2), ("hdist_filt", C.c_uint32 * 2),
                ("n_lookups", C.c_uint32), ("n_minfo", C.c_uint32), ("n_sel", C.c_uint32), ("n_place", C.c_uint32),
                ("closest", C.c_int32), ("lookups", C.POINTER(Lookup)), ("minfo", C.POINTER(Minfo)),
                ("sel", C.POINTER(Sel)), ("place", C.POINTER(Place))]


class Seek(C.Structure):
    _fields_ = [("onmers", C.c_uint64), ("hist", (C.c_double * 8) * 2), ("match", C.c_double * 2), ("d", C.c_double * 2), ("v", C.c_double * 2),
                ("dist", C.c_double), ("found", C.c_int)]


def build() -> str:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.ko_index_load.restype = C.c_void_p
        L.ko_index_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.ko_index_free.argtypes = [C.c_void_p]
        for f in ("k", "h", "m", "nnodes"):
            getattr(L, "ko_index_" + f).restype = C.c_uint32
            getattr(L, "ko_index_" + f).argtypes = [C.c_void_p]
        for f in ("mask_hash_bp", "mask_drop_lr"):
            getattr(L, "ko_index_" + f).restype = C.c_uint64
            getattr(L, "ko_index_" + f).argtypes = [C.c_void_p]
        L.ko_index_node_name.restype = C.c_char_p
        L.ko_index_node_name.argtypes = [C.c_void_p, C.c_uint32]
        L.ko_index_is_leaf.argtypes = [C.c_void_p, C.c_uint32]
        L.ko_index_parent.restype = C.c_uint32
        L.ko_index_parent.argtypes = [C.c_void_p, C.c_uint32]
        L.ko_index_blen.restype = C.c_double
        L.ko_index_blen.argtypes = [C.c_void_p, C.c_uint32]
        L.ko_index_bucket.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ko_index_jplace_tree.restype = C.c_void_p
        L.ko_index_jplace_tree.argtypes = [C.c_void_p]
        L.ko_query_read.argtypes = [C.c_void_p, C.POINTER(Params), C.c_char_p, C.c_uint64, C.POINTER(Read)]
        L.ko_read_free.argtypes = [C.POINTER(Read)]
        L.ko_pext64.restype = C.c_uint64
        L.ko_pext64.argtypes = [C.c_uint64, C.c_uint64]
        L.ko_revcomp_bp64.restype = C.c_uint64
        L.ko_revcomp_bp64.argtypes = [C.c_uint64, C.c_uint32]
        L.ko_conv_bp64_lr64.restype = C.c_uint64
        L.ko_conv_bp64_lr64.argtypes = [C.c_uint64]
        L.ko_popcount_lr32.restype = C.c_uint32
        L.ko_popcount_lr32.argtypes = [C.c_uint32]
        L.ko_xur64_hash.restype = C.c_uint64
        L.ko_xur64_hash.argtypes = [C.c_uint64]
        L.ko_llh.restype = C.c_double
        L.ko_llh.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double]
        L.ko_brent.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_double, C.c_double,
                               C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
        L.ko_llh_tables.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ko_geom_new.restype = C.c_void_p
        L.ko_geom_new.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_char_p]
        L.ko_extract_mers.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint32, C.POINTER(C.POINTER(C.c_uint64)),
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ko_extract_mers_rho.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint32, C.POINTER(C.POINTER(C.c_uint64)),
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.ko_sketch_load.restype = C.c_void_p
        L.ko_sketch_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.ko_sketch_free.argtypes = [C.c_void_p]
        L.ko_sketch_k.restype = C.c_uint32
        L.ko_sketch_k.argtypes = [C.c_void_p]
        L.ko_sketch_rho.restype = C.c_double
        L.ko_sketch_rho.argtypes = [C.c_void_p]
        L.ko_seek_read.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint64, C.POINTER(Seek)]
        _lib = L
    return _lib


def default_params(**kw) -> Params:
    p = Params(4, 2.706, float("nan"), 2, 1, 1, 0, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def read_sketch_file(path: str) -> dict:
    """The fields of a sketch file as `krepp sketch` writes them (ref src/table.cpp:35-41, src/krepp.cpp:18-29,121-128)."""
    import numpy as np
    import struct
    with open(path, "rb") as f:
        buf = f.read()
    nk = struct.unpack_from("<Q", buf, 0)[0]
    enc = np.frombuffer(buf, "<u4", nk, 8)
    at = 8 + 4 * nk
    nrows = struct.unpack_from("<I", buf, at)[0]
    inc = np.frombuffer(buf, "<u8", nrows, at + 4)
    at += 4 + 8 * nrows
    k, w, h = buf[at], buf[at + 1], buf[at + 2]
    m, r = struct.unpack_from("<II", buf, at + 3)
    frac = buf[at + 11]
    nrows2 = struct.unpack_from("<I", buf, at + 12)[0]
    ppos = bytes(buf[at + 16:at + 16 + h])
    npos = bytes(buf[at + 16 + h:at + 16 + k])
    rho = struct.unpack_from("<d", buf, at + 16 + k)[0]
    assert at + 16 + k + 8 == len(buf) and nrows2 == nrows
    return dict(enc=enc, inc=inc, k=k, w=w, h=h, m=m, r=r, frac=frac, nrows=nrows, ppos=ppos, npos=npos, rho=rho)


def oracle_sketch_table(meta: dict, seqs) -> tuple:
    """(sorted unique row << 32 | enc keys, rho) of the sequences under a sketch's geometry, from the oracle's extract_mers walk."""
    import numpy as np
    L = lib()
    geom = L.ko_geom_new(meta["k"], meta["h"], meta["m"], meta["r"], int(meta["frac"]), meta["ppos"])
    out, n, cap = C.POINTER(C.c_uint64)(), C.c_uint64(0), C.c_uint64(0)
    est = (C.c_double * 2)(0.0, 0.0)
    for s in seqs:
        L.ko_extract_mers_rho(geom, s, len(s), meta["w"], C.byref(out), C.byref(n), C.byref(cap), est)
    keys = np.unique(np.ctypeslib.as_array(out, shape=(n.value,)).copy()) if n.value else np.zeros(0, np.uint64)
    return keys, est[1] / est[0]


class OracleSketch:
    """The sketch of one genome (`krepp sketch`) and `krepp seek` on it."""

    def __init__(self, path: str):
        err = C.create_string_buffer(512)
        self.h = lib().ko_sketch_load(path.encode(), err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.k, self.rho = lib().ko_sketch_k(self.h), lib().ko_sketch_rho(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ko_sketch_free(self.h)
            self.h = None

    def seek(self, seq: bytes, th: int = 4) -> dict:
        r = Seek()
        lib().ko_seek_read(self.h, th, seq, len(seq), C.byref(r))
        return {"onmers": r.onmers, "hist": [[int(r.hist[s][x]) for x in range(th + 1)] for s in range(2)], "match": [int(r.match[0]), int(r.match[1])],
                "d": [r.d[0], r.d[1]], "v": [r.v[0], r.v[1]], "dist": r.dist, "found": bool(r.found)}

    def tsv_row(self, name: str, seq: bytes, th: int = 4) -> str:
        r = self.seek(seq, th)
        return f"{name}\t{r['dist']:.5f}" if r["found"] else f"{name}\tNaN"


class OracleIndex:
    def __init__(self, path: str):
        err = C.create_string_buffer(512)
        self.h = lib().ko_index_load(path.encode(), err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())
        L = lib()
        self.k, self.hh, self.m, self.nnodes = (L.ko_index_k(self.h), L.ko_index_h(self.h), L.ko_index_m(self.h),
                                                 L.ko_index_nnodes(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ko_index_free(self.h)
            self.h = None

    def name(self, se: int) -> str:
        return lib().ko_index_node_name(self.h, se).decode()

    def jplace_tree(self) -> str:
        p = lib().ko_index_jplace_tree(self.h)
        s = C.string_at(p).decode()
        C.CDLL(None).free(C.c_void_p(p))
        return s

    def query(self, seq: bytes, params: Params | None = None) -> dict:
        """Runs one read; returns plain-python stage dicts (integers as int, floats as float)."""
        p = params or default_params()
        th = p.hdist_th
        r = Read()
        lib().ko_query_read(self.h, C.byref(p), seq, len(seq), C.byref(r))
        out = {
            "len": r.len, "onmers": r.onmers, "wn": (r.wn[0], r.wn[1]), "hdist_filt": (r.hdist_filt[0], r.hdist_filt[1]),
            "lookups": [(l.strand, l.pos, l.rix, l.enc) for l in (r.lookups[i] for i in range(r.n_lookups))],
            "minfo": [dict(strand=m.strand, leaf_se=m.leaf_se, hdist_min=m.hdist_min, solved=m.solved,
                           match=int(m.match_count), mismatch=m.mismatch_count, rho=m.rho,
                           hist=[int(m.hist[x]) for x in range(th + 1)], d=m.d_llh, v=m.v_llh)
                      for m in (r.minfo[i] for i in range(r.n_minfo))],
            "sel": [dict(leaf_se=s.leaf_se, strand=s.strand, is_closest=s.is_closest, d=s.d_llh, v=s.v_llh, chisq=s.chisq)
                    for s in (r.sel[i] for i in range(r.n_sel))],
            "closest": r.closest,
            "place": [dict(se=q.se, edge=q.edge, d=q.d_llh, v=q.v_llh, chisq=q.chisq, lwr=q.lwr, pendant=q.pendant,
                           distal=q.distal) for q in (r.place[i] for i in range(r.n_place))],
        }
        lib().ko_read_free(C.byref(r))
        return out


def parse_ref_dump(text: str) -> dict[int, dict]:
    """Parses the line format documented at the top of oracle/ref_dump.cpp into {read idx: stage dict}."""
    reads: dict[int, dict] = {}
    info = None
    for line in text.splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "I":
            info = dict(k=int(t[1]), h=int(t[2]), m=int(t[3]), nnodes=int(t[5]), mask_hash_bp=int(t[6], 16),
                        mask_drop_lr=int(t[7], 16))
            continue
        idx = int(t[1])
        if t[0] == "R":
            reads[idx] = dict(name=t[2], len=int(t[3]), onmers=int(t[4]), wn=(int(t[5]), int(t[6])),
                              hdist_filt=(int(t[7]), int(t[8])), lookups=[], minfo=[], href=[], cref=None, sel=[], place=[])
        elif t[0] == "L":
            reads[idx]["lookups"].append((int(t[2]), int(t[3]), int(t[4]), int(t[5])))
        elif t[0] == "M":
            reads[idx]["minfo"].append(dict(strand=int(t[2]), leaf_se=int(t[3]), match=int(float(t[4])), hdist_min=int(t[5]),
                                            rho=float(t[6]), hist=[int(float(x)) for x in t[7:]]))
        elif t[0] == "H":
            reads[idx]["href"].append(dict(leaf_se=int(t[2]), d=float(t[3]), v=float(t[4])))
        elif t[0] == "C":
            reads[idx]["cref"] = int(t[2])
        elif t[0] == "D":
            reads[idx]["sel"].append(dict(leaf_se=int(t[2]), strand=int(t[3]), d=float(t[4]), v=float(t[5]), chisq=float(t[6]),
                                          is_closest=int(t[7])))
        elif t[0] == "P":
            reads[idx]["place"].append(dict(se=int(t[2]), edge=int(t[3]), d=float(t[4]), v=float(t[5]), chisq=float(t[6]),
                                            lwr=float(t[7]), pendant=float(t[8]), distal=float(t[9])))
    return {"info": info, "reads": reads}
