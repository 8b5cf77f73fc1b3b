"""Helpers shared by the GPU parity tests: run the CUDA path through the C ABI and compare with the oracle."""
from __future__ import annotations

import math

import numpy as np

import krepp_b200
import oracle_lib as O

REL_TOL = 1e-5  # north_star: distances, likelihoods and like_weight_ratio within 1e-5 relative


def rel_close(a: float, b: float, tol: float = REL_TOL) -> bool:
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-300)


def gpu_stage_dicts(batch: "krepp_b200.IBatch", res: dict, tap: np.ndarray | None):
    """Per read: dict with the same keys/layout as OracleIndex.query()."""
    out = []
    lookups = {}
    if tap is not None and len(tap):
        order = np.lexsort((tap[:, 1] & 0x7FFFFFFF, tap[:, 1] >> 31, tap[:, 0]))
        t = tap[order]
        bounds = np.searchsorted(t[:, 0], np.arange(batch.n_reads + 1))
        for i in range(batch.n_reads):
            seg = t[bounds[i]:bounds[i + 1]]
            lookups[i] = [(int(r[1] >> 31), int(r[1] & 0x7FFFFFFF), int(r[2]), int(r[3])) for r in seg]
    recs, hist, reads = res["records"], res["hist"], res["reads"]
    for i in range(batch.n_reads):
        s = reads[i]
        b, n = int(s["rec_begin"]), int(s["rec_count"])
        minfo, sel = [], []
        for j in range(b, b + n):
            r = recs[j]
            minfo.append(dict(strand=int(r["strand"]), leaf_se=int(r["leaf_se"]), hdist_min=int(r["hdist_min"]),
                              solved=int(r["flags"] & 1), match=int(r["match_count"]), rho=float(r["rho"]),
                              hist=[int(x) for x in hist[j]], d=float(r["d_llh"]), v=float(r["v_llh"])))
            if r["flags"] & 2:
                sel.append(dict(leaf_se=int(r["leaf_se"]), strand=int(r["strand"]), is_closest=int(bool(r["flags"] & 4)),
                                d=float(r["d_llh"]), v=float(r["v_llh"]), chisq=float(r["chisq"])))
        sel.sort(key=lambda e: e["leaf_se"])
        assert int(s["n_selected"]) == len(sel), (i, "n_selected", int(s["n_selected"]), len(sel))
        pb, pn = int(s["place_begin"]), int(s["place_count"])
        place = [dict(se=int(q["se"]), edge=int(q["se"]) - 1, d=float(q["d_llh"]), v=-float(q["loglik"]), chisq=float(q["chisq"]),
                      lwr=float(q["lwr"]), pendant=float(q["pendant"]), distal=float(q["distal"])) for q in res["placements"][pb:pb + pn]]
        out.append(dict(onmers=int(s["onmers"]), wn=(int(s["wn"][0]), int(s["wn"][1])),
                        hdist_filt=(int(s["hdist_filt"][0]), int(s["hdist_filt"][1])), lookups=lookups.get(i, []),
                        minfo=minfo, sel=sel, place=place))
    return out


def compare_read(i: int, g: dict, o: dict, check_lookups: bool, check_chisq: bool, stats: dict):
    """Integer stages bit-exact, floats within REL_TOL.  Raises AssertionError with context."""
    assert g["onmers"] == o["onmers"], (i, "onmers", g["onmers"], o["onmers"])
    assert g["wn"] == o["wn"], (i, "wn", g["wn"], o["wn"])
    assert g["hdist_filt"] == o["hdist_filt"], (i, "hdist_filt", g["hdist_filt"], o["hdist_filt"])
    if check_lookups:
        ol = sorted(o["lookups"], key=lambda l: (l[0], l[1]))
        assert g["lookups"] == ol, (i, "lookups", len(g["lookups"]), len(ol))
    gi = [(m["strand"], m["leaf_se"], m["match"], m["hdist_min"], tuple(m["hist"]), m["solved"]) for m in g["minfo"]]
    oi = [(m["strand"], m["leaf_se"], m["match"], m["hdist_min"], tuple(m["hist"]), m["solved"]) for m in o["minfo"]]
    assert gi == oi, (i, "minfo", gi, oi)
    for a, b in zip(g["minfo"], o["minfo"]):
        assert a["rho"] == b["rho"], (i, "rho", a["rho"], b["rho"])
        if a["solved"]:
            stats["solves"] += 1
            stats["bitexact_d"] += a["d"] == b["d"]
            stats["max_rel_d"] = max(stats["max_rel_d"], abs(a["d"] - b["d"]) / max(abs(b["d"]), 1e-300))
            assert rel_close(a["d"], b["d"]), (i, "d_llh", a, b)
            assert rel_close(a["v"], b["v"]), (i, "v_llh", a, b)
    gs = [(s["leaf_se"], s["strand"], s["is_closest"]) for s in g["sel"]]
    os_ = [(s["leaf_se"], s["strand"], s["is_closest"]) for s in o["sel"]]
    assert gs == os_, (i, "sel", gs, os_)
    if check_chisq:
        for a, b in zip(g["sel"], o["sel"]):
            assert rel_close(a["chisq"], b["chisq"], 1e-5) or abs(a["chisq"] - b["chisq"]) < 1e-9, (i, "chisq", a, b)


def compare_place(i: int, g: dict, o: dict, stats: dict):
    """Candidate edges must be the same set; d / loglik / lwr / pendant within REL_TOL (chisq and lwr are differences /
    exponentials of likelihoods, so they get an absolute floor as well)."""
    ge, oe = [p["se"] for p in g["place"]], [p["se"] for p in o["place"]]
    assert ge == oe, (i, "place edges", ge, oe)
    for a, b in zip(g["place"], o["place"]):
        stats["placements"] += 1
        assert rel_close(a["d"], b["d"]), (i, "place d", a, b)
        assert rel_close(a["v"], b["v"]), (i, "place loglik", a, b)
        assert rel_close(a["lwr"], b["lwr"]) or abs(a["lwr"] - b["lwr"]) < 1e-9, (i, "place lwr", a, b)
        assert rel_close(a["pendant"], b["pendant"]) or abs(a["pendant"] - b["pendant"]) < 1e-9, (i, "pendant", a, b)
        assert a["distal"] == b["distal"], (i, "distal", a, b)
        assert rel_close(a["chisq"], b["chisq"]) or abs(a["chisq"] - b["chisq"]) < 1e-7, (i, "place chisq", a, b)


def check_gated_subset(gated: dict, full: dict):
    """The default (gated) output must be exactly the solved records of the keep-all output: per read the same records
    in the same order with bit-identical numbers (reads land in the record arrays in completion order, which differs
    from run to run), the same per-read scalars, the same closest record and the same placements."""
    gr, fr = gated["reads"], full["reads"]
    for name in ("onmers", "wn", "hdist_filt", "place_count"):
        assert np.array_equal(gr[name], fr[name]), ("gated read summaries differ in", name)
    assert len(gated["records"]) == int(((full["records"]["flags"] & 1) != 0).sum()), "gated record count"
    fields = [n for n in full["records"].dtype.names]
    for i in range(len(fr)):
        b, n = int(fr["rec_begin"][i]), int(fr["rec_count"][i])
        keep = np.nonzero((full["records"]["flags"][b:b + n] & 1) != 0)[0] + b
        gb, gn = int(gr["rec_begin"][i]), int(gr["rec_count"][i])
        assert gn == len(keep), (i, "gated rec_count", gn, len(keep))
        for name in fields:
            a, c = gated["records"][name][gb:gb + gn], full["records"][name][keep]
            assert np.array_equal(a, c, equal_nan=(a.dtype.kind == "f")), (i, "gated records differ in", name, a, c)
        assert np.array_equal(gated["hist"][gb:gb + gn], full["hist"][keep]), (i, "gated histograms differ")
        cl, gcl = int(fr["closest"][i]), int(gr["closest"][i])
        assert (cl < 0) == (gcl < 0), (i, "closest")
        if cl >= 0:
            assert gcl - gb == int(np.searchsorted(keep, cl)), (i, "closest index")
        pb, pn, qb = int(fr["place_begin"][i]), int(fr["place_count"][i]), int(gr["place_begin"][i])
        for name in full["placements"].dtype.names:
            a, c = gated["placements"][name][qb:qb + pn], full["placements"][name][pb:pb + pn]
            assert np.array_equal(a, c, equal_nan=(a.dtype.kind == "f")), (i, "gated placements differ in", name)


def run_and_compare(index_dir: str, reads: list[bytes], oracle: "O.OracleIndex", gindex: "krepp_b200.Index",
                    check_lookups: bool = True, **params) -> dict:
    th = params.get("hdist_th", 4)
    batch = krepp_b200.IBatch(gindex, reads, **params)
    if check_lookups:
        batch.enable_tap(sum(max(len(r) - oracle.k + 1, 0) for r in reads) * 2 + 16)
    # default output first: only the (strand, leaf) pairs that pass the hdist_filt gate leave the device ...
    batch.submit()
    res = batch.wait()
    gated = {k: np.array(res[k], copy=True) for k in ("reads", "records", "hist", "placements")}
    # ... then parity tap 2: every pair the reference holds before the gate, compared with the oracle stage by stage
    batch.keep_all_records()
    batch.submit()
    res = batch.wait()
    check_gated_subset(gated, res)
    tap = batch.read_tap() if check_lookups else None
    g = gpu_stage_dicts(batch, res, tap)
    place = bool(params.get("place", False))
    p = O.default_params(hdist_th=th, want_lookups=int(check_lookups), no_filter=int(params.get("no_filter", True)),
                         want_place=int(place), tau=params.get("tau", 2), chisq=params.get("chisq", 2.706))
    stats = dict(solves=0, bitexact_d=0, max_rel_d=0.0, reads=len(reads), records=len(res["records"]), placements=0)
    for i, s in enumerate(reads):
        o = oracle.query(s, p)
        compare_read(i, g[i], o, check_lookups, place or not params.get("no_filter", True), stats)
        if place:
            compare_place(i, g[i], o, stats)
    stats["alg"] = batch.algorithmic_bytes()
    batch.close()
    return stats
