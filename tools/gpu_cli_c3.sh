#!/bin/bash
# The drop-in check on config 3: both command lines (reference CPU binary, krepp_b200) on the same 200k-read FASTQ against
# the same 1,000-genome index; wall time, their own elapsed lines, and the two outputs compared line by line after sorting.
# usage: bash tools/gpu_cli_c3.sh <outdir> [reads]
O=${1:-gpurun_out/cli}; N=${2:-200000}; mkdir -p $O
D=$(python -c "import sys; sys.path.insert(0,'tools'); import workload as W; print(W.ensure_c3($N, fastq_reads=$N)[0])")
T=$(nproc)
( TIMEFORMAT="wall %R s"; time oracle/_ref/krepp --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/ref_dist.tsv ) > $O/ref_dist.log 2>&1
( TIMEFORMAT="wall %R s"; time krepp_b200/_build/krepp_b200 --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/gpu_dist.tsv ) > $O/gpu_dist.log 2>&1
( TIMEFORMAT="wall %R s"; time oracle/_ref/krepp --num-threads $T place -i $D/index -q $D/reads.fq -o /tmp/ref_place.jplace ) > $O/ref_place.log 2>&1
( TIMEFORMAT="wall %R s"; time krepp_b200/_build/krepp_b200 --num-threads $T place -i $D/index -q $D/reads.fq -o /tmp/gpu_place.jplace ) > $O/gpu_place.log 2>&1
python - $O $N $T <<'PY' | tee $O/cli_c3.txt
import re, sys, json
O, N, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
def t(log, what):
    s = open(f"{O}/{log}.log").read()
    el = re.search(r"Done (?:estimating distances|placing queries), elapsed: ([0-9.eE+-]+) sec", s)
    wall = re.search(r"wall ([0-9.]+) s", s)
    w = float(wall.group(1)) if wall else float("nan")
    return (float(el.group(1)) if el else float("nan")), w
print(f"# config 3 drop-in check: {N} reads (150 bp) against the 1,000-genome index, --num-threads {T}")
for cmd in ("dist", "place"):
    re_, rw = t(f"ref_{cmd}", cmd); ge, gw = t(f"gpu_{cmd}", cmd)
    print(f"{cmd:5s} reference CPU : query {re_:8.2f} s ({N / re_:10.0f} reads/s)   wall incl. index load {rw:7.2f} s")
    print(f"{cmd:5s} krepp_b200    : query {ge:8.2f} s ({N / ge:10.0f} reads/s)   wall incl. index load {gw:7.2f} s")
a = sorted(l for l in open("/tmp/ref_dist.tsv") if not l.startswith("#"))
b = sorted(l for l in open("/tmp/gpu_dist.tsv") if not l.startswith("#"))
sa, sb = set(a), set(b)
print(f"dist TSV lines: reference {len(a)}, krepp_b200 {len(b)}, identical {len(sa & sb)}, only reference {len(sa - sb)}, only krepp_b200 {len(sb - sa)}")
for l in sorted(sa - sb)[:5]: print("  ref only:", l.rstrip())
for l in sorted(sb - sa)[:5]: print("  b200 only:", l.rstrip())
def placements(path):
    j = json.load(open(path)); out = {}
    for p in j["placements"]:
        out[p["n"][0]] = sorted(tuple(r) for r in p["p"])
    return out
pa, pb = placements("/tmp/ref_place.jplace"), placements("/tmp/gpu_place.jplace")
same = sum(1 for k in pa if pb.get(k) == pa[k])
edges = sum(1 for k in pa if k in pb and [r[0] for r in pa[k]] == [r[0] for r in pb[k]])
print(f"place jplace: reads placed reference {len(pa)}, krepp_b200 {len(pb)}; identical rows {same}; same edge sets {edges} (a tied closest reference may differ by the reference's hash order, SURVEY section 0 fact 6)")
PY
