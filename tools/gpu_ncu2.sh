#!/bin/bash
# ncu on the config-3 chain at the bench's launch size (1M reads per batch): the launch list of one step and one `--set full`
# capture of every heavy kernel of one batch; the raw page is exported to CSV on the box (the .ncu-rep comes back too).
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_ncu2.sh <tag> [kernel regex] [launches to capture] [extra bench args]'
TAG=${1:-ncu2}; RE=${2:-lookup|join_kernel|hit_scatter_kernel|resolve_kernel|gate_kernel|solve_kernel|alias_kernel|merge_kernel|dist_|bin_sort}; CNT=${3:-12}; EXTRA=${4:-}
O=gpurun_out/$TAG; mkdir -p $O
CMD="python bench.py --reads 2000000 --batch 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-place $EXTRA"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# skip the warm-up step's two batches: capture the kernels of the third batch on
SKIP=$(python - $O/launches.csv "$RE" <<'PY'
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ik = rows[h].index("Kernel Name")
names = [r[ik] for r in rows[h + 1:] if len(r) > ik]
m = [n for n in names if re.search(sys.argv[2], n)]
print(len(m) // 2)   # two steps of two batches: the second half is the timed step
PY
)
echo "matching launches to skip: $SKIP"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -f -o $O/chain_full $CMD > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/chain_full.ncu-rep --page raw --csv > $O/chain_full_raw.csv 2>/dev/null
ls -la $O; tail -3 $O/ncu_full.log
