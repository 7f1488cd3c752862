#!/usr/bin/env bash
# ncu launch list (per-kernel device times) for a workload. usage: bash tools/gpu_list.sh TAG WORKLOAD
set -u
TAG=$1; WL=$2; OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload $WL --steps 5 --warmup 3 --no-variants --no-cpu-baseline --no-graph"
timeout 300 $CMD > $OUT/plain_${WL}_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${WL}_$TAG.csv $CMD > $OUT/ncu_list_${WL}_$TAG.log 2>&1
echo "ncu list $WL exit $?"
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("$OUT/launches_${WL}_$TAG.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except Exception: continue
    agg[r[ki][:80]][0] += 1; agg[r[ki][:80]][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:8]:
    print(f"{t/tot*100:5.1f}%  n={n:3d} avg={t/n/1e3:9.1f}us  {k}")
PY
