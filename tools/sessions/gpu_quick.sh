#!/usr/bin/env bash
# Quick GPU check: parity tests + bench lines for selected workloads (no ncu).
# usage: bash tools/gpu_quick.sh TAG "workload1 workload2 ..."
set -u
TAG=${1:-q}
WLS=${2:-"patch16-fft-256-b256"}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; tail -n 4 $OUT/pytest_$TAG.log
for WL in $WLS; do
  timeout 300 python bench.py --workload $WL --steps 300 --warmup 20 --no-variants --no-cpu-baseline > $OUT/bench_${WL}_$TAG.json 2>> $OUT/bench_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${WL}_$TAG.json"))
    print("$WL", round(d["value"]), "img/s  frac", round(d["roofline"]["frac"], 4), " ms/step", round(d["ms_per_step"], 4), " e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$WL failed", e)
PY
done
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
