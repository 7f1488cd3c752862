#!/usr/bin/env bash
# GPU session (final records of round 2): suite, smoke, three short bench lines, default bench line, reference arm,
# ncu launch lists, one --set full capture of the three global-256 launches.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-final}
export TFC_SAMPLES_DIR=$PWD/tests/_local_samples
timeout 400 python tools/pipe_check.py > $OUT/pipecheck_$TAG.log 2>&1; echo "pipe_check exit $?"
grep -E "MISMATCH|PASS|FAIL" $OUT/pipecheck_$TAG.log | head -5
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $OUT/pytest_$TAG.log 2>&1
PT=$?; echo "pytest exit $PT"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 2 $OUT/smoke_$TAG.log
run() {
  WL=$1; V=$2
  F=$OUT/bench_${WL}_${TAG}_$(echo "$V" | tr -c 'A-Za-z0-9' '_').json
  env $V timeout 300 python bench.py --workload $WL --steps 500 --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_$TAG.err
  python - "$F" "$WL" "$V" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    mp = d.get("module_path", {})
    print(f"{wl:22s} [{v:16s}] {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  module {mp.get('ms_per_step'):.4f} graph {d.get('graph',{}).get('ms_per_step'):.4f} eager {d.get('eager',{}).get('ms_per_step'):.4f}")
except Exception as e:
    print(wl, v, "failed", e)
PY
}
run global-fft-256-b64 "X=1"
run global-fft-256-b64-rgb "X=1"
run patch16-fft-256-b256 "X=1"
timeout 900 python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "default bench exit $?"
cut -c1-400 $OUT/bench_default_$TAG.json
timeout 300 python bench.py --impl reference --steps 50 --warmup 3 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_default_$TAG.err; echo "ref bench exit $?"
cut -c1-300 $OUT/bench_ref_$TAG.json
for WL in global-fft-256-b64 patch16-fft-256-b256; do
  CMD="python bench.py --workload $WL --steps 5 --warmup 3 --no-variants --no-cpu-baseline --no-graph"
  timeout 300 $CMD > $OUT/plain_${WL}_$TAG.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/launches_${WL}_$TAG.csv $CMD > $OUT/ncu_list_${WL}_$TAG.log 2>&1
  echo "ncu launch list $WL exit $?"
done
bash tools/gpu_ncu.sh $TAG global-fft-256-b64 "combine_kernel|sub_fwd4|sub_inv4" 6 3
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
