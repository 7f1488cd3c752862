#!/usr/bin/env bash
# GPU session (round 2, third sitting): full validation of the tree, module-path check, phase trace, even lane splits.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-s3}
export TFC_SAMPLES_DIR=$PWD/tests/_local_samples
timeout 400 python tools/pipe_check.py > $OUT/pipecheck_$TAG.log 2>&1; echo "pipe_check exit $?"
grep -E "MISMATCH|PASS|FAIL" $OUT/pipecheck_$TAG.log | head -5
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 2 $OUT/smoke_$TAG.log
timeout 200 python tools/trace_sub.py 64 1 > $OUT/trace_sub_$TAG.log 2>&1; cat $OUT/trace_sub_$TAG.log
run() {
  WL=$1; V=$2; X=${3:---no-graph}
  F=$OUT/bench_${WL}_${TAG}_$(echo "$V$X" | tr -c 'A-Za-z0-9' '_').json
  env $V timeout 300 python bench.py --workload $WL --steps 300 --warmup 20 --no-variants --no-cpu-baseline $X > $F 2>> $OUT/bench_$TAG.err
  python - "$F" "$WL" "$V $X" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    mp = d.get("module_path", {})
    print(f"{wl:22s} [{v:44s}] {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  module {mp.get('ms_per_step')} graph {d.get('graph',{}).get('ms_per_step')}")
except Exception as e:
    print(wl, v, "failed", e)
PY
}
run global-fft-256-b64 "" ""
run global-fft-256-b64 "TFCFFT_SUB_LANES=2 TFCFFT_SUB_LANE_TILES=32"
run global-fft-256-b64 "TFCFFT_SUB_LANES=2 TFCFFT_SUB_LANE_TILES=16"
run global-fft-256-b64 "TFCFFT_SUB_LANES=2 TFCFFT_SUB_LANE_TILES=22"
run global-fft-256-b64-rgb "TFCFFT_SUB_LANES=2 TFCFFT_SUB_LANE_TILES=32"
run global-fft-256-b64-rgb ""
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
