#!/usr/bin/env bash
# GPU session: L2 look-ahead of the sub-tile forward launches (TFCFFT_SUB_LOOKAHEAD=0..3) -- parity subset, A/B.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-la}
export TFC_SAMPLES_DIR=$PWD/tests/_local_samples
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_bench_shapes_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | head -20
run() {
  WL=$1; V=$2
  F=$OUT/bench_${WL}_${TAG}_$(echo "$V" | tr -c 'A-Za-z0-9' '_' | tail -c 40).json
  env $V timeout 300 python bench.py --workload $WL --steps 300 --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_$TAG.err
  python - "$F" "$WL" "$V" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    print(f"{wl:22s} [{v[-28:]:28s}] {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  graph {d.get('graph',{}).get('ms_per_step'):.4f} eager {d.get('eager',{}).get('ms_per_step'):.4f}")
except Exception as e:
    print(wl, v, "failed", e)
PY
}
for WL in ${WLS:-global-fft-256-b64 global-fft-256-b64-rgb patch4-fft-256-b256 global-fft-512-b32 patch16-fft-512-b64}; do
  for LA in ${LAS:-0 3 1 0 3}; do
    run $WL "TFCFFT_SUB_LOOKAHEAD=$LA"
  done
done
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
