#!/usr/bin/env bash
# GPU session: parity suite (+ tools/pipe_check.py), then A/B of an environment switch of the library, alternating runs.
#   ALT="TFCFFT_NO_DEFER=1" WLS="..." tools/gpu_envab.sh <tag>
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-envab}; ALT=${ALT:-TFCFFT_NO_DEFER=1}
export TFC_SAMPLES_DIR=$PWD/tests/_local_samples
if [ "${PIPE:-1}" = "1" ]; then timeout 400 python tools/pipe_check.py > $OUT/pipecheck_$TAG.log 2>&1; echo "pipe_check exit $?"; grep -E "MISMATCH|PASS|FAIL" $OUT/pipecheck_$TAG.log | head -5; fi
timeout 1200 python -m pytest ${PYT:-tests} -m gpu -q --timeout 300 -p no:cacheprovider -x > $OUT/pytest_$TAG.log 2>&1
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
    mp = d.get("module_path", {})
    print(f"{wl:22s} [{v[-24:]:24s}] {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  module {mp.get('ms_per_step', 0):.4f} graph {d.get('graph',{}).get('ms_per_step'):.4f} eager {d.get('eager',{}).get('ms_per_step'):.4f}")
except Exception as e:
    print(wl, v, "failed", e)
PY
}
for WL in ${WLS:-global-fft-256-b64 global-fft-256-b64-rgb patch4-fft-256-b256 global-fft-512-b32 patch16-fft-512-b64}; do
  run $WL "$ALT"; run $WL "X=1"; run $WL "$ALT"; run $WL "X=1"
done
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
