#!/usr/bin/env bash
# A/B session on the GPU box: parity tests, then bench lines for (workload, env) pairs.
# usage: bash tools/gpu_ab.sh TAG "wl1 wl2" "ENV1=.. ENV2=..|ENVB=..|"   (variants separated by '|'; empty = default)
set -u
TAG=${1:-ab}
WLS=${2:-"patch16-fft-256-b256"}
VARS=${3:-"|TFCFFT_LINE_V1=1"}
STEPS=${4:-300}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python tools/ring_check.py > $OUT/ringcheck_$TAG.log 2>&1
RC=$?
cat $OUT/ringcheck_$TAG.log | tail -n 12
if [ $RC -ne 0 ]; then echo "ring_check failed or hung (exit $RC): stopping"; exit 1; fi
if [ "${SKIP_PYTEST:-0}" != "1" ]; then
timeout 700 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | head -20
fi
IFS='|' read -ra VV <<< "$VARS"
[ ${#VV[@]} -eq 0 ] && VV=("")
for WL in $WLS; do
  for V in "${VV[@]}"; do
    F=$OUT/bench_${WL}_${TAG}_$(echo "$V" | tr -c 'A-Za-z0-9' '_').json
    env $V timeout 300 python bench.py --workload $WL --steps $STEPS --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_$TAG.err
    python - "$F" "$WL" "$V" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    print(f"{wl:28s} [{v or 'default':24s}] {d['value']:12.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:.0f}  sm {d['clocks']['sm_mhz']}")
except Exception as e:
    print(wl, v, "failed", e)
PY
  done
done
tail -n 5 $OUT/bench_$TAG.err 2>/dev/null
