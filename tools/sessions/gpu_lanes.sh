#!/usr/bin/env bash
# GPU session: scheduling-variant self-check of the sub-tile path, then bench A/B of the lane schedule.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/pipe_check.py > $OUT/pipecheck_lanes.log 2>&1; echo "pipe_check exit $?"
grep -E "MISMATCH|PASS|FAIL" $OUT/pipecheck_lanes.log | head -20
STEPS=${1:-300}
for WL in global-fft-256-b64 global-fft-256-b64-rgb patch4-fft-256-b256 patch16-fft-512-b64 global-fft-512-b32; do
  for V in "" "TFCFFT_SUB_LANES=2" "TFCFFT_SUB_LANES=2 TFCFFT_SUB_WAVES=2"; do
    F=$OUT/bench_${WL}_lanes_$(echo "$V" | tr -c 'A-Za-z0-9' '_').json
    env $V timeout 300 python bench.py --workload $WL --steps $STEPS --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_lanes.err
    python - "$F" "$WL" "$V" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    print(f"{wl:28s} [{v or 'default':24s}] {d['value']:12.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f} eager {d.get('eager')} graph {d.get('graph')}")
except Exception as e:
    print(wl, v, "failed", e)
PY
  done
done
tail -n 5 $OUT/bench_lanes.err 2>/dev/null
