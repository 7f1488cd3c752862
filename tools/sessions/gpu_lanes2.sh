#!/usr/bin/env bash
# GPU session: workspace chunk size x lane schedule sweep on the multi-wave sub-tile workloads.
set -u
OUT=gpurun_out; mkdir -p $OUT
STEPS=${1:-200}
run() {
  WL=$1; V=$2
  F=$OUT/bench_${WL}_l2_$(echo "$V" | tr -c 'A-Za-z0-9' '_').json
  env $V timeout 300 python bench.py --workload $WL --steps $STEPS --warmup 20 --no-variants --no-cpu-baseline --no-graph > $F 2>> $OUT/bench_l2.err
  python - "$F" "$WL" "$V" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    print(f"{wl:26s} [{v or 'default':52s}] {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}")
except Exception as e:
    print(wl, v, "failed", e)
PY
}
for WL in global-fft-512-b32 patch4-fft-256-b256 global-fft-256-b64-rgb patch16-fft-256-b256-rgb; do
  [ $WL = patch16-fft-256-b256-rgb ] && continue
  for V in "" "TFCFFT_WS_CHUNK_MB=32" "TFCFFT_WS_CHUNK_MB=16" "TFCFFT_SUB_LANES=2 TFCFFT_SUB_WAVES=2" "TFCFFT_SUB_LANES=2 TFCFFT_SUB_WAVES=3" "TFCFFT_SUB_LANES=2 TFCFFT_SUB_WAVES=4" "TFCFFT_SUB_LANES=2 TFCFFT_SUB_WAVES=2 TFCFFT_WS_CHUNK_MB=128"; do
    run $WL "$V"
  done
done
tail -n 3 $OUT/bench_l2.err 2>/dev/null
