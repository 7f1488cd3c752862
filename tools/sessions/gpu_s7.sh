#!/usr/bin/env bash
# GPU session: quad combine for 256 x 256 tiles -- parity, scheduling self-check, A/B against the one-thread item and
# over the occupancy target of the launch (variant libraries).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-s7}
export TFC_SAMPLES_DIR=$PWD/tests/_local_samples
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | head -20
timeout 500 python tools/pipe_check.py > $OUT/pipecheck_$TAG.log 2>&1; echo "pipe_check exit $?"
grep -E "MISMATCH|PASS|FAIL|^quad" $OUT/pipecheck_$TAG.log | head -30
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
    print(f"{wl:22s} [{v:48s}] {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  module {mp.get('ms_per_step'):.4f} graph {d.get('graph',{}).get('ms_per_step'):.4f} eager {d.get('eager',{}).get('ms_per_step'):.4f} host_us {d.get('eager',{}).get('host_us_per_call'):.1f}")
except Exception as e:
    print(wl, v, "failed", e)
PY
}
L=$PWD/tfc-gan_b200
for WL in global-fft-256-b64 global-fft-256-b64-rgb; do
  run $WL ""
  run $WL "TFCFFT_COMBINE_V1=1"
  run $WL "TFCFFT_LIB=$L/libtfcfft_cq4.so"
  run $WL "TFCFFT_LIB=$L/libtfcfft_cq6.so"
  run $WL "TFCFFT_LIB=$L/libtfcfft_cq8.so"
  run $WL ""
done
CMD="python bench.py --workload global-fft-256-b64 --steps 5 --warmup 3 --no-variants --no-cpu-baseline --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_quad_$TAG.csv $CMD > $OUT/ncu_list_quad_$TAG.log 2>&1
grep -E "combine|sub_fwd4|sub_inv4" $OUT/launches_quad_$TAG.csv | tail -6 | awk -F'","' '{print $5, $NF}'
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
