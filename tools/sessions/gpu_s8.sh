#!/usr/bin/env bash
# GPU session: quad combine with rows-per-CTA loop -- A/B over rows per CTA x occupancy target, then ncu of the launch.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-s8}
export TFC_SAMPLES_DIR=$PWD/tests/_local_samples
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_bench_shapes_gpu.py tests/test_module_path_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | head -20
run() {
  WL=$1; V=$2
  F=$OUT/bench_${WL}_${TAG}_$(echo "$V" | tr -c 'A-Za-z0-9' '_' | tail -c 40).json
  env $V timeout 300 python bench.py --workload $WL --steps 500 --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_$TAG.err
  python - "$F" "$WL" "$V" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    mp = d.get("module_path", {})
    print(f"{wl:22s} [{v[-28:]:28s}] {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  module {mp.get('ms_per_step'):.4f} graph {d.get('graph',{}).get('ms_per_step'):.4f} eager {d.get('eager',{}).get('ms_per_step'):.4f}")
except Exception as e:
    print(wl, v, "failed", e)
PY
}
L=$PWD/tfc-gan_b200
for WL in global-fft-256-b64 global-fft-256-b64-rgb; do
  run $WL ""
  run $WL "TFCFFT_COMBINE_V1=1"
  for T in r8m5 r4m8 r8m8; do run $WL "TFCFFT_LIB=$L/libtfcfft_$T.so"; done
done
bash tools/gpu_ncu.sh $TAG global-fft-256-b64 "combine_quad" 2 1
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
