#!/usr/bin/env bash
# GPU session: identical-tile flags (new test), full suite, headline regression check, ncu --set full of the three launches.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-s4}
export TFC_SAMPLES_DIR=$PWD/tests/_local_samples
timeout 300 python -m pytest tests/test_module_path_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "equals_real" > $OUT/pytest_eq_$TAG.log 2>&1
echo "pytest eq exit $?"; tail -n 15 $OUT/pytest_eq_$TAG.log
timeout 400 python tools/pipe_check.py > $OUT/pipecheck_$TAG.log 2>&1; echo "pipe_check exit $?"
grep -E "MISMATCH|PASS|FAIL" $OUT/pipecheck_$TAG.log | head -5
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | head -20
for WL in global-fft-256-b64 patch4-fft-256-b256 global-fft-512-b32; do
  F=$OUT/bench_${WL}_$TAG.json
  timeout 300 python bench.py --workload $WL --steps 300 --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_$TAG.err
  python - "$F" "$WL" <<'PY'
import json, sys
f, wl = sys.argv[1:3]
try:
    d = json.load(open(f))
    mp = d.get("module_path", {})
    print(f"{wl:22s} {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  module {mp.get('ms_per_step')} graph {d.get('graph',{}).get('ms_per_step')} eager {d.get('eager',{}).get('ms_per_step')}")
except Exception as e:
    print(wl, "failed", e)
PY
done
bash tools/gpu_ncu.sh $TAG global-fft-256-b64 "combine_kernel|sub_fwd4|sub_inv4" 6 3
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
