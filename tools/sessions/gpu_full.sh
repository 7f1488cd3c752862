#!/usr/bin/env bash
# GPU session: scheduling self-check, the whole GPU suite, then the default bench line and the main workloads.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-full}
timeout 400 python tools/pipe_check.py > $OUT/pipecheck_$TAG.log 2>&1; echo "pipe_check exit $?"
grep -E "MISMATCH|PASS|FAIL" $OUT/pipecheck_$TAG.log | head
timeout 120 python tools/ring_check.py > $OUT/ringcheck_$TAG.log 2>&1; echo "ring_check exit $?"; tail -n 1 $OUT/ringcheck_$TAG.log
export TFC_SAMPLES_DIR=$PWD/tests/_local_samples
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 2 $OUT/smoke_$TAG.log
for WL in global-fft-512-b32 combined-512-b32 patch4-fft-256-b256 patch16-fft-512-b64 global-fft-256-b64; do
  F=$OUT/bench_${WL}_$TAG.json
  timeout 300 python bench.py --workload $WL --steps 300 --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_$TAG.err
  python - "$F" "$WL" <<'PY'
import json, sys
f, wl = sys.argv[1:3]
try:
    d = json.load(open(f))
    print(f"{wl:28s} {d['value']:12.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f}  launches/step {d['roofline'].get('launches_per_step')}")
except Exception as e:
    print(wl, "failed", e)
PY
done
tail -n 3 $OUT/bench_$TAG.err 2>/dev/null
