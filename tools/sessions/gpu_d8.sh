#!/usr/bin/env bash
# GPU session: the D = 8 sub-tile path (512 x 512 tiles): parity tests, then bench against the split kernels.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider -k "512 or fuzz or dist" > $OUT/pytest_d8.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_d8.log | head -20
STEPS=${1:-200}
for WL in global-fft-512-b32 combined-512-b32; do
  for V in "" "TFCFFT_NO_D8=1"; do
    F=$OUT/bench_${WL}_d8_$(echo "$V" | tr -c 'A-Za-z0-9' '_').json
    env $V timeout 300 python bench.py --workload $WL --steps $STEPS --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_d8.err
    python - "$F" "$WL" "$V" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    print(f"{wl:28s} [{v or 'default':24s}] {d['value']:12.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f} launches {d.get('gpu_launches')}")
except Exception as e:
    print(wl, v, "failed", e)
PY
  done
done
tail -n 5 $OUT/bench_d8.err 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_global512_d8.csv python bench.py --workload global-fft-512-b32 --steps 3 --warmup 2 --no-variants --no-cpu-baseline --no-graph > $OUT/ncu_d8.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_global512_d8.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-9:]: print(r[4][:60], r[-1])
PY
