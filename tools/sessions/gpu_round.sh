#!/usr/bin/env bash
# One GPU-box session: parity tests -> smoke -> microbench -> bench -> (only if all green) ncu.
# Usage (under gpurun):  bash tools/gpu_round.sh [tag] [ncu:0|1|2]   2 = launch list + one --set full capture
set -u
TAG=${1:-r1}
NCU=${2:-1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $OUT/pytest_$TAG.log 2>&1
PT=$?
tail -n 40 $OUT/pytest_$TAG.log
echo "pytest exit $PT"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
SM=$?
tail -n 5 $OUT/smoke_$TAG.log
echo "smoke exit $SM"
[ -x tools/ubench ] && timeout 120 ./tools/ubench > $OUT/ubench_$TAG.txt 2>&1 && cat $OUT/ubench_$TAG.txt
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
BE=$?
cat $OUT/bench_$TAG.json; tail -n 5 $OUT/bench_$TAG.err
echo "bench exit $BE"
timeout 300 python bench.py --impl reference --steps 50 --warmup 3 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
cat $OUT/bench_ref_$TAG.json
if [ "$NCU" != "0" ] && [ $PT -eq 0 ] && [ $SM -eq 0 ] && [ $BE -eq 0 ]; then
  for WL in global-fft-256-b64 patch16-fft-256-b256; do
    CMD="python bench.py --workload $WL --steps 5 --warmup 3 --no-variants --no-cpu-baseline"
    timeout 300 $CMD > $OUT/plain_$WL.log 2>&1 && \
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file $OUT/launches_${WL}_$TAG.csv $CMD > $OUT/ncu_list_$WL.log 2>&1
    echo "ncu launch list $WL exit $?"
  done
  if [ "$NCU" = "2" ]; then
    WL=${3:-patch16-fft-256-b256}
    KR=${4:-resident_kernel}
    CMD="python bench.py --workload $WL --steps 5 --warmup 3 --no-variants --no-cpu-baseline"
    timeout 300 $CMD > $OUT/plain2_$WL.log 2>&1 && \
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s 4 -c 2 \
        -f -o $OUT/prof_${WL}_$TAG $CMD > $OUT/ncu_full_$WL.log 2>&1
    echo "ncu full $WL exit $?"
  fi
fi
echo done
