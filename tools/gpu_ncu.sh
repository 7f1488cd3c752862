#!/usr/bin/env bash
# One ncu --set full capture of a kernel in a bench workload (after the same command exits 0 plain).
# usage: bash tools/gpu_ncu.sh TAG WORKLOAD KERNEL_REGEX [skip] [count]
set -u
TAG=$1; WL=$2; KR=$3; SKIP=${4:-4}; CNT=${5:-1}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload $WL --steps 5 --warmup 3 --no-variants --no-cpu-baseline --no-graph"
timeout 300 $CMD > $OUT/plain_${WL}_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c $CNT \
    -f -o $OUT/prof_${WL}_$TAG $CMD > $OUT/ncu_full_${WL}_$TAG.log 2>&1
echo "ncu full $WL exit $?"
tail -n 3 $OUT/ncu_full_${WL}_$TAG.log
