#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
run() {
  WL=$1; V=$2
  F=$OUT/bench_${WL}_cmb_$(echo "$V" | tr -c 'A-Za-z0-9' '_').json
  env $V timeout 300 python bench.py --workload $WL --steps 200 --warmup 20 --no-variants --no-cpu-baseline > $F 2>> $OUT/bench_cmb.err
  python - "$F" "$WL" "$V" <<'PY'
import json, sys
f, wl, v = sys.argv[1:4]
try:
    d = json.load(open(f))
    print(f"{wl:26s} [{v or 'default':30s}] {d['value']:10.0f} img/s  frac {d['roofline']['frac']:.4f}  ms/step {d['ms_per_step']:.4f} host_us {d['eager']['host_us_per_call']:.1f} eager {d['eager']['value']:.0f} graph {d.get('graph',{}).get('value',0):.0f} module {d.get('module_path',{}).get('value',0):.0f}")
except Exception as e:
    print(wl, v, "failed", e)
PY
}
run combined-512-b32 "TFCFFT_COMBINED_CHUNK=8"
run combined-512-b32 "TFCFFT_COMBINED_CHUNK=16"
run combined-512-b32 "TFCFFT_COMBINED_CHUNK=32"
run global-fft-256-b64 ""
run patch16-fft-256-b256 ""
tail -n 3 $OUT/bench_cmb.err 2>/dev/null
