#!/usr/bin/env bash
# GPU session: ncu --set full captures (source-level stall pages) of the patch-16 ring kernel, the three global-256
# launches and the three 4-patch launches of the current build.
set -u
TAG=${1:-p3}
bash tools/gpu_ncu.sh $TAG patch16-fft-256-b256 "line_ring" 3 1
bash tools/gpu_ncu.sh $TAG global-fft-256-b64 "combine_kernel|sub_fwd4|sub_inv4" 6 3
bash tools/gpu_ncu.sh $TAG patch4-fft-256-b256 "combine_kernel|sub_fwd|sub_inv" 12 3
