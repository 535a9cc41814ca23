#!/bin/bash
# ncu --set full of the two k_cascade_tiles launches of one bench step (batch 8), for A/B runs:
#   tools/ncu_tiles.sh <tag> [ENV=VALUE ...]      -> gpurun_out/<tag>.ncu-rep, gpurun_out/<tag>_raw.csv
tag=$1; shift
env "$@" python bench.py --batch 8 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || exit 1
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_cascade_tiles -s 8 -c 2 -f -o gpurun_out/${tag} \
    python bench.py --batch 8 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
