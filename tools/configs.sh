#!/bin/bash
# Runs the five BASELINE.json configurations on one B200 and prints one line each
# (value = device-resident frames/s, e2e = host frames in / rects out, cpu = the reference's own cvHaarDetectObjects from oracle/_ref on the host cores, clod_cpu = CLOD-CPU).
# usage: ./tools/configs.sh > gpurun_out/configs.jsonl
run() { python bench.py --steps "$1" --warmup 3 "${@:2}" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'workload': d['config']['workload'], 'windows_per_frame': d['config']['windows_per_frame'], 'value_fps': d['value'],
                  'e2e_fps': d['e2e']['value'], 'windows_per_sec': d['windows_per_sec'], 'ms_per_step': d['ms_per_step'],
                  'cpu_fps': (d.get('cpu_baseline') or {}).get('value'), 'cpu_cores': (d.get('cpu_baseline') or {}).get('cores'), 'cpu_kind': (d.get('cpu_baseline') or {}).get('kind'),
                  'clod_cpu_fps': ((d.get('cpu_baseline') or {}).get('clod_cpu') or {}).get('value'),
                  'kernels_ms': {k: v['ms'] for k, v in d['kernels'].items()}}))"; }
run 20 --cascade frontalface_alt --size 640x480 --min-size 24x24 --batch 1 --cpu-baseline-frames 16           # configs[0]
run 10 --cascade frontalface_default --batch 64 --cpu-baseline-frames 16                                      # configs[1]
run 5 --cascade frontalface_alt_tree,eye --size 3840x2160 --batch 8 --cpu-baseline-frames 4                   # configs[2]
run 5 --cascade profileface,fullbody --scale 1.1 --batch 16 --cpu-baseline-frames 8                           # configs[3]
run 128 --cascade frontalface_alt --batch 64 --no-cpu-baseline                                                # configs[4]: 8192 frames on this GPU
