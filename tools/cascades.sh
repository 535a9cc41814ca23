#!/bin/bash
# Every cascade file of the reference (19) at 1080p, scale 1.2, batch 16 on one B200: one line each.
# usage: ./tools/cascades.sh > gpurun_out/cascades.jsonl
for c in frontalface_alt frontalface_default profileface eye fullbody mcs_nose frontalface_alt2 eye_tree_eyeglasses frontalface_alt_tree \
         lefteye_2splits righteye_2splits lowerbody upperbody mcs_eyepair_big mcs_eyepair_small mcs_lefteye mcs_righteye mcs_mouth mcs_upperbody; do
  python bench.py --steps 10 --warmup 3 --cascade $c --batch 16 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'cascade': '$c', 'workload': d['config']['workload'], 'value_fps': d['value'], 'e2e_fps': d['e2e']['value'],
                  'windows_per_sec': d['windows_per_sec'], 'ms_per_step': d['ms_per_step'],
                  'kernels_ms': {k: v['ms'] for k, v in d['kernels'].items()}}))"
done
