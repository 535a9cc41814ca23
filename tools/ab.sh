#!/bin/bash
# A/B of library builds and packer hooks: one bench line per argument "ENV=.. ENV=.. [-- bench args]"
# (CLFD_LIB=ab/libclfd_b200_x.so selects a build of the same ABI).  usage: tools/ab.sh "CLFD_LIB=ab/x.so" "CLFD_N_FIXED=2 -- --cascade frontalface_default"
for cfg in "$@"; do
  envs="${cfg%%--*}"; args=""; [[ "$cfg" == *--* ]] && args="${cfg#*--}"
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra $args 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
k=d['kernels']
print('$cfg', '| fps', d['value'], 'e2e', d['e2e']['value'], 'tiles_ms', k['cascade_tiles']['ms'], 'rects', d.get('rects_total'))
"
done
