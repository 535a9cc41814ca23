#!/bin/bash
# like tools/ab.sh, printing the pyramid kernels' times too
for cfg in "$@"; do
  envs="${cfg%%--*}"; args=""; [[ "$cfg" == *--* ]] && args="${cfg#*--}"
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra $args 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
k=d['kernels']
print('$cfg', '| fps', d['value'], 'e2e', d['e2e']['value'], 'resize', k['resize_colsum']['ms'], k['resize_colsum'].get('frac_of_hbm_peak'), 'colscan', k['colscan']['ms'], 'rows', k['integral_rows']['ms'], 'tilted', k['tilted']['ms'], 'tiles', k['cascade_tiles']['ms'])
"
done
