#!/bin/bash
# usage: tools_sweep.sh "ENV1=.. ENV2=.." ...   (one bench run per argument, prints tile/deep ms)
for cfg in "$@"; do
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
k=d['kernels']
print('$cfg', 'fps', d['value'], 'e2e', d['e2e']['value'], 'tiles', k['cascade_tiles']['ms'], 'deep', k.get('cascade_deep',{}).get('ms'), 'deepwin', d['deep_windows_per_step'])
"
done
