#!/bin/bash
# per-launch device times (ncu, cold cache, serialised) of a bench command's last step:
#   tools/ncu_launches.sh <tag> <kernel regex> <bench args...>   -> gpurun_out/<tag>_launches.csv
tag=$1; re=$2; shift 2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:$re --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra "$@" > gpurun_out/${tag}_ncu.log 2>&1
python - "$tag" <<'PY'
import csv, sys, collections
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_launches.csv")) if len(r) > 10]
hdr = rows[0]; i_k = hdr.index("Kernel Name"); i_m = hdr.index("Metric Name"); i_v = hdr.index("Metric Value"); i_id = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[i_id], {"k": r[i_k]})[r[i_m]] = float(r[i_v].replace(",", ""))
last = list(per.values())
for v in last[-12:]:
    print(f"{v['k'][:60]:60s} {v.get('gpu__time_duration.sum', 0)/1e3:9.1f} us  rd {v.get('dram__bytes_read.sum', 0)/1e6:8.1f} MB  wr {v.get('dram__bytes_write.sum', 0)/1e6:8.1f} MB")
PY
