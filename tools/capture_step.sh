#!/bin/bash
# One whole step of the default bench under ncu (run on the GPU box, via gpurun), after the same command exited 0 without it:
#   tools/capture_step.sh <tag> [bench args...]
# -> gpurun_out/<tag>_step.ncu-rep       ncu --set full of the step's launches (batch 8)
#    gpurun_out/<tag>_launches.csv       per-launch device times of two steps (cold cache, serialised: compare shares)
#    gpurun_out/<tag>_stamp.txt          bench.kernel_source_stamp() of the tree that ran (profiles/make_traffic.py)
# then, back in the container:
#   python profiles/summarize_ncu.py gpurun_out/<tag>_step.ncu-rep profiles/<tag>_step_full_batch8.csv
#   python profiles/make_traffic.py 8 $(cat gpurun_out/<tag>_stamp.txt) profiles/<tag>_step_full_batch8.csv > profiles/traffic.json
tag=$1; shift
export CLFD_NO_OVERLAP=1   # serial order: under ncu the chunks of the overlapped schedule (2 frames each at batch 8) would be profiled as separate, tail-dominated launches
B="python bench.py --batch 8 --steps 2 --warmup 3 --no-cpu-baseline --no-extra $*"
$B > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || exit 1
python -c "import bench; print(bench.kernel_source_stamp())" > gpurun_out/${tag}_stamp.txt
n=$(python -c "import json; print(json.load(open('gpurun_out/${tag}_plain.json'))['gpu_launches'] // 2)")   # launches per step (gpu_launches counts the 2 timed steps)
ncu --set full --clock-control none --import-source on -s $((4 * n)) -c $n -f -o gpurun_out/${tag}_step $B > gpurun_out/${tag}_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * n)) -c $((2 * n)) --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu2.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
