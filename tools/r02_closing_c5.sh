#!/bin/bash
# round 2, closing session, C5 part: full-set captures of the traversal kernel on the 10 M-triangle soup (sorted and unsorted queue)
O=gpurun_out/r02z
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 4 -c 1 -f -o $O/prof_trace_c5_sorted \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 4 -c 1 -f -o $O/prof_trace_c5 \
    python bench.py --workload c5 --sort 0 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full_c5u.log 2>&1; echo "ncu c5 unsorted rc=$?"
ls -la $O/*c5*
