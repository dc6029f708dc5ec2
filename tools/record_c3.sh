#!/bin/bash
# Developer tool, run under gpurun: the C3 bench line + ncu launch list + one full-set capture (short form of record_round.sh)
R=${1:-r01}
O=gpurun_out/$R
mkdir -p $O
timeout 400 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 4 -c 1 -o $O/prof_trace_c3 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full_c3.log 2>&1
tail -c 1500 $O/bench_c3.json
