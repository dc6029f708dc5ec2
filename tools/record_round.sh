#!/bin/bash
# Developer tool, run under gpurun: produce the numbers + ncu evidence that get copied into profiles/.
#   tools/record_round.sh rNN
R=${1:-r01}
O=gpurun_out/$R
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 400 python bench.py --impl reference > $O/bench_c3_reference.json 2> $O/bench_c3_reference.err
timeout 400 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 400 python bench.py --flush-l2 --no-cpu-baseline > $O/bench_c3_flushl2.json 2>> $O/bench_c3.err
for w in c1 c2 c3path c4 soup1m; do timeout 400 python bench.py --workload $w --steps 100 > $O/bench_$w.json 2> $O/bench_$w.err; done
timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python tools/e2e_sweep.py > $O/e2e_sweep.txt 2>&1
# every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1
# the top kernel, full set, one launch each for C3 and the 1M soup
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 4 -c 1 -o $O/prof_trace_c3 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full_c3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 4 -c 1 -o $O/prof_trace_soup1m \
    python bench.py --workload soup1m --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full_soup1m.log 2>&1
tail -2 $O/pytest_gpu.log; cat $O/bench_c3_reference.json $O/bench_c3.json
