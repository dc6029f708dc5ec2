#!/bin/bash
# round 2, fifth GPU session (1 GPU): e2e investigation, lane statistics, ncu captures per workload
O=gpurun_out/r02e
mkdir -p $O
python tools/pcie_probe.py > $O/pcie_probe.txt 2>&1; cat $O/pcie_probe.txt
for t in 1 0; do echo "taper=$t"; TRQ_CHUNK_TAPER=$t timeout 300 python tools/e2e_sweep.py 262144 524288 1048576 2>&1 | tail -3; done > $O/e2e_sweep.txt 2>&1
cat $O/e2e_sweep.txt
TRQ_LIB=$PWD/tools/variants/libtracer_rq_tl.so timeout 300 python tools/variants/tl.py > $O/timeline.txt 2>&1; tail -45 $O/timeline.txt
TRQ_LIB=$PWD/tools/variants/libtracer_rq_stats.so timeout 300 python tools/stats_probe.py c3 > $O/stats.txt 2>&1
TRQ_LIB=$PWD/tools/variants/libtracer_rq_stats.so timeout 300 python tools/stats_probe.py soup >> $O/stats.txt 2>&1; cat $O/stats.txt
# bench on the remaining workloads as standalone lines (e2e included)
for w in c1 c2 c4 soup1m; do timeout 400 python bench.py --workload $w --steps 50 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; done
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
# launch list of the default command + full captures of the traversal kernel per workload
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_list.log 2>&1
for w in c3 soup1m c1 c2 c4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 6 -c 1 -f -o $O/prof_trace_$w \
    python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_full_$w.log 2>&1; echo "ncu $w rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 4 -c 1 -f -o $O/prof_trace_c5_sorted \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 4 -c 1 -f -o $O/prof_trace_c5 \
    python bench.py --workload c5 --sort 0 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full_c5u.log 2>&1; echo "ncu c5 unsorted rc=$?"
ls -la $O | head -40
