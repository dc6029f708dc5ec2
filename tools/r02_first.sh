#!/bin/bash
# round 2, first GPU session: parity of the rewritten kernel + configuration sweep
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 900 python tools/cfg_perf.py c3 soup1m c1 c4 c2 > $O/cfg_perf.jsonl 2> $O/cfg_perf.err
cat $O/cfg_perf.jsonl
timeout 600 python tools/cfg_perf.py soup10m --iters 3 > $O/cfg_perf_c5.jsonl 2>> $O/cfg_perf.err
cat $O/cfg_perf_c5.jsonl
tail -5 $O/cfg_perf.err
