#!/bin/bash
# round 2, third GPU session: short-stack configurations (bigger L1), L2 streaming hints
O=gpurun_out/r02c
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 900 python tools/cfg_perf.py c3 soup1m c4 c1 c2 > $O/cfg_perf.jsonl 2> $O/cfg_perf.err
cat $O/cfg_perf.jsonl
timeout 600 python tools/cfg_perf.py soup10m --iters 3 > $O/cfg_perf_c5.jsonl 2>> $O/cfg_perf.err
cat $O/cfg_perf_c5.jsonl
for h in 1 2 3; do echo "l2hint=$h"; TRQ_L2_HINT=$h timeout 300 python tools/cfg_perf.py c3 soup1m --cfgs 0,4 2>/dev/null | grep -v primary; done > $O/sweep_l2hint.txt 2>&1
for c in 50 100; do echo "carveout=$c"; TRQ_CARVEOUT=$c timeout 300 python tools/cfg_perf.py c3 soup1m --cfgs 0,4 2>/dev/null | grep -v primary; done > $O/sweep_carve.txt 2>&1
cat $O/sweep_l2hint.txt $O/sweep_carve.txt
tail -3 $O/cfg_perf.err
