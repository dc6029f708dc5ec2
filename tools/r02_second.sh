#!/bin/bash
# round 2, second GPU session: NEXT / TOP(SoA) configurations
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 900 python tools/cfg_perf.py c3 soup1m c4 c1 c2 > $O/cfg_perf.jsonl 2> $O/cfg_perf.err
cat $O/cfg_perf.jsonl
timeout 600 python tools/cfg_perf.py soup10m --iters 3 > $O/cfg_perf_c5.jsonl 2>> $O/cfg_perf.err
cat $O/cfg_perf_c5.jsonl
for sm in 4 12 16; do echo "swapMin=$sm"; TRQ_SWAP_MIN=$sm timeout 300 python tools/cfg_perf.py c3 soup1m --cfgs 1 2>/dev/null | grep -v primary; done > $O/sweep_swap.txt 2>&1
for rm in 8 24 32; do echo "refillMin=$rm"; TRQ_REFILL_MIN=$rm timeout 300 python tools/cfg_perf.py c3 soup1m --cfgs 1 2>/dev/null | grep -v primary; done > $O/sweep_refill.txt 2>&1
cat $O/sweep_swap.txt $O/sweep_refill.txt
tail -3 $O/cfg_perf.err
