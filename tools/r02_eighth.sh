#!/bin/bash
# round 2, session 8 (1 GPU): (a) can TMA gather 64-byte nodes faster than 2 x LDG.256?  (b) what can SM-driven copies move over
# PCIe?  (c) short shared-memory stack + staged top of the tree at full occupancy (cfg 2: 640x2, cfg 3: 256x5), top-size sweep
O=gpurun_out/r02o
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/smi.txt
timeout 300 tools/gather_bench 64 > $O/gather_bench.jsonl 2>&1; echo "gather_bench rc=$?"; grep -E "table_MB\": (4|64)," $O/gather_bench.jsonl | grep "1024" 
timeout 120 tools/hostlink_probe > $O/hostlink_probe.txt 2>&1; echo "hostlink rc=$?"; cat $O/hostlink_probe.txt
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q > $O/pytest_configs.log 2>&1; echo "rc=$?" >> $O/pytest_configs.log; tail -3 $O/pytest_configs.log
timeout 900 python tools/cfg_perf.py c3 c4 soup1m --cfgs 0,2,3 --tops 64,128,256,384,512,768,1024 > $O/cfg_perf.jsonl 2> $O/cfg_perf.err; echo "cfg_perf rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02o/cfg_perf.jsonl'):
    d=json.loads(l); print(d.get('workload'), d.get('cfg'), d.get('staged_nodes'), d.get('mrays_s'), d.get('same_result'), d.get('error',''))
PY
