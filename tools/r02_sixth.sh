#!/bin/bash
# round 2, sixth GPU session (1 GPU): device-resident create / build / refit tests, new host staging
O=gpurun_out/r02f
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
grep -E "device build|passed|failed|rc=|Error|error" $O/pytest_gpu.log | tail -15
for t in 1 0; do echo "taper=$t"; TRQ_CHUNK_TAPER=$t timeout 300 python tools/e2e_sweep.py 131072 262144 524288 1048576 2>&1 | tail -4; done > $O/e2e_sweep.txt 2>&1
cat $O/e2e_sweep.txt
TRQ_LIB=$PWD/tools/variants/libtracer_rq_tl.so timeout 300 python tools/variants/tl.py > $O/timeline.txt 2>&1; tail -18 $O/timeline.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f/bench_c3.json'))
print('value', d['value'], 'e2e', d['e2e'], '\ne2e16', d.get('e2e_hit16'), '\npipe', d.get('e2e_pipelined'))
PY
timeout 300 python tools/build_perf.py c3 3 2>&1 | tail -5
