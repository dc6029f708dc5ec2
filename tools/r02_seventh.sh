#!/bin/bash
# round 2, seventh GPU session (1 GPU): full suite on the current build, contract bench, creation timing
O=gpurun_out/r02i
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python tools/create_perf.py c3 2>/dev/null | tee $O/create_perf.txt
timeout 300 python tools/create_perf.py soup1m 2>/dev/null | tail -2 | tee -a $O/create_perf.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i/bench_c3.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['host_path_equals_device_path'], 'e2e16', d['e2e_hit16']['value'], 'pipe', d['e2e_pipelined']['value'], 'clocks', d['clocks'])
print('roofline', {k: d['roofline'][k] for k in ('bound','achieved','peak','frac','traffic','dram_frac','ncu')})
for w,e in (d.get('workloads') or {}).items(): print(w, e.get('value'), e.get('ms_per_step'), e.get('parity_on_sample'), (e.get('roofline') or {}).get('frac'), (e.get('roofline') or {}).get('dram_frac'), e.get('error'), e.get('skipped'))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 12 -c 2 -f -o $O/prof_trace_c2 \
    python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_full_c2.log 2>&1; echo "ncu c2 rc=$?"
