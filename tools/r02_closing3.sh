#!/bin/bash
# round 2, last refresh after the staged configuration became "whole tree only": tests, the default line, the C1 / C2 lines
O=gpurun_out/r02zzz
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?"
for w in c1 c2; do timeout 400 python bench.py --workload $w --steps 50 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zzz/bench_c3.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e_hit16']['value'], d['e2e_pipelined']['value'], [(k, v.get('value'), v.get('kernel_config')) for k, v in d['workloads'].items()], d['workloads']['c5'].get('unhinted', {}).get('value'))
for w in ('c1','c2'):
    e=json.load(open(f'gpurun_out/r02zzz/bench_{w}.json')); print(w, e['value'], e['ms_per_step'], e['kernel_config'], e['e2e']['value'], e['roofline']['frac'])
PY
