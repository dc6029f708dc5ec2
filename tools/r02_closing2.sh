#!/bin/bash
# round 2, closing session repeated after the automatic ray ordering went in (trace kernel: one predicate; sort passes rewritten):
# tests, smoke, both bench arms, the C5 line with its unhinted figure, launch list, sanitizer. (ncu full captures: r02_closing.sh)
O=gpurun_out/r02zz
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_c3_reference.json 2> $O/bench_c3_reference.err ) 2>&1 | grep real
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zz/bench_c3.json'))
print('value', d['value'], 'kernel frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['host_path_equals_device_path'], 'e2e16', d['e2e_hit16']['value'], 'pipe', d['e2e_pipelined']['value'])
print('parity', d['cpu_baseline']['parity_on_sample'], 'cpu', d['cpu_baseline']['value'])
for w,e in (d.get('workloads') or {}).items(): print(w, e.get('value'), e.get('ms_per_step'), e.get('parity_on_sample'), e.get('unhinted'), e.get('error'), e.get('skipped'))
PY
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02zz/bench_c5.json')); print('c5', d['value'], d.get('unhinted'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_list.log 2>&1
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py > $O/sanitize_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done' $O/sanitize_$tool.log | tr '\n' ' ')"
done
