#!/bin/bash
# round 2, closing session on 1 GPU: the driver's checks on the committed build, every bench line, ncu evidence, sanitizer pass
O=gpurun_out/r02z
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/smi.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_c3_reference.json 2> $O/bench_c3_reference.err ) 2>&1 | grep real
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z/bench_c3.json'))
print('value', d['value'], 'kernel frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['host_path_equals_device_path'], 'e2e16', d['e2e_hit16']['value'], 'pipe', d['e2e_pipelined']['value'], 'clocks', d['clocks'])
print('parity', d['cpu_baseline']['parity_on_sample'], 'cpu', d['cpu_baseline']['value'])
for w,e in (d.get('workloads') or {}).items(): print(w, e.get('value'), e.get('ms_per_step'), e.get('parity_on_sample'), e.get('error'), e.get('skipped'))
PY
for w in c1 c2 c4 soup1m; do timeout 400 python bench.py --workload $w --steps 50 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; done
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
# launch list of the default command (cold-cache, serialised: compare SHARES) + full captures of the traversal kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_list.log 2>&1
for w in c3 soup1m c1 c2 c4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_packed -s 6 -c 1 -f -o $O/prof_trace_$w \
    python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_full_$w.log 2>&1; echo "ncu $w rc=$?"
done
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py > $O/sanitize_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done' $O/sanitize_$tool.log | tr '\n' ' ')"
done
ls -la $O | head -50
