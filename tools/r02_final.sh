#!/bin/bash
# round 2, closing session on 1 GPU: the driver's three checks on the committed build + sanitizer pass
O=gpurun_out/r02k
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k/bench_c3.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['host_path_equals_device_path'], 'e2e16', d['e2e_hit16']['value'], 'pipe', d['e2e_pipelined']['value'])
for w,e in (d.get('workloads') or {}).items(): print(w, e.get('value'), e.get('ms_per_step'), e.get('parity_on_sample'), e.get('error'), e.get('skipped'))
PY
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py > $O/sanitize_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done' $O/sanitize_$tool.log | tr '\n' ' ')"
done
