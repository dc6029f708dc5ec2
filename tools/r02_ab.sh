#!/bin/bash
# A/B of the shipped library against tools/variants/libtracer_rq_old.so (the commit before the automatic ordering) on the small-tree workloads
O=gpurun_out/r02ab; mkdir -p $O
for rep in 1 2 3; do
  for v in base old; do
    if [ "$v" = base ]; then unset TRQ_LIB; else export TRQ_LIB=$PWD/tools/variants/libtracer_rq_old.so; fi
    timeout 300 python tools/cfg_perf.py c1 c2 2>/dev/null | sed "s/^{/{\"variant\": \"$v\", /" >> $O/ab.jsonl
    timeout 300 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v bench c2', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
  done
done
python - <<'PY'
import json, collections
rows = collections.defaultdict(list)
for l in open('gpurun_out/r02ab/ab.jsonl'):
    d = json.loads(l); rows[(d['workload'], d['cfg'], d['variant'])].append(d.get('mrays_s'))
for k, v in sorted(rows.items()): print(k, v)
PY
