#!/bin/bash
O=gpurun_out/r02v; mkdir -p $O
for rm in 16 18 20 22; do
  TRQ_REFILL_MIN=$rm timeout 300 python tools/cfg_perf.py c1 c2 c3 c4 2>/dev/null | sed "s/^{/{\"refill\": $rm, /" >> $O/sweep_b.jsonl
done
python - <<'PY'
import json, collections
t = collections.defaultdict(dict)
for l in open('gpurun_out/r02v/sweep_b.jsonl'):
    d = json.loads(l); t[(d['workload'], d['cfg'])][d['refill']] = d['mrays_s']
for w, m in t.items():
    print(w, ' '.join(f"{k}:{v:.0f}" for k, v in sorted(m.items())))
PY
