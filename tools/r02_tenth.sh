#!/bin/bash
# round 2, session 10 (1 GPU): refill / leaf-batch thresholds again on the leaf-specialised kernels
O=gpurun_out/r02v
mkdir -p $O
for rm in 12 16 20 24; do for lb in 8 12 16; do
  TRQ_REFILL_MIN=$rm TRQ_LEAF_BATCH=$lb timeout 300 python tools/cfg_perf.py c3 soup1m c4 --cfgs 0 2>/dev/null | sed "s/^{/{\"refill\": $rm, \"leaf\": $lb, /" >> $O/sweep.jsonl
done; done
python - <<'PY'
import json, collections
t = collections.defaultdict(dict)
for l in open('gpurun_out/r02v/sweep.jsonl'):
    d = json.loads(l); t[d['workload']][(d['refill'], d['leaf'])] = d['mrays_s']
for w, m in t.items():
    print(w, ' '.join(f"{k[0]}/{k[1]}:{v:.0f}" for k, v in sorted(m.items())))
PY
