#!/bin/bash
# round 2, session 9 (1 GPU): the TMA unit as a second node-fetch path (cfg 2): correctness, then who fetches what
O=gpurun_out/r02p
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q > $O/pytest_configs.log 2>&1; echo "rc=$?" >> $O/pytest_configs.log; tail -3 $O/pytest_configs.log
timeout 300 python tools/cfg_perf.py c3 soup1m c4 --cfgs 0 > $O/base.jsonl 2>/dev/null
for lanes in 0xffffffff 0xaaaaaaaa 0x88888888 0x0; do
  for from in 0 2047; do
    [ "$lanes" = "0x0" ] && [ "$from" = "2047" ] && continue
    TRQ_TMA_LANES=$lanes TRQ_TMA_FROM=$from timeout 300 python tools/cfg_perf.py c3 soup1m c4 --cfgs 2 2>/dev/null | sed "s/^{/{\"lanes\": \"$lanes\", \"from\": $from, /" >> $O/tma.jsonl
  done
done
python - <<'PY'
import json
for f in ('base','tma'):
    for l in open(f'gpurun_out/r02p/{f}.jsonl'):
        d=json.loads(l); print(d.get('lanes','-'), d.get('from','-'), d.get('workload'), d.get('cfg'), d.get('mrays_s'), d.get('error',''))
PY
