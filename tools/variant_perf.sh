#!/bin/bash
# Developer tool: tools/variant_perf.sh OUTDIR "workloads" variant [variant ...]  -- cfg 0 of the shipped library and of each
# instrumented build in tools/variants/ (tools/build_variant.sh) on the given workloads, twice each (run-to-run spread)
O=$1; W=$2; shift 2
mkdir -p $O
for rep in 1 2; do
  for v in base "$@"; do
    if [ "$v" = base ]; then unset TRQ_LIB; else export TRQ_LIB=$PWD/tools/variants/libtracer_rq_$v.so; fi
    timeout 300 python tools/cfg_perf.py $W --cfgs 0 2>/dev/null | sed "s/^{/{\"variant\": \"$v\", /" >> $O/variants.jsonl
  done
done
python - "$O" <<'PY'
import json, sys, collections
rows = collections.defaultdict(list)
for l in open(sys.argv[1] + '/variants.jsonl'):
    d = json.loads(l); rows[(d['workload'], d['variant'])].append(d.get('mrays_s'))
for k, v in rows.items(): print(k[0], k[1], v)
PY
