#!/bin/bash
# round 2, closing multi-GPU session: gather tests on real peers (N = 2) + the contract bench line.  usage: tools/r02_multi_closing.sh N
N=${1:-2}
O=gpurun_out/r02z_$N
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.csv
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log; tail -5 $O/pytest_multi.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?"
tail -3 $O/bench_c3.err
python - <<PY
import json
d=json.load(open('$O/bench_c3.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e16', d.get('e2e_hit16',{}).get('value'))
print('gather', json.dumps(d.get('with_hit_allgather'))[:1500])
for w,e in (d.get('workloads') or {}).items(): print(w, e.get('value'), e.get('ms_per_step'), e.get('parity_on_sample'), e.get('error'), e.get('skipped'))
PY
