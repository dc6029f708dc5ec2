#!/bin/bash
# round 2, fourth GPU session: full GPU test suite + the contract bench at N=1 (both arms)
O=gpurun_out/r02d
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err ) 2>&1 | grep real
echo "rc=$?"; tail -5 $O/bench_c3.err; cat $O/bench_c3.json
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_c3_reference.json 2> $O/bench_c3_reference.err ) 2>&1 | grep real
cat $O/bench_c3_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
