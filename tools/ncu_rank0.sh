#!/bin/bash
# Developer tool for torchrun --no-python: rank 0 runs the python script under ncu (a few NVLink / time counters on the
# gather kernels), the other ranks run it plainly.   usage: torchrun --no-python ... bash tools/ncu_rank0.sh OUT.csv script.py [args]
OUT=$1; shift
if [ "${RANK:-0}" = "0" ]; then
  exec ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum \
       --clock-control none -k regex:"gather_send|trace_packed" -s 20 -c 8 --csv --log-file "$OUT" python "$@"
else
  exec python "$@"
fi
