#!/bin/bash
# developer tool: sweep runtime knobs of the packed kernel on one workload
W=${1:-c3}
for lb in 1 4 8 12 16 24; do for rm in 4 8 16; do
  echo "leafBatch=$lb refillMin=$rm $(TRQ_LEAF_BATCH=$lb TRQ_REFILL_MIN=$rm python tools/quick_perf.py $W 2>/dev/null | tr '\n' ' ')"
done; done
