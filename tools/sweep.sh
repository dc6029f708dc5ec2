#!/bin/bash
# developer tool: sweep runtime knobs of the packed kernel on one workload
W=${1:-c3}
for lb in 8 12 16; do for rm in 12 16 20 24; do
  echo "leafBatch=$lb refillMin=$rm $(TRQ_LEAF_BATCH=$lb TRQ_REFILL_MIN=$rm python tools/quick_perf.py $W --iters 7 2>/dev/null | grep -o 'bounce.*mrays_s": [0-9.]*\|random.*mrays_s": [0-9.]*' | grep -o '[0-9.]*$' | tr '\n' ' ')"
done; done
