#!/bin/bash
# Development script: A/B of the sub-batch sweep size (bench.py --sweep-rows) on ONE box, interleaved twice.
for i in 1 2; do
  for r in 0 151296 302592 75648; do
    python bench.py --quick --sweep-rows $r --steps 6 --warmup 3 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sweep_rows=$r', round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms')"
  done
done
