#!/bin/bash
# Development script: clips per step A/B (bench.py --batch) on ONE box, interleaved twice.
for i in 1 2; do
  for b in 512 768 1024; do
    python bench.py --quick --batch $b --steps 4 --warmup 3 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch=$b', round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms')"
  done
done
