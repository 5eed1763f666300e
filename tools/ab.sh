#!/bin/bash
# A/B: run the bench at a reduced edge count for each library given; prints GCUPS per variant
for lib in "$@"; do
  HASLR_B200_LIB=$lib timeout 300 python bench.py --edges 50000 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'gcups', round(d['roofline']['gcups'],1), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
