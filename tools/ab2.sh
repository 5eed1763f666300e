#!/bin/bash
# occupancy sweep: resident warps per device (148 SMs x warps/SM)
for w in "$@"; do
  HGPU_MAX_WARPS=$w timeout 300 python bench.py --edges 50000 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warps $w', 'gcups', round(d['roofline']['gcups'],1), 'value', round(d['value'],1))"
done
