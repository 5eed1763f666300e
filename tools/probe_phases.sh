#!/bin/bash
# per-phase kernel time of round K (default: the last round of a 6-read edge) on the cfg3 shape; differences between successive lines = cost of a phase
K=${1:-5}
for p in 1 2 3 4 5 0; do echo "round $K stop-after-phase $p (0 = full edge)"; HGPU_PROBE=$p HGPU_PROBE_ROUND=$K python tools/fill_probe.py 40000 2>&1 | grep cfg3; done
