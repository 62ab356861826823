#!/bin/bash
# rows_kernel launch-shape sweep (raw rows and fp16 rows): envs per group (log2) x warps per block.  Development tooling.
for lgg in 1 2 0; do for w in 8 6 12 16; do
echo "AT_ROWS_LGG=$lgg AT_ROWS_WARPS=$w AT_ROWS_WARPS_H=$w"
AT_ROWS_LGG=$lgg AT_ROWS_WARPS=$w AT_ROWS_WARPS_H=$w timeout 100 python tools/rows_probe.py 2>&1 | grep -E "at_step: raw|norm16: normalised fp16 rows only" | head -2
done; done
