for lgg in 1 0; do for wh in 0 4 7 8 14 16; do
echo "AT_ROWS_LGG=$lgg AT_ROWS_WARPS_H=$wh"
AT_ROWS_LGG=$lgg AT_ROWS_WARPS_H=$wh timeout 100 python tools/rows_probe.py 2>&1 | grep -E "norm16: normalised fp16 rows only|at_step: raw"
done; done
