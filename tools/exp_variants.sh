#!/bin/bash
# experiment harness: time bench.py with alternative builds of the library (alphatank_b200/lib_<TAG>.so)
cp alphatank_b200/libalphatank_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp alphatank_b200/lib_$v.so alphatank_b200/libalphatank_b200.so
  touch alphatank_b200/libalphatank_b200.so
  echo -n "$v: "; python bench.py --steps 200 --warmup 10 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('%.4f ms/step' % d['ms_per_step'])"
done
cp /tmp/lib_orig.so alphatank_b200/libalphatank_b200.so
