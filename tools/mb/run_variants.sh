#!/bin/bash
# A/B the number of floors moved from the XU pipe to the FMA pipe (TW_MAGIC_FLOORS = 0, 3, 6)
for K in 0 3 6; do
  cp tools/mb/lib_mf$K.so 3dworld_b200/lib3dworld_b200.so
  echo "== TW_MAGIC_FLOORS=$K"; timeout 100 python -m pytest tests/test_gpu_heightgen.py -m gpu -q -x 2>&1 | tail -1
  timeout 100 python bench.py --kernel-only | cut -c60-175
done
