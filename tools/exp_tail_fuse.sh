#!/bin/bash
# developer build: the caller's loop (10 steps per graph) with and without the input build fused into the stencil launch
PBMC_EXTRA_NVCC_FLAGS="-DPBMC_DEV_BUILD" python pbml_mantle_convection_b200/build.py --force > /dev/null 2>&1
for f in 1 0 1 0; do
  PBMC_TAIL_FUSE=$f python bench.py --steps 20 --warmup 5 --no-sub-records --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('fuse=$f ms/step', round(d['ms_per_step'],5), 'resident_loop', round(d['resident_loop']['ms_per_step'],5))
"
done
