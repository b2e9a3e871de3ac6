#!/bin/bash
# developer build: programmatic-dependent-launch mask of the forward's chain (api.cu chain_pdl), whole-step time
PBMC_EXTRA_NVCC_FLAGS="-DPBMC_DEV_BUILD" python pbml_mantle_convection_b200/build.py --force > /dev/null 2>&1
for m in 29 31 29 31; do
  PBMC_CHAIN_PDL=$m python bench.py --steps 100 --warmup 10 --no-sub-records --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chain_pdl=$m ms/step', round(d['ms_per_step'],5), 'resident_loop', round(d['resident_loop']['ms_per_step'],5))
"
done
