#!/bin/bash
# developer build: per-level CTA budgets of the persistent trunk kernels (PBMC_BUDGETS), whole-step time
PBMC_EXTRA_NVCC_FLAGS="-DPBMC_DEV_BUILD" python pbml_mantle_convection_b200/build.py --force > /dev/null 2>&1
for b in "" "100,32,8,4,2,2" "96,32,8,4,4,2" "104,28,8,4,2,2" "100,30,8,4,3,3" "92,36,10,4,3,3" "112,24,6,3,2,1"; do
  PBMC_BUDGETS=$b python bench.py --steps 100 --warmup 10 --no-sub-records --no-cpu-baseline > /tmp/b.json 2> /tmp/b.err
  python - <<PY
import json
d=json.load(open("/tmp/b.json"))
print("budgets [$b] ms/step", round(d["ms_per_step"],5))
PY
done
