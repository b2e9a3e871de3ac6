#!/bin/bash
# is conv[1] bound by DRAM?  developer build; aliased sources + no flush = inputs L2 resident
PBMC_EXTRA_NVCC_FLAGS="-DPBMC_DEV_BUILD" python pbml_mantle_convection_b200/build.py --force > /dev/null 2>&1
for raw in 0; do
  PBMC_ROW_RAW=$raw python tools/conv1_time.py | tail -1
  PBMC_ROW_RAW=$raw CONV1_NOFLUSH=1 python tools/conv1_time.py | tail -1
  PBMC_ROW_RAW=$raw CONV1_ALIAS=1 CONV1_NOFLUSH=1 python tools/conv1_time.py | tail -1
done
