#!/bin/bash
# what-if runs of conv[1] on a trace + developer build (results deliberately wrong for flags != 0):
# 1 no MMA issue, 2 no global loads, 4 no shared-memory stores, 8 no tcgen05.ld, 16 no global stores, 32 no proxy fence
PBMC_EXTRA_NVCC_FLAGS="-DPBMC_ROW_TRACE -DPBMC_DEV_BUILD" python pbml_mantle_convection_b200/build.py --force > /dev/null 2>&1
for fl in 0 1 2 4 6 38 24 25 39 63; do PBMC_ROW_DBG_FLAGS=$fl python tools/conv1_time.py 2>&1 | tail -1; done
