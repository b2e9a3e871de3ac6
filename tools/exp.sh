for fl in 0 1 2 4 6 8 16 24 32 7 31 63; do echo "== flags=$fl"; PBMC_ROW_DBG_FLAGS=$fl python tools/kbench.py row_f16x2 2>&1 | grep -E "plain +B32|plain +B1 512x512|gelu +B32"; done
