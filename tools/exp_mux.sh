# developer what-if runs of conv_mux_kernel (needs a -DPBMC_ROW_TRACE build; results of flagged runs are WRONG by design)
for fl in 0 4 8 16 2 12 28 30; do echo "== PBMC_MUX_DBG_FLAGS=$fl"; PBMC_MUX_DBG_FLAGS=$fl python tools/kbench.py mux_f16x2 2>&1 | grep -E "B1 512x512|gn\+gelu +B32"; done
