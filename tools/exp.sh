for dx in 16 0 128; do echo "== dbg_dx=$dx"; PBMC_ROW_DBG_DX=$dx python tools/kbench.py row_f16x2 2>&1 | grep -E "B32|B1 512x512"; done
