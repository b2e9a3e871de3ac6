#!/bin/bash
# trunk kernel with 6 groups of worker warps (832 threads, 72 registers) against the shipped 5 (704 threads, 80 registers)
for n in 6 5; do
  PBMC_EXTRA_NVCC_FLAGS="-DPBMC_CT_SETS=$n" python pbml_mantle_convection_b200/build.py --force > /dev/null 2>&1
  echo "== CT_SETS=$n"
  python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "trunk" 2>&1 | tail -1
  python tools/one_trunk.py 512 512 4 1 0 mux_f16x2 threads
  python tools/one_trunk.py 512 512 4 1 96 mux_f16x2 threads
  python bench.py --steps 100 --warmup 10 --no-sub-records --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],5), 'resident_loop', round(d['resident_loop']['ms_per_step'],5), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
"
done
