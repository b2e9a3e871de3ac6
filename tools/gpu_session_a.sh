#!/bin/bash
# one-GPU session: trunk loader A/B (TMA bulk copies vs thread loads), tests, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "trunk" 2>&1 | tail -3
for ld in bulk threads; do
  python tools/one_trunk.py 512 512 4 1 0 mux_f16x2 $ld
  python tools/one_trunk.py 512 512 4 1 108 mux_f16x2 $ld
done
for t in auto auto_thread_loader per_layer; do
  python bench.py --steps 100 --warmup 10 --no-sub-records --no-cpu-baseline --trunk $t > gpurun_out/r2_bench_trunk_$t.json 2> gpurun_out/r2_bench_trunk_$t.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_trunk_$t.json"))
print("$t", "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
PY
done
python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" > gpurun_out/r2_t6.log; tail -3 gpurun_out/r2_t6.log
