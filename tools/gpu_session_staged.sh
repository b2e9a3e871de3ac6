#!/bin/bash
mkdir -p gpurun_out
for opt in "" "--up-staged"; do
python bench.py --steps 100 --warmup 10 --no-sub-records --no-cpu-baseline $opt > gpurun_out/r2_bench_staged.json 2> gpurun_out/r2_bench_staged.err; echo "bench [$opt] rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_staged.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
PY
done
