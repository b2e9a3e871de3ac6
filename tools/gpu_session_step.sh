#!/bin/bash
# one-GPU session after a change of the rollout chain: every -m gpu test, then the step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 100 --warmup 10 --no-sub-records --no-cpu-baseline > gpurun_out/r2_bench_step.json 2> gpurun_out/r2_bench_step.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_step.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
PY
