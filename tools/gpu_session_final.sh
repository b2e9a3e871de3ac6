#!/bin/bash
# one-GPU validation of the tree as committed: every -m gpu test, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" > gpurun_out/r2_final_tests.log; tail -3 gpurun_out/r2_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference_arm.json 2> gpurun_out/r2_final_reference_arm.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_final_bench_1gpu.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "e2e64", d["e2e"]["float64_host"]["value"], "cpu", d["cpu_baseline"]["value"])
print("roofline", d["roofline"]); print("clocks", d["clocks"], "launches", d["gpu_launches"])
r=json.load(open("gpurun_out/r2_final_reference_arm.json")); print("reference arm", r["value"], r["unit"], r["cpu_baseline"])
PY
