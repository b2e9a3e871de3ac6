#!/bin/bash
# one-GPU session for conv[1] work: parity tests of the conv operators and the network, conv[1] alone, the step
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py -m gpu -q -x 2>&1 | tail -3
python tools/conv1_time.py 2>&1 | tail -1
python bench.py --steps 100 --warmup 10 --no-sub-records --no-cpu-baseline > gpurun_out/r2_bench_conv1.json 2> gpurun_out/r2_bench_conv1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_conv1.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
PY
