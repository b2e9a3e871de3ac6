#!/bin/bash
# four-GPU session: slab surrogate with real NCCL (1 and 3 steps), then the default bench at N = 4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
F='slab_check|slab_surrogate|rows with|Error|error|Traceback'
timeout 300 $TR --master-port 29531 tools/slab_surrogate_check.py 1024 1024 1 6 2>&1 | grep -E "$F" | head -6
timeout 300 $TR --master-port 29532 tools/slab_surrogate_check.py 1024 1024 3 6 2>&1 | grep -E "$F" | head -6
timeout 300 $TR --master-port 29533 tools/slab_surrogate_check.py 2048 2048 3 6 2>&1 | grep -E "$F" | head -6
timeout 600 $TR --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
echo "bench 4gpu rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_4gpu.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "e2e64", d["e2e"]["float64_host"]["value"])
for k,v in d["sub_records"].items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","finite","bounded","identical_to_single_gpu","n_gpus","issue_mode","ms_per_step_by_issue_mode")})
PY
