#!/bin/bash
# two-GPU session: bisect the slab-surrogate mismatch seen with real NCCL at 8 ranks; bench at N = 2 with the final code
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='slab_check|slab_surrogate|rows with|Error|error|Traceback'
python tools/slab_surrogate_threads.py 2048 2048 8 6 1 2>&1 | grep -E "$F" | head -4
python tools/slab_surrogate_threads.py 512 512 2 4 1 2>&1 | grep -E "$F" | head -4
timeout 300 $TR --master-port 29531 tools/slab_surrogate_check.py 512 512 1 4 2>&1 | grep -E "$F" | head -6
timeout 300 $TR --master-port 29532 tools/slab_surrogate_check.py 2048 2048 1 6 2>&1 | grep -E "$F" | head -6
timeout 300 $TR --master-port 29512 tools/slab_check.py 8192 8192 100 p2p flags 0 2>&1 | grep -E "$F" | head -5
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench 2gpu rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_2gpu.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "e2e64", d["e2e"]["float64_host"]["value"])
for k,v in d["sub_records"].items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","finite","bounded","identical_to_single_gpu","n_gpus","issue_mode","ms_per_step_by_issue_mode")})
PY
