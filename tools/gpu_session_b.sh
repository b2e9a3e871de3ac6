#!/bin/bash
# two-GPU session: the real peer-memory slab path (flags and NCCL dt sync), then the default bench at N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29520 tools/slab_debug.py 2>&1 | grep -E "rank|identical|Error|error" | grep -v "^W1018\|OMP_NUM" | head -40
for mode in "p2p flags" "p2p nccl"; do
  timeout 300 $TR --master-port 29511 tools/slab_check.py 2048 2048 40 $mode 2>&1 | grep -E "slab_check|Error|error|Traceback" | head -5
done
timeout 300 $TR --master-port 29512 tools/slab_check.py 8192 8192 100 p2p flags 2>&1 | grep -E "slab_check|Error|error|Traceback" | head -5
timeout 300 $TR --master-port 29513 tools/slab_check.py 8192 8192 100 p2p nccl 2>&1 | grep -E "slab_check|Error|error|Traceback" | head -5
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench 2gpu rc=$?"; grep -v "^W1018\|^\[W" gpurun_out/r2_bench_2gpu.err | tail -3
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_2gpu.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "e2e64", d["e2e"]["float64_host"]["value"])
for k,v in d["sub_records"].items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","finite","bounded","identical_to_single_gpu","n_gpus")})
PY
