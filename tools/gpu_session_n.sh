#!/bin/bash
# N-GPU session (N = $1): the flag-synchronised slab stencil, the slab-decomposed surrogate, then the default bench at N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='slab_check|slab_surrogate_check|Error|error|Traceback'
timeout 300 $TR --master-port 29511 tools/slab_check.py 8192 8192 100 p2p flags 0 2>&1 | grep -E "$F" | head -5
timeout 300 $TR --master-port 29512 tools/slab_check.py 8192 8192 100 p2p flags 20 2>&1 | grep -E "$F" | head -5
timeout 300 $TR --master-port 29513 tools/slab_check.py 8192 8192 100 p2p nccl 20 2>&1 | grep -E "$F" | head -5
timeout 300 $TR --master-port 29515 tools/slab_surrogate_check.py 2048 2048 3 6 2>&1 | grep -E "$F" | head -5
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench ${N}gpu rc=$?"; grep -v "^W1018\|^\[W\|OMP_NUM\|^\*\*\*" gpurun_out/r2_bench_${N}gpu.err | tail -3
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "e2e64", d["e2e"]["float64_host"]["value"])
for k,v in d["sub_records"].items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","finite","bounded","identical_to_single_gpu","n_gpus","issue_mode","ms_per_step_by_issue_mode")})
PY
