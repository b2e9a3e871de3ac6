#!/bin/bash
# two-GPU session: where do the ~45 us per slab step go?  graph length, eager launches, symmetric-memory probe
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='slab_check|probe|rank [01]\]|Error|error|Traceback'
timeout 200 $TR --master-port 29520 tools/symm_probe.py 4096 8192 2>&1 | grep -E "$F" | head
timeout 200 $TR --master-port 29521 tools/symm_probe.py 1024 2048 2>&1 | grep -E "$F" | head
for gs in 2 20 0; do
  timeout 300 $TR --master-port 29511 tools/slab_check.py 2048 2048 80 p2p flags $gs 2>&1 | grep -E "$F" | head -5
  timeout 300 $TR --master-port 29512 tools/slab_check.py 8192 8192 80 p2p flags $gs 2>&1 | grep -E "$F" | head -5
done
timeout 300 $TR --master-port 29513 tools/slab_check.py 8192 8192 80 p2p nccl 20 2>&1 | grep -E "$F" | head -5
