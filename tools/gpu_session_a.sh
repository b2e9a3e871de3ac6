#!/bin/bash
# one-GPU session: tests, A/B of the trunk modes, stand-alone trunk timings, learned-network timing, ncu launch list + one full capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" > gpurun_out/r2_t4.log; tail -4 gpurun_out/r2_t4.log
python tools/one_trunk.py 512 512 4 1 0
python tools/one_trunk.py 512 512 4 1 108
python tools/learned_time.py 2>&1 | tail -4
for t in auto per_layer; do
  python bench.py --steps 100 --warmup 10 --no-sub-records --no-cpu-baseline --trunk $t > gpurun_out/r2_bench_trunk_$t.json 2> gpurun_out/r2_bench_trunk_$t.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_trunk_$t.json"))
print("$t", "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:round(v["ms"]*1e3,1) for k,v in d["kernels"].items()})
PY
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r2_plain_short.json 2> gpurun_out/r2_plain_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:(avgpool2|bicubic|build_input|conv_|head_kernel|stencil_march|uvmax|pack_)" -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/one_trunk.py 512 512 4 1 0 > gpurun_out/r2_one_trunk_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_trunk -s 2 -c 1 -o gpurun_out/r2_trunk_b python tools/one_trunk.py 512 512 4 1 0 > gpurun_out/r2_ncu_trunk.log 2>&1
echo "ncu trunk rc=$?"
