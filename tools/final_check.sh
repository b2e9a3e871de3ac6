timeout 60 python tools/learned_time.py > gpurun_out/learned_time.log 2>&1; tail -2 gpurun_out/learned_time.log
timeout 150 python -m pytest tests/test_gpu_net.py -m gpu -x -q -k "learned or module_level or fallback" 2>&1 | tail -4 > gpurun_out/pytest_learned.log; cat gpurun_out/pytest_learned.log
timeout 120 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_final_r1.json 2> gpurun_out/bench_final_r1.err; wc -l gpurun_out/bench_final_r1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_final_r1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])"
