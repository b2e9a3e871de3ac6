# end-of-round GPU check (one short gpurun call): learned-network timing + the network-level GPU tests
timeout 50 python tools/learned_time.py > gpurun_out/learned_time2.log 2>&1; tail -3 gpurun_out/learned_time2.log
timeout 100 python -m pytest tests/test_gpu_net.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_net_final.log; cat gpurun_out/pytest_net_final.log
