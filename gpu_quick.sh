timeout 300 python -m pytest tests/test_gpu_apply.py tests/test_gpu_smooth.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t.log
timeout 120 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | python tools/summ.py "bench fused" >> gpurun_out/t.log
IB_SPMM_NOFUSE=1 timeout 120 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | python tools/summ.py "bench nofuse" >> gpurun_out/t.log
