python -m pytest tests/test_gpu_apply.py tests/test_gpu_smooth.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t.log
python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | python tools/summ.py "bench" >> gpurun_out/t.log
