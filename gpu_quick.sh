python -m pytest tests/test_gpu_apply.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t.log
python bench.py --steps 300 --warmup 10 2>&1 | tail -1 > gpurun_out/bench_r1.json
python tools/summ.py bench < gpurun_out/bench_r1.json >> gpurun_out/t.log
