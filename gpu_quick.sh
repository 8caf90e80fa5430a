python -m pytest tests/test_gpu_apply.py -x -q -m gpu 2>&1 | tail -2 > gpurun_out/t.log
python tools/bench_extra.py --skip-1km --out gpurun_out/extra_r1.json >> gpurun_out/t.log 2>&1
