python -m pytest tests/test_gpu_apply.py tests/test_gpu_smooth.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t.log
python tools/bench_extra.py --skip-1km --out gpurun_out/extra_r1.json >> gpurun_out/t.log 2>&1
