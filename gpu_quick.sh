python -m pytest tests/test_gpu_hntr.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/t.log
python tools/bench_extra.py --only hntr >> gpurun_out/t.log 2>&1
