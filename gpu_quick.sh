python -m pytest tests/test_gpu_build.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/t.log
