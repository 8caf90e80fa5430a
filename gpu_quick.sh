set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests/test_gpu_build.py -x -q -m gpu 2>&1 | tail -30
