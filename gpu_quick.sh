set -x
python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -3
