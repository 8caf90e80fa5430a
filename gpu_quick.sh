python -m pytest tests/test_gpu_apply.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t.log
for bps in 2 3 4; do echo "== TMA bps=$bps"; IB_SPMM_VARIANT=2 IB_SPMM_BPS=$bps python tools/spmm_explore.py 512; 
IB_SPMM_BPS=$bps python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | python tools/summ.py "bench bps=$bps"; done >> gpurun_out/t.log 2>&1
