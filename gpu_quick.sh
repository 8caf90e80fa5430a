python -m pytest tests/test_gpu_merge.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 >> gpurun_out/t.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference >> gpurun_out/t.log 2>&1
