set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --shard time 2>&1 | tail -2 | tee gpurun_out/bench_c2_n2_time.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c2_n2.json
