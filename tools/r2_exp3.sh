#!/bin/bash
# round-2 GPU session 3 (8 GPUs): the multi-GPU tests and the N=8 bench line the driver will run
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/r2_pytest3.txt 2>&1
tail -8 gpurun_out/r2_pytest3.txt
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench n8 rc=$?"
tail -c 2500 gpurun_out/r2_bench_n8.json; tail -5 gpurun_out/r2_bench_n8.err
