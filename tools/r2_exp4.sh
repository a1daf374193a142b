#!/bin/bash
# round-2 GPU session 4 (2 GPUs): teardown order of the NCCL communicators (the N = 8 run hung for 600 s at exit)
mkdir -p gpurun_out
date +%s > gpurun_out/r2_t0.txt
timeout 300 python -m pytest tests/test_gpu_cli.py -m gpu -q -x -k two_ranks > gpurun_out/r2_pytest4.txt 2>&1
tail -3 gpurun_out/r2_pytest4.txt
SECONDS=0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$? in ${SECONDS}s"
tail -c 1500 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
