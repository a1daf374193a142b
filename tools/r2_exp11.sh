#!/bin/bash
# round-2 GPU session (8 GPUs): the N = 8 bench line again with the collective communicator teardown (dm_reduce_finalize)
mkdir -p gpurun_out
SECONDS=0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err
echo "bench n8 rc=$? in ${SECONDS}s"
tail -c 600 gpurun_out/r2_bench_n8b.json; grep -v "^$" gpurun_out/r2_bench_n8b.err | tail -3
