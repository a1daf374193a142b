#!/bin/bash
# round-2 GPU experiment 1: TMEM-load and MUFU micro-benchmarks, kernel variants back to back, attribution runs
mkdir -p gpurun_out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3"
$NV -o /tmp/ldtm_bench tools/micro/ldtm_bench.cu && timeout 120 /tmp/ldtm_bench > gpurun_out/r2_ldtm.txt 2>&1
$NV -o /tmp/mufu_bench tools/micro/mufu_bench.cu && timeout 120 /tmp/mufu_bench > gpurun_out/r2_mufu.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
SWEEP_REPS=2 timeout 900 python tools/tc_sweep.py run > gpurun_out/r2_sweep1.txt 2>&1
timeout 600 python tools/quick_gpu_check.py > gpurun_out/r2_quick1.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_cli.py > gpurun_out/r2_pytest1.txt 2>&1
tail -5 gpurun_out/r2_pytest1.txt
cat gpurun_out/r2_ldtm.txt gpurun_out/r2_sweep1.txt
