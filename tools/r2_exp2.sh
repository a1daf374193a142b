#!/bin/bash
# round-2 GPU session 2: the whole GPU test suite, the new bench line, launch list + one full ncu capture of k_lstm_tc
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest2.txt 2>&1
tail -15 gpurun_out/r2_pytest2.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_a.csv \
  python bench.py --steps 2 --warmup 3 --reads 1500 --no-cpu-baseline --no-next-rows --no-parity-leg > gpurun_out/r2_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lstm_tc -s 3 -c 1 -o gpurun_out/tc_r2a -f \
  python bench.py --steps 2 --warmup 3 --reads 1500 --no-cpu-baseline --no-next-rows --no-parity-leg > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out | tail -8
