#!/bin/bash
# round-2 GPU session: final kernel -- GPU test suite, bench line as the driver runs it, launch list + ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest6.txt 2>&1
tail -4 gpurun_out/r2_pytest6.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
tail -c 1200 gpurun_out/r2_bench_c.json; tail -3 gpurun_out/r2_bench_c.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c.csv \
  python bench.py --steps 2 --warmup 3 --reads 1500 --no-cpu-baseline --no-next-rows --no-parity-leg > gpurun_out/r2_launch_bench_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lstm_tc -s 3 -c 1 -o gpurun_out/tc_r2c -f \
  python bench.py --steps 2 --warmup 3 --reads 1500 --no-cpu-baseline --no-next-rows --no-parity-leg > gpurun_out/r2_ncu_bench_c.log 2>&1
ls -la gpurun_out/tc_r2c.ncu-rep
