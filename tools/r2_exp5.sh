#!/bin/bash
mkdir -p gpurun_out
SWEEP_REPS=2 SWEEP_PREC=3 timeout 600 python tools/tc_sweep.py run > gpurun_out/r2_sweep2.txt 2>&1
cat gpurun_out/r2_sweep2.txt
