#!/bin/bash
# sustained (power-capped) throughput of kernel variants: the bench's value leg, back to back on one box
mkdir -p gpurun_out
for rep in 1 2; do for v in a_ship b_gate16; do
  DEEPMOD_B200_LIB=tools/variants/$v.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-next-rows --no-parity-leg 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(d['roofline']['kernel_ms'],2))"
done; done | tee gpurun_out/r2_sustained.txt
