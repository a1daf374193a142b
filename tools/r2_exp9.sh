#!/bin/bash
# round-2 GPU session: new edge-case tests, sustained A/B of three builds, configs[3] at its stated size, CLI with two loader threads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_errors.py -m gpu -q -x > gpurun_out/r2_pytest7.txt 2>&1; tail -5 gpurun_out/r2_pytest7.txt
for v in a_ship b_unitmap c_ld1 a_ship; do
  DEEPMOD_B200_LIB=tools/variants/$v.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-next-rows --no-parity-leg 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(d['roofline']['kernel_ms'],2))"
done | tee gpurun_out/r2_sustained2.txt
timeout 600 python - > gpurun_out/r2_config3_full.json 2> gpurun_out/r2_config3_full.err <<'PY'
import importlib.util, json, os, sys
spec = importlib.util.spec_from_file_location("dm_bench", "bench.py"); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
print(json.dumps(b.mixed_length_leg(0, 100000)))
import torch
from deepmod_b200 import capi, checkpoint
ctx = capi.Context(checkpoint.Model.from_dict(b.load_weights()), device=0, precision=capi.F16)
ctx.set_genome([b.GENOME_LEN], "C")
for rep in range(2):
    print(json.dumps(b.cli_leg(ctx, ctx.synth_spec(**b.SYNTH), 0)))
PY
cat gpurun_out/r2_config3_full.json | cut -c1-700; tail -3 gpurun_out/r2_config3_full.err
