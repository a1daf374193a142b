#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_predetail.py tests/test_gpu_cli.py tests/test_gpu_errors.py -m gpu -q -x > gpurun_out/r2_pytest10.txt 2>&1; tail -3 gpurun_out/r2_pytest10.txt
timeout 600 python - > gpurun_out/r2_cli2.jsonl 2> gpurun_out/r2_cli2.err <<'PY'
import importlib.util, json
spec = importlib.util.spec_from_file_location("dm_bench", "bench.py"); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
from deepmod_b200 import capi, checkpoint
ctx = capi.Context(checkpoint.Model.from_dict(b.load_weights()), device=0, precision=capi.F16)
ctx.set_genome([b.GENOME_LEN], "C")
for rep in range(3):
    print(json.dumps(b.cli_leg(ctx, ctx.synth_spec(**b.SYNTH), 0)))
PY
cut -c1-330 gpurun_out/r2_cli2.jsonl; tail -3 gpurun_out/r2_cli2.err
