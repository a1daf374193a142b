#!/bin/bash
# round-2 final validation: GPU test suite, smoke, the bench line as the driver runs it, and the bf16-operand value leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_final.txt 2>&1; tail -3 gpurun_out/r2_pytest_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.txt 2>&1; tail -2 gpurun_out/r2_smoke_final.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_final.json')); print(d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['next_rows']['cli'].get('value'), d['cpu_baseline']['value'])"
timeout 300 python bench.py --precision bf16 --steps 8 --warmup 3 --no-cpu-baseline --no-next-rows --no-parity-leg > gpurun_out/r2_bench_final_bf16.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2_bench_final_bf16.json')); print('bf16', d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac'])"
