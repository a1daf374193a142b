#!/bin/bash
# ncu --set full of two kernel variants (tools/variants/*.so), 1 M windows through dm_forward_windows
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from deepmod_b200 import capi, checkpoint
with np.load("tests/golden/model_conmodC_P100.npz") as z:
    model = checkpoint.Model.from_dict({k: z[k] for k in z.files})
with np.load("tests/golden/windows_conmodC_P100.npz") as z:
    X = z["X"]
ctx = capi.Context(model, 0, capi.F16)
big = np.tile(X, (512, 1, 1))
for _ in range(3):
    ctx.forward_windows(big)
PY
for v in ${VARIANTS:-a_both b_fmaf}; do
  DEEPMOD_B200_LIB=tools/variants/$v.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lstm_tc -s 2 -c 1 \
    -o gpurun_out/tc_r2_$v -f python /tmp/one.py > gpurun_out/r2_ncu_$v.log 2>&1
  tail -2 gpurun_out/r2_ncu_$v.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
