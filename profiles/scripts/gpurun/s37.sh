#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "t32 or T32" > gpurun_out/s37_tests.log 2>&1; echo tests rc=$?
tail -n 4 gpurun_out/s37_tests.log
cat > /tmp/t32leg.py <<'P'
import json, sys
sys.path.insert(0, ".")
import bench
import informationbottleneckdecodingldpc_b200 as pkg
import torch
torch.cuda.set_device(0)
r = bench.leg_ib(pkg, "wlan", 3, 0, T_=32)
print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "gpu_launches_per_step", "parity_sample")}))
wl = bench.workload("wlan")
t, tb, quanti, decodi = bench.build_ib(pkg, wl, 2000, 0, 32)
ch = quanti.quantize_direct_OpenCL(t.n_var, 2000)
ms, out = bench.time_steps(lambda: decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True), 20, 3, min_warm_s=0.2)
print(json.dumps({"B2000_ms_per_decode": ms / 20, "launches": decodi.info()[1]}))
P
python /tmp/t32leg.py > gpurun_out/s37_t32_fused.json 2> gpurun_out/s37.err; echo rc=$?
IBLDPC_T32_NO_PHASE=1 python /tmp/t32leg.py > gpurun_out/s37_t32_perclass.json 2>> gpurun_out/s37.err; echo rc=$?
cat gpurun_out/s37_t32_fused.json gpurun_out/s37_t32_perclass.json
