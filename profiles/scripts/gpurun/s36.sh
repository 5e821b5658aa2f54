#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "per_frame or fused_phase or small_batch or policy" > gpurun_out/s36_tests.log 2>&1; echo tests rc=$?
tail -n 4 gpurun_out/s36_tests.log
cat > /tmp/etleg.py <<'P'
import json, sys
sys.path.insert(0, ".")
import bench
import informationbottleneckdecodingldpc_b200 as pkg
import torch
torch.cuda.set_device(0)
r = bench.leg_early_termination(pkg, 3, 0)
print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ("value", "ms_per_step", "mean_i_num", "gpu_launches_per_step", "equal")}) for k, v in r.items()}))
P
python /tmp/etleg.py > gpurun_out/s36_et_tri.json 2> gpurun_out/s36.err; echo rc=$?
IBLDPC_NO_PF_TRIPLE=1 python /tmp/etleg.py > gpurun_out/s36_et_notri.json 2>> gpurun_out/s36.err; echo rc=$?
cat gpurun_out/s36_et_tri.json gpurun_out/s36_et_notri.json
