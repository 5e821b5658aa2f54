#!/bin/bash
# s45: A/B of the dp4a address arithmetic in the tail-pair variable-node kernels now that degree 8 runs 24 warps per SM
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 5 --warmup 3 --no-legs --no-cpu-baseline --no-e2e --workload dvbs2"
$B > gpurun_out/s45_dvbs2_default.json 2>gpurun_out/s45.err; echo rc=$?
IBLDPC_LIB=$GRAFT_REPO_ROOT/informationbottleneckdecodingldpc_b200/libibldpc_dp4a_all.so $B > gpurun_out/s45_dvbs2_dp4a_all.json 2>>gpurun_out/s45.err; echo rc=$?
python - <<'P'
import json
for f in ("default","dp4a_all"):
    d=json.load(open(f"gpurun_out/s45_dvbs2_{f}.json")); r=d["roofline"]
    print(f, round(d["value"],3), round(d["ms_per_step"],2), r.get("cn_avg_ms"), r.get("vn_avg_ms"), d["parity_sample"]["equal"])
P
