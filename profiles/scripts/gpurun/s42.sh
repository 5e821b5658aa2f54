#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "fused or per_frame or policy or small_batch or t32 or T32 or (golden_device_buffers and n4_nocoop) or irregular_vs_oracle" > gpurun_out/s42_tests.log 2>&1; echo tests rc=$?
tail -n 2 gpurun_out/s42_tests.log
B="python bench.py --steps 5 --warmup 3 --no-legs --no-cpu-baseline --no-e2e"
$B --workload wlan > gpurun_out/s42_wlan.json 2>gpurun_out/s42.err; echo rc=$?
$B --workload wlan1944 > gpurun_out/s42_wlan1944.json 2>>gpurun_out/s42.err; echo rc=$?
$B --workload wlan --card 32 > gpurun_out/s42_wlan_T32.json 2>>gpurun_out/s42.err; echo rc=$?
python - <<'P'
import json
for f in ("wlan","wlan1944","wlan_T32"):
    d=json.load(open(f"gpurun_out/s42_{f}.json")); r=d["roofline"]
    print(f, round(d["value"],3), round(d["ms_per_step"],2), r.get("cn_avg_ms"), r.get("vn_avg_ms"), d["parity_sample"]["equal"])
P
