#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vn_tail_pair or dvbs2_full_size or irregular_vs_oracle or (golden_device_buffers and (nophase or vpair3))" > gpurun_out/s43_tests.log 2>&1; echo tests rc=$?
tail -n 2 gpurun_out/s43_tests.log
B="python bench.py --steps 5 --warmup 3 --no-legs --no-cpu-baseline --no-e2e --workload dvbs2"
$B > gpurun_out/s43_dvbs2_768.json 2>gpurun_out/s43.err; echo rc=$?
IBLDPC_VN_PAIR_THREADS=256 $B > gpurun_out/s43_dvbs2_256.json 2>>gpurun_out/s43.err; echo rc=$?
python - <<'P'
import json
for f in ("768","256"):
    d=json.load(open(f"gpurun_out/s43_dvbs2_{f}.json")); r=d["roofline"]
    print(f, round(d["value"],3), round(d["ms_per_step"],2), r.get("cn_avg_ms"), r.get("vn_avg_ms"), d["parity_sample"]["equal"])
P
