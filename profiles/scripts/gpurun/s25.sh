#!/bin/bash
# s25: float64 BP in the likelihood-ratio domain: LLR parity tests + BP leg A/B
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "llr or minsum or bp or layered or driver" > gpurun_out/s25_tests.log 2>&1; echo tests rc=$?
tail -n 3 gpurun_out/s25_tests.log
cat > /tmp/bpleg.py <<'P'
import json, sys
sys.path.insert(0, ".")
import bench
import informationbottleneckdecodingldpc_b200 as pkg
import torch
torch.cuda.set_device(0)
print(json.dumps({"bp_f64": bench.leg_llr(pkg, "bp", 3, 0), "minsum_f64": bench.leg_llr(pkg, "minsum", 3, 0)}))
P
python /tmp/bpleg.py > gpurun_out/s25_bp_ratio.json 2>gpurun_out/s25.err; echo rc=$?
IBLDPC_BP_LOGDOMAIN=1 python /tmp/bpleg.py > gpurun_out/s25_bp_log.json 2>>gpurun_out/s25.err; echo rc=$?
