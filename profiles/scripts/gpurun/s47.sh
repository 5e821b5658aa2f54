#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "t32 or T32" > gpurun_out/s47_tests.log 2>&1; echo tests rc=$?
tail -n 4 gpurun_out/s47_tests.log
cat > /tmp/t32small.py <<'P'
import json, sys
sys.path.insert(0, ".")
import bench
import informationbottleneckdecodingldpc_b200 as pkg
import torch
torch.cuda.set_device(0)
wl = bench.workload("wlan")
for B in (2, 100, 512, 2000, 4096):
    t, tb, quanti, decodi = bench.build_ib(pkg, wl, B, 0, 32)
    ch = quanti.quantize_direct_OpenCL(t.n_var, B)
    ms, out = bench.time_steps(lambda: decodi.decode_OpenCL(ch, buffer_in=True, return_buffer=True), 20, 3, min_warm_s=0.2)
    par = bench.parity_sample_ib(decodi, ch, out, t, tb, wl, 32)
    print(json.dumps({"B": B, "ms_per_decode": round(ms / 20, 4), "launches": decodi.info()[1], "parity": par["equal"]}), flush=True)
P
echo "== default" > gpurun_out/s47_t32_small.txt; python /tmp/t32small.py >> gpurun_out/s47_t32_small.txt 2> gpurun_out/s47.err
echo "== IBLDPC_T32_COOP_MAX_B=0 (one launch per phase)" >> gpurun_out/s47_t32_small.txt; IBLDPC_T32_COOP_MAX_B=0 python /tmp/t32small.py >> gpurun_out/s47_t32_small.txt 2>> gpurun_out/s47.err
echo "== IBLDPC_T32_COOP_MAX_B=8192" >> gpurun_out/s47_t32_small.txt; IBLDPC_T32_COOP_MAX_B=8192 python /tmp/t32small.py >> gpurun_out/s47_t32_small.txt 2>> gpurun_out/s47.err
cat gpurun_out/s47_t32_small.txt; tail -n 3 gpurun_out/s47.err
