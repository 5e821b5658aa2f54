#!/bin/bash
cd $GRAFT_REPO_ROOT
IBLDPC_COOP_MAX_B=8192 timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "small_batch or batch_size_policy" > gpurun_out/s31_tests.log 2>&1; echo tests rc=$?
tail -n 5 gpurun_out/s31_tests.log
echo "== coop phase up to 8192" > gpurun_out/s31_mid.txt
IBLDPC_COOP_MAX_B=8192 timeout 900 python profiles/scripts/mid_batch.py >> gpurun_out/s31_mid.txt 2> gpurun_out/s31.err
echo "== default (coop phase <= 256, fused phase kernels / per-class above)" >> gpurun_out/s31_mid.txt
timeout 900 python profiles/scripts/mid_batch.py >> gpurun_out/s31_mid.txt 2>> gpurun_out/s31.err
cat gpurun_out/s31_mid.txt
