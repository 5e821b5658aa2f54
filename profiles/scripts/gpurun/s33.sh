#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "small_batch or batch_size_policy or host_pipeline or golden_host" > gpurun_out/s33_tests.log 2>&1; echo tests rc=$?
tail -n 4 gpurun_out/s33_tests.log
timeout 600 python profiles/scripts/small_batch_irreg.py 2>&1 | grep dvbs2
