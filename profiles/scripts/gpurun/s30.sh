#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "small_batch or batch_size_policy" > gpurun_out/s30_tests.log 2>&1; echo tests rc=$?
tail -n 5 gpurun_out/s30_tests.log
timeout 600 python profiles/scripts/small_batch_irreg.py > gpurun_out/s30_small.txt 2> gpurun_out/s30.err
timeout 600 python profiles/scripts/small_batch.py >> gpurun_out/s30_small.txt 2>> gpurun_out/s30.err
cat gpurun_out/s30_small.txt | grep -v "B=  512\|B= 2048"
