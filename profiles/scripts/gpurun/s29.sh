#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s29_tests.log 2>&1; echo tests rc=$?
tail -n 3 gpurun_out/s29_tests.log
python profiles/scripts/small_batch_irreg.py > gpurun_out/s29_small.txt 2> gpurun_out/s29.err
python profiles/scripts/small_batch.py >> gpurun_out/s29_small.txt 2>> gpurun_out/s29.err
echo done
