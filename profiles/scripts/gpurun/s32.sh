#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s32_tests.log 2>&1; echo tests rc=$?
tail -n 3 gpurun_out/s32_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s32_smoke.log 2>&1; echo smoke rc=$?
timeout 600 python profiles/scripts/small_batch_irreg.py > gpurun_out/s32_small.txt 2> gpurun_out/s32.err
timeout 600 python profiles/scripts/small_batch.py >> gpurun_out/s32_small.txt 2>> gpurun_out/s32.err
