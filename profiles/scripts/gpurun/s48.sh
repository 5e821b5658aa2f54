#!/bin/bash
# s48: validation (after the last change, the T32 cooperative kernel) of the final tree of round 2 (library rebuilt from scratch)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s48_tests.log 2>&1; echo tests rc=$?
tail -n 2 gpurun_out/s48_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s48_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/s48_bench.json 2> gpurun_out/s48_bench.err; echo bench rc=$?
