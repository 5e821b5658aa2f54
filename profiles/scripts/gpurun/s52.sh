#!/bin/bash
# s52: smoke + the small-batch / policy tests on the final library (after the launch-planner fix)
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s52_smoke.log 2>&1; echo smoke rc=$?
timeout 80 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "batch_size_policy or timed_bench_geometry" 2>&1 | tail -n 2
