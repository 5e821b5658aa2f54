#!/bin/bash
# s51: launch planner fix for 24-warp CTAs: regression test + the tests around the tail-pair variable-node kernels + smoke
cd $GRAFT_REPO_ROOT
timeout 170 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vn_tail_pair or dvbs2_full_size or irregular_vs_oracle or full_size_properties" > gpurun_out/s51_tests.log 2>&1; echo tests rc=$?
tail -n 3 gpurun_out/s51_tests.log
