#!/bin/bash
# s24: funnel-shift nibble push: full GPU suite + benches of the three workloads
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s24_tests.log 2>&1; echo tests rc=$?
B="python bench.py --steps 5 --warmup 3 --no-legs --no-cpu-baseline --no-e2e"
$B > gpurun_out/s24_c1.json 2>gpurun_out/s24.err; echo rc=$?
$B --workload dvbs2 > gpurun_out/s24_dvbs2.json 2>>gpurun_out/s24.err; echo rc=$?
$B --workload wlan > gpurun_out/s24_wlan.json 2>>gpurun_out/s24.err; echo rc=$?
tail -n 3 gpurun_out/s24_tests.log
