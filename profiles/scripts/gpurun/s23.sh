#!/bin/bash
# s23: three-input-table check-node kernels of degree 7 / 8 (per-class path): parity + A/B
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail_pair_variant_all_degrees or irregular_vs_oracle or dvbs2_full_size or (golden_device_buffers and (nophase or cn_tri6 or notriple))" > gpurun_out/s23_tests.log 2>&1; echo tests rc=$?
B="python bench.py --steps 5 --warmup 3 --no-legs --no-cpu-baseline --no-e2e"
$B --workload dvbs2 > gpurun_out/s23_tri8_dvbs2.json 2>gpurun_out/s23.err; echo rc=$?
IBLDPC_CN_TRI_MAX_DEGREE=6 $B --workload dvbs2 > gpurun_out/s23_tri6_dvbs2.json 2>>gpurun_out/s23.err; echo rc=$?
$B --workload wlan > gpurun_out/s23_phase_wlan.json 2>>gpurun_out/s23.err; echo rc=$?
IBLDPC_NO_PHASE=1 $B --workload wlan > gpurun_out/s23_perclass_tri8_wlan.json 2>>gpurun_out/s23.err; echo rc=$?
IBLDPC_NO_PHASE=1 IBLDPC_CN_TRI_MAX_DEGREE=6 $B --workload wlan > gpurun_out/s23_perclass_tri6_wlan.json 2>>gpurun_out/s23.err; echo rc=$?
tail -3 gpurun_out/s23_tests.log
