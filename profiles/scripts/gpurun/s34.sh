#!/bin/bash
# s34: compute-sanitizer memcheck over the kernels added this round (small cases)
cd $GRAFT_REPO_ROOT
which compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "batch_size_policy or (tail_pair_variant_all_degrees and n4) or (golden_device_buffers and ib_wlan_T16_match and not True)" > gpurun_out/s34_memcheck.log 2>&1; echo memcheck rc=$?
tail -n 25 gpurun_out/s34_memcheck.log
