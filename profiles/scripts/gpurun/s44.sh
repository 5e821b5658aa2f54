#!/bin/bash
# s44: final capture after the last kernel change (768-thread d_v = 8-9 kernels): default bench, ncu --set full + launch list of C1, DVB-S2 leg
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/s44_bench.json 2> gpurun_out/s44_bench.err; echo bench rc=$?
ncu --set full --clock-control none --import-source on -k regex:"ib_(cn|vn)[0-9]*_n4" -s 12 -c 2 -f -o gpurun_out/prof_c1_r02f python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s44_ncu_c1.log 2>&1; echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s44_launches_c1.csv python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s44_ncu_launch.log 2>&1; echo launches rc=$?
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_device_buffers and n4 and not generic" 2>&1 | tail -n 2
