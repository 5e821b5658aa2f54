#!/bin/bash
# s26: default bench (all legs), then ncu --set full of the two dominant kernels and the launch list (each after its own plain run)
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/s26_bench.json 2> gpurun_out/s26_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s26_ref.json 2>> gpurun_out/s26_bench.err; echo ref rc=$?
ncu --set full --clock-control none --import-source on -k regex:"ib_(cn|vn)[0-9]*_n4" -s 12 -c 2 -f -o gpurun_out/prof_c1_r02d python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s26_ncu_c1.log 2>&1; echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s26_launches_c1.csv python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s26_ncu_launch.log 2>&1; echo launches rc=$?
