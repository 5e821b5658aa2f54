#!/bin/bash
# s50: launch list of one DVB-S2-like decode with the final kernels (which degree class costs what)
cd $GRAFT_REPO_ROOT
python bench.py --workload dvbs2 --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s50_plain.json 2> gpurun_out/s50.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/s50_launches_dvbs2.csv python bench.py --workload dvbs2 --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s50_ncu.log 2>&1; echo launches rc=$?
