#!/bin/bash
# s39: final state of round 2: full GPU suite, smoke, default bench, reference arm, ncu --set full + launch list (each after its own plain run)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s39_tests.log 2>&1; echo tests rc=$?
tail -n 2 gpurun_out/s39_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s39_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/s39_bench.json 2> gpurun_out/s39_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s39_ref.json 2>> gpurun_out/s39_bench.err; echo ref rc=$?
ncu --set full --clock-control none --import-source on -k regex:"ib_(cn|vn)[0-9]*_n4" -s 12 -c 2 -f -o gpurun_out/prof_c1_r02e python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s39_ncu_c1.log 2>&1; echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s39_launches_c1.csv python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s39_ncu_launch.log 2>&1; echo launches rc=$?
