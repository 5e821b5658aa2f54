#!/bin/bash
# s38: why is the fused per-phase kernel of the (3,6) set slower than the per-class kernel?  ncu of both check-node kernels
cd $GRAFT_REPO_ROOT
IBLDPC_PHASE=1 python bench.py --steps 3 --warmup 3 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s38_c1_phase.json 2> gpurun_out/s38.err; echo rc=$?
IBLDPC_PHASE=1 ncu --set full --clock-control none --import-source on -k regex:"ib_phase_kernel" -s 12 -c 2 -f -o gpurun_out/prof_c1_phase python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline --no-e2e > gpurun_out/s38_ncu.log 2>&1; echo ncu rc=$?
