#!/bin/bash
# s41: ncu --set full of the per-frame-early-termination kernels at full width (passes 3 and 4 of the first decode), with and
# without the three-input-table images
cd $GRAFT_REPO_ROOT
python profiles/scripts/pf_profile.py c1 frame > gpurun_out/s41_pf.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"ib_phase_pf_kernel" -s 6 -c 2 -f -o gpurun_out/prof_pf_tri python profiles/scripts/pf_profile.py c1 frame > gpurun_out/s41_ncu_tri.log 2>&1; echo ncu rc=$?
IBLDPC_NO_PF_TRIPLE=1 ncu --set full --clock-control none --import-source on -k regex:"ib_phase_pf_kernel" -s 6 -c 2 -f -o gpurun_out/prof_pf_notri python profiles/scripts/pf_profile.py c1 frame > gpurun_out/s41_ncu_notri.log 2>&1; echo ncu rc=$?
tail -n 3 gpurun_out/s41_pf.log
