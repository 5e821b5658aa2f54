#!/bin/bash
cd $GRAFT_REPO_ROOT
for mode in nocoop_phase nocoop; do
  case $mode in
    nocoop_phase) envs="IBLDPC_COOP_MAX_B=0 IBLDPC_PHASE=1";;
    nocoop) envs="IBLDPC_COOP_MAX_B=0 IBLDPC_NO_PHASE=1";;
  esac
  echo "== $mode" >> gpurun_out/s28_mid.txt
  env $envs python profiles/scripts/mid_batch.py >> gpurun_out/s28_mid.txt 2>> gpurun_out/s28.err
done
echo done
