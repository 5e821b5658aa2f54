#!/bin/bash
# s27: small batches -- cooperative whole-decode kernel (default) vs per-phase launches
cd $GRAFT_REPO_ROOT
for mode in default nocoop_phase nocoop; do
  case $mode in
    default) envs="";;
    nocoop_phase) envs="IBLDPC_COOP_MAX_B=0 IBLDPC_PHASE=1";;
    nocoop) envs="IBLDPC_COOP_MAX_B=0";;
  esac
  echo "== $mode" >> gpurun_out/s27_small.txt
  env $envs python profiles/scripts/small_batch_irreg.py >> gpurun_out/s27_small.txt 2>> gpurun_out/s27.err
  env $envs python profiles/scripts/small_batch.py >> gpurun_out/s27_small.txt 2>> gpurun_out/s27.err
done
echo done
