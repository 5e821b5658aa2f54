#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s40_2gpu.json 2> gpurun_out/s40_2gpu.err; echo rc2=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/s40_2gpu_ref.json 2>> gpurun_out/s40_2gpu.err; echo rcref=$?
python bench.py > gpurun_out/s40_bench.json 2> gpurun_out/s40_bench.err; echo rc1=$?
timeout 300 python -m pytest tests -x -q -m gpu -k "two_ranks or allreduce" 2>&1 | tail -n 2
