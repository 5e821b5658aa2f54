#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/s49_4gpu.json 2> gpurun_out/s49_4gpu.err; echo rc4=$?
