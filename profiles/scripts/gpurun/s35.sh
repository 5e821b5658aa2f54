#!/bin/bash
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/s35_bench.json 2> gpurun_out/s35_bench.err; echo bench rc=$?
python - <<'P'
import json
d=json.load(open('gpurun_out/s35_bench.json'))
print(d['value'], d['e2e']['value'])
print(json.dumps(d['workloads']['reference_batch_sizes'], indent=1))
P
tail -n 5 gpurun_out/s35_bench.err
