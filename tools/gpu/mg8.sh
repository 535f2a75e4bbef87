#!/bin/bash
# one short 8-GPU pass: cfg3 (parity check of the 8-rank path), cfg5 (N=32768, m=4096), cfg4 -- one step each, device-resident
N=${1:-8}; tag=${2:-v1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
show() { python - "$1" <<P
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(l["metric"], "n_gpus", l["n_gpus"], "ms", round(l["ms_per_step"],1), l["stages_ms"], l.get("parity",{}).get("family_C_same_order"))
except Exception as e:
    print("no line:", e, open(sys.argv[1].replace(".json",".err")).read()[-1500:])
P
}
(timeout 120 $TR --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_s8_z8192_g${N}_$tag.json 2> gpurun_out/r02_s8_z8192_g${N}_$tag.err); show gpurun_out/r02_s8_z8192_g${N}_$tag.json
(timeout 200 $TR --dtype z --order 32768 --wanted 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --no-1gpu-compare > gpurun_out/r02_s8_z32768_g${N}_$tag.json 2> gpurun_out/r02_s8_z32768_g${N}_$tag.err); show gpurun_out/r02_s8_z32768_g${N}_$tag.json
(timeout 120 $TR --dtype d --order 16384 --wanted 2048 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_s8_d16384_g${N}_$tag.json 2> gpurun_out/r02_s8_d16384_g${N}_$tag.err); show gpurun_out/r02_s8_d16384_g${N}_$tag.json
