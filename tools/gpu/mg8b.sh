#!/bin/bash
# cfg5 only: ZHEGVDX N=32768, il=1..4096 on N GPUs, one timed step, device-resident
N=${1:-8}; tag=${2:-v1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
(timeout 260 $TR --dtype z --order 32768 --wanted 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --no-1gpu-compare > gpurun_out/r02_s8_z32768_g${N}_$tag.json 2> gpurun_out/r02_s8_z32768_g${N}_$tag.err)
python - gpurun_out/r02_s8_z32768_g${N}_$tag.json <<P
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(l["metric"], "n_gpus", l["n_gpus"], "ms", round(l["ms_per_step"],1), l["stages_ms"], l.get("parity"))
except Exception as e:
    print("no line:", e); print("\n".join([x for x in open(sys.argv[1].replace(".json",".err")).read().splitlines() if "rror" in x or "failed" in x][:12]))
P
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
