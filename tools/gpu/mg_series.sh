#!/bin/bash
# multi-GPU series on an N-GPU box: MG parity tests for this world size, then cfg3 / cfg4 (and cfg5 on 8 GPUs) benches
N=${1:-2}; tag=${2:-v1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
show() { python - "$1" <<P
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(l["metric"], "n_gpus", l["n_gpus"], "ms", round(l["ms_per_step"],1), "e2e", l.get("e2e",{}).get("ms_per_step"), l["stages_ms"], l.get("parity",{}).get("family_C_same_order"))
except Exception as e:
    print("no line:", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
P
}
(timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -q -x -k "${TESTK:-$N]}" 2>&1 | tail -60) | tee gpurun_out/r02_mgtests_g${N}_$tag.log
(timeout 400 $TR --steps 3 --warmup 2 > gpurun_out/r02_scale_z8192_g${N}_$tag.json 2> gpurun_out/r02_scale_z8192_g${N}_$tag.err); show gpurun_out/r02_scale_z8192_g${N}_$tag.json
(timeout 400 $TR --dtype d --order 16384 --wanted 2048 --steps 2 --warmup 1 > gpurun_out/r02_scale_d16384_g${N}_$tag.json 2> gpurun_out/r02_scale_d16384_g${N}_$tag.err); show gpurun_out/r02_scale_d16384_g${N}_$tag.json
if [ "$N" = "8" ]; then
  (timeout 900 $TR --dtype z --order 32768 --wanted 4096 --steps 1 --warmup 1 --no-e2e --no-1gpu-compare > gpurun_out/r02_scale_z32768_g${N}_$tag.json 2> gpurun_out/r02_scale_z32768_g${N}_$tag.err); show gpurun_out/r02_scale_z32768_g${N}_$tag.json
fi
