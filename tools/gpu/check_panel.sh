#!/bin/bash
# 1-GPU check after a panel-kernel change: parity suites that go through hetrd, then the headline bench and cfg4.
tag=${1:-x}
(timeout 700 python -m pytest tests/test_sytrd_gpu.py tests/test_driver_gpu.py tests/test_large_gpu.py tests/test_golden.py tests/test_stages_gpu.py -q -m gpu 2>&1 | tail -12) | tee gpurun_out/r02_t_$tag.log
(timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err)
python - <<P
import json
l=json.loads(open("gpurun_out/r02_bench_$tag.json").read().strip().splitlines()[-1])
print(l["ms_per_step"], l["e2e"]["ms_per_step"], l["stages_ms"], l["roofline"]["frac"], l["parity"]["family_C_same_order"])
P
tail -3 gpurun_out/r02_bench_$tag.err
(timeout 200 python bench.py --dtype d --n 16384 --m 2048 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02_bench_d16384_$tag.json 2>> gpurun_out/r02_bench_$tag.err)
python - <<P
import json
l=json.loads(open("gpurun_out/r02_bench_d16384_$tag.json").read().strip().splitlines()[-1])
print(l["ms_per_step"], l["stages_ms"], l["roofline"]["frac"], l["parity"]["family_C_same_order"])
P
