#!/bin/bash
# option sweep of the distributed tridiagonalization: tools/gpu/mg_opts.sh P N z|d "opts1" "opts2" ...
P=$1; N=$2; T=$3; shift 3
for o in "$@"; do
  echo "== $o"
  EIGB_OPTS="$o" python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29533 tools/test_mg_hetrd.py $N $T 2>&1 | grep "cols\|rank 0: n=" | cut -c1-260
done
