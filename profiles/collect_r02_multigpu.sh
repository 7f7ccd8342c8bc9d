#!/bin/bash
# multi-GPU legs of the round-2 collection: bash profiles/collect_r02_multigpu.sh N   (under gpurun --gpus N)
N=$1
O=gpurun_out
mkdir -p $O
for w in c3 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload $w > $O/r02_bench_${w}_${N}gpu.json 2> $O/r02_bench_${w}_${N}gpu.err || echo "FAIL $w $N"
  tail -c 600 $O/r02_bench_${w}_${N}gpu.json
done
