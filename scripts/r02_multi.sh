#!/bin/bash
# Multi-GPU call (N = $1, default 2): slab parity against one GPU + block oracle (tests/test_multi_gpu.py), bench line with in-bench parity.
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r02_multi_pytest_n$N.log 2>&1; tail -8 gpurun_out/r02_multi_pytest_n$N.log
cp gpurun_out/slab_check_$N.json gpurun_out/r02_slab_check_n$N.json 2>/dev/null
timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r02_bench_n$N.err; head -c 3000 gpurun_out/r02_bench_n$N.json
