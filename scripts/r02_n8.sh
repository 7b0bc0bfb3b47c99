#!/bin/bash
set -x
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_final_bench_n$N.json 2> gpurun_out/r02_final_bench_n$N.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_final_bench_n$N.err
python scripts/bench_digest.py gpurun_out/r02_final_bench_n$N.json
XSB_BENCH_EXTRA_OPTS="-xsb_pdist_min_nodes 30000" timeout 400 $TR --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --no-assembled > gpurun_out/r02_pd30k_bench_n$N.json 2> gpurun_out/r02_pd30k_bench_n$N.err; echo "bench rc=$?"
python scripts/bench_digest.py gpurun_out/r02_pd30k_bench_n$N.json
