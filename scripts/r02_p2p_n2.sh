#!/bin/bash
# first GPU run of the peer-memory halo exchange + plane-distributed coarse levels: 2 GPUs, guarded by timeouts
set -x
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29531 scripts/slab_check.py 8 > gpurun_out/r02_p2p_slab_check_n$N.log 2>&1; echo "slab_check rc=$?"
tail -15 gpurun_out/r02_p2p_slab_check_n$N.log
cp gpurun_out/slab_check_$N.json gpurun_out/r02_p2p_slab_check_n$N.json
timeout 600 $TR --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_p2p_bench_n$N.json 2> gpurun_out/r02_p2p_bench_n$N.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_p2p_bench_n$N.err
python scripts/bench_digest.py gpurun_out/r02_p2p_bench_n$N.json
