#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "ilu or vcycle or abf_solve or baseline_size" > gpurun_out/r02_ilu_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_ilu_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ilu_bench.json 2> gpurun_out/r02_ilu_bench.err; echo "bench rc=$?"
tail -c 500 gpurun_out/r02_ilu_bench.err
python scripts/bench_digest.py gpurun_out/r02_ilu_bench.json
