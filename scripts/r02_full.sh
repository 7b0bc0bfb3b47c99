#!/bin/bash
# whole GPU suite + default bench on one GPU (what the driver runs at round end)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -x -m gpu > gpurun_out/r02_full_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02_full_pytest.log
timeout 900 python bench.py > gpurun_out/r02_full_bench.json 2> gpurun_out/r02_full_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_full_bench.err
python scripts/bench_digest.py gpurun_out/r02_full_bench.json
