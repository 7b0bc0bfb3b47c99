#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/r02_c3_pytest.log 2>&1; tail -5 gpurun_out/r02_c3_pytest.log
python scripts/mf_one.py 64 4 5 > gpurun_out/r02_c3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mf_onepass -s 3 -c 1 -o gpurun_out/r02_onepass_v0 python scripts/mf_one.py 64 4 5 > gpurun_out/r02_c3_ncu.log 2>&1
tail -3 gpurun_out/r02_c3_plain.log gpurun_out/r02_c3_ncu.log
