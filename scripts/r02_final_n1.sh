#!/bin/bash
# end-of-round record on one GPU: whole GPU suite, smoke(), default bench, then (only after those exited 0) the ncu launch list
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -x -m gpu > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; rc=$?; echo "bench rc=$rc"
tail -c 300 gpurun_out/r02_final_bench_n1.err
python scripts/bench_digest.py gpurun_out/r02_final_bench_n1.json
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r02_launches_operator_free_64cubed.csv python bench.py --steps 1 --warmup 3 --no-assembled --no-strong128 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu rc=$?"
  wc -l gpurun_out/r02_launches_operator_free_64cubed.csv
fi
