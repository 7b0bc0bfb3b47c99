timeout 300 python -m pytest tests -m gpu -q > gpurun_out/c14_pytest.log 2>&1; tail -3 gpurun_out/c14_pytest.log | cut -c 1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 23000 -c 2500 --csv --log-file gpurun_out/c14_launches_assembled.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-matrix-free > gpurun_out/c14_ncu_bench.log 2>&1; wc -l gpurun_out/c14_launches_assembled.csv
