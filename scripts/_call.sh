timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c11_pytest.log 2>&1; tail -4 gpurun_out/c11_pytest.log | cut -c 1-300
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; tail -c 1200 gpurun_out/c11_bench.json; tail -3 gpurun_out/c11_bench.err
