timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c9_pytest.log 2>&1; tail -40 gpurun_out/c9_pytest.log | cut -c 1-400
