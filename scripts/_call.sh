timeout 300 python -m pytest tests -m gpu -x -q -k "matrix_free" > gpurun_out/c3_pytest.log 2>&1; tail -1 gpurun_out/c3_pytest.log
for rv in 1 0; do echo "reverse $rv"; timeout 300 python scripts/mf_bench.py 64 20 "-xsb_mf_reverse $rv" 2>/dev/null; done
echo "32^3"; timeout 300 python scripts/mf_bench.py 32 20 2>/dev/null
