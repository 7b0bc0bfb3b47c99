timeout 600 python scripts/fixture_probe.py > gpurun_out/c12_fixture_probe.json 2> gpurun_out/c12_fixture_probe.err; cat gpurun_out/c12_fixture_probe.json; tail -3 gpurun_out/c12_fixture_probe.err
