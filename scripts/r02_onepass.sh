#!/bin/bash
# One-pass element kernel: parity tests, product timing, operator-free solve.  gpurun --timeout 900 -- 'bash scripts/r02_onepass.sh'
set -x
mkdir -p gpurun_out
ABF="-saddle_ksp_type fgmres -fs -saddle_fieldsplit_u_pc_type mg -saddle_fieldsplit_u_ksp_type gcr -saddle_fieldsplit_u_ksp_rtol 1e-2 -saddle_fieldsplit_u_mg_levels_pc_type jacobi -saddle_fieldsplit_u_mg_levels_ksp_type chebyshev -saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig 0,0.2,0,1.1 -saddle_fieldsplit_u_mg_levels_ksp_max_it 8 -saddle_fieldsplit_u_mg_levels_ksp_norm_type none -saddle_fieldsplit_u_pc_mg_galerkin -saddle_fieldsplit_p_ksp_type preonly -saddle_fieldsplit_p_pc_type bjacobi"
W64="$ABF -saddle_fieldsplit_u_pc_mg_levels 6 -mx 64 -model 6 -eta0 1 -eta1 1e6 -saddle_ksp_rtol 1e-8"
timeout 500 python -m pytest tests/test_gpu_parity.py -q -x -k "matrix_free or operator_free" > gpurun_out/r02_onepass_pytest.log 2>&1; tail -15 gpurun_out/r02_onepass_pytest.log
timeout 200 python scripts/mf_bench.py 64 20 2>&1 | tail -2 | tee gpurun_out/r02_onepass_mfbench.json
for k in 3 4; do timeout 200 python scripts/run_case.py --solves 3 -- $W64 -xsb_matrix_free full -xsb_mf_kernel $k 2>&1 | tail -1 | tee -a gpurun_out/r02_onepass_solve.json; done
