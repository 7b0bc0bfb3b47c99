#!/bin/bash
set -x
mkdir -p gpurun_out
ABF="-saddle_ksp_type fgmres -fs -saddle_fieldsplit_u_pc_type mg -saddle_fieldsplit_u_ksp_type gcr -saddle_fieldsplit_u_ksp_rtol 1e-2 -saddle_fieldsplit_u_mg_levels_pc_type jacobi -saddle_fieldsplit_u_mg_levels_ksp_type chebyshev -saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig 0,0.2,0,1.1 -saddle_fieldsplit_u_mg_levels_ksp_max_it 8 -saddle_fieldsplit_u_mg_levels_ksp_norm_type none -saddle_fieldsplit_u_pc_mg_galerkin -saddle_fieldsplit_p_ksp_type preonly -saddle_fieldsplit_p_pc_type bjacobi"
W64="$ABF -saddle_fieldsplit_u_pc_mg_levels 6 -mx 64 -model 6 -eta0 1 -eta1 1e6 -saddle_ksp_rtol 1e-8"
for t in 0 1 2 3 4; do
  timeout 120 python scripts/mf_one.py 64 4 20 $t 2>&1 | tail -1 | tee -a gpurun_out/r02_c4_tiles.json
  timeout 200 python scripts/run_case.py --solves 3 -- $W64 -xsb_matrix_free full -xsb_mf_tile $t 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('its','solve_s','true_rel_res')}))" | tee -a gpurun_out/r02_c4_tiles.json
done
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "matrix_free or operator_free or lame_64" > gpurun_out/r02_c4_pytest.log 2>&1; tail -3 gpurun_out/r02_c4_pytest.log
