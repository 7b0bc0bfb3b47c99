#!/bin/bash
# Comprehensive GPU call: all GPU tests, one-pass kernel tile variants, Chebyshev-mode profile, bench line.
set -x
mkdir -p gpurun_out
ABF="-saddle_ksp_type fgmres -fs -saddle_fieldsplit_u_pc_type mg -saddle_fieldsplit_u_ksp_type gcr -saddle_fieldsplit_u_ksp_rtol 1e-2 -saddle_fieldsplit_u_mg_levels_pc_type jacobi -saddle_fieldsplit_u_mg_levels_ksp_type chebyshev -saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig 0,0.2,0,1.1 -saddle_fieldsplit_u_mg_levels_ksp_max_it 8 -saddle_fieldsplit_u_mg_levels_ksp_norm_type none -saddle_fieldsplit_u_pc_mg_galerkin -saddle_fieldsplit_p_ksp_type preonly -saddle_fieldsplit_p_pc_type bjacobi"
W64="$ABF -saddle_fieldsplit_u_pc_mg_levels 6 -mx 64 -model 6 -eta0 1 -eta1 1e6 -saddle_ksp_rtol 1e-8"
rm -f gpurun_out/r02_c5_tiles.json
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02_c5_pytest.log 2>&1; tail -15 gpurun_out/r02_c5_pytest.log
for t in 0 1 2 3 4; do
  timeout 120 python scripts/mf_one.py 64 4 20 $t 2>&1 | tail -1 | tee -a gpurun_out/r02_c5_tiles.json
  for gr in 1 0; do
  timeout 200 python scripts/run_case.py --solves 3 -- $W64 -xsb_matrix_free full -xsb_mf_tile $t -xsb_graph $gr 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'tile': $t, 'graph': $gr, **{k:d[k] for k in ('its','solve_s','true_rel_res')}}))" | tee -a gpurun_out/r02_c5_tiles.json
  done
done
for gr in 1 0; do timeout 200 python scripts/run_case.py --solves 3 -- $W64 -xsb_graph $gr 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'path': 'assembled', 'graph': $gr, **{k:d[k] for k in ('its','solve_s','true_rel_res')}}))" | tee -a gpurun_out/r02_c5_tiles.json; done
python scripts/run_case.py --solves 1 -- $W64 -xsb_matrix_free full -xsb_mf_tile 1 -xsb_graph 0 > gpurun_out/r02_c5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mf_onepass -s 14 -c 1 -o gpurun_out/r02_onepass_cheb_t1 python scripts/run_case.py --solves 1 -- $W64 -xsb_matrix_free full -xsb_mf_tile 1 -xsb_graph 0 > gpurun_out/r02_c5_ncu.log 2>&1
tail -2 gpurun_out/r02_c5_ncu.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_c5_bench.json 2> gpurun_out/r02_c5_bench.err; tail -c 600 gpurun_out/r02_c5_bench.err; head -c 1500 gpurun_out/r02_c5_bench.json
