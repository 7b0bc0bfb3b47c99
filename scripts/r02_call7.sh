#!/bin/bash
# Guarded GPU call for the two-barrier staged kernel (version C): parity first, then tile timings and solves.
set -x
mkdir -p gpurun_out
ABF="-saddle_ksp_type fgmres -fs -saddle_fieldsplit_u_pc_type mg -saddle_fieldsplit_u_ksp_type gcr -saddle_fieldsplit_u_ksp_rtol 1e-2 -saddle_fieldsplit_u_mg_levels_pc_type jacobi -saddle_fieldsplit_u_mg_levels_ksp_type chebyshev -saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig 0,0.2,0,1.1 -saddle_fieldsplit_u_mg_levels_ksp_max_it 8 -saddle_fieldsplit_u_mg_levels_ksp_norm_type none -saddle_fieldsplit_u_pc_mg_galerkin -saddle_fieldsplit_p_ksp_type preonly -saddle_fieldsplit_p_pc_type bjacobi"
W64="$ABF -saddle_fieldsplit_u_pc_mg_levels 6 -mx 64 -model 6 -eta0 1 -eta1 1e6 -saddle_ksp_rtol 1e-8"
rm -f gpurun_out/r02_c8_*.json
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x --timeout 60 -k "matrix_free_apply" > gpurun_out/r02_c8_pytest_mf.log 2>&1
rc=$?; tail -5 gpurun_out/r02_c8_pytest_mf.log
if [ $rc -ne 0 ]; then echo "ELEMENT KERNEL PARITY FAILED: stopping"; exit 1; fi
for t in 0 1 2 4; do timeout 100 python scripts/mf_one.py 64 4 20 $t 2>&1 | tail -1 | tee -a gpurun_out/r02_c8_tiles.json; done
for t in 0 1 2; do
  timeout 120 python scripts/run_case.py --solves 3 -- $W64 -xsb_matrix_free full -xsb_mf_tile $t 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'tile': $t, **{k:d[k] for k in ('its','solve_s','true_rel_res')}}))" | tee -a gpurun_out/r02_c8_tiles.json
done
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -k "matrix_free or operator_free or lame_64 or abf_opts or baseline_size" > gpurun_out/r02_c8_pytest.log 2>&1; tail -5 gpurun_out/r02_c8_pytest.log
timeout 120 python scripts/run_case.py --solves 1 -- $W64 -xsb_matrix_free full -xsb_graph 0 > gpurun_out/r02_c8_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mf_onepass -s 14 -c 1 -o gpurun_out/r02_onepass_cheb_b python scripts/run_case.py --solves 1 -- $W64 -xsb_matrix_free full -xsb_graph 0 > gpurun_out/r02_c8_ncu.log 2>&1
tail -2 gpurun_out/r02_c8_ncu.log
