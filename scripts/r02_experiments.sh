#!/bin/bash
# First GPU call of round 2: measure the opt-in variants written (unmeasured) at the end of round 1.
#   gpurun --timeout 900 -- 'bash scripts/r02_experiments.sh'
# Everything lands in gpurun_out/r02_*.  One GPU, ~4 minutes; no ncu here (bounded captures come after, with -c).
set -x
mkdir -p gpurun_out
ABF="-saddle_ksp_type fgmres -fs -saddle_fieldsplit_u_pc_type mg -saddle_fieldsplit_u_ksp_type gcr -saddle_fieldsplit_u_ksp_rtol 1e-2 -saddle_fieldsplit_u_mg_levels_pc_type jacobi -saddle_fieldsplit_u_mg_levels_ksp_type chebyshev -saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig 0,0.2,0,1.1 -saddle_fieldsplit_u_mg_levels_ksp_max_it 8 -saddle_fieldsplit_u_mg_levels_ksp_norm_type none -saddle_fieldsplit_u_pc_mg_galerkin -saddle_fieldsplit_p_ksp_type preonly -saddle_fieldsplit_p_pc_type bjacobi"
W64="$ABF -saddle_fieldsplit_u_pc_mg_levels 6 -mx 64 -model 6 -eta0 1 -eta1 1e6 -saddle_ksp_rtol 1e-8"
# 1. correctness of the experimental kernels (bitwise / 1e-12 against the default kernels and the oracle)
XSB_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k experimental > gpurun_out/r02_experimental_pytest.log 2>&1; tail -3 gpurun_out/r02_experimental_pytest.log
# 2. element kernel: accumulator pinned in L2 or not (stand-alone product times)
for o in "" "-xsb_mf_fused_zero" "-xsb_mf_l2_persist" "-xsb_mf_fused_zero -xsb_mf_l2_persist"; do timeout 200 python scripts/mf_bench.py 64 20 "$o" 2>/dev/null | tee -a gpurun_out/r02_mf_l2.json; done
# 3. ILU(0): cluster kernel vs single-CTA ring kernel, operator-free 64^3 solve (17 % of it was ILU)
for k in 1 2; do timeout 200 python scripts/run_case.py --solves 3 -- $W64 -xsb_matrix_free full -xsb_ilu_kernel $k 2>/dev/null | tee -a gpurun_out/r02_ilu.json; done
# 4. fine-level BAIJ product with closed-form columns, assembled 64^3 solve
for o in "" "-xsb_baij_closed_form"; do timeout 200 python scripts/run_case.py --solves 3 -- $W64 $o 2>/dev/null | tee -a gpurun_out/r02_baij_cf.json; done
for o in "" "-xsb_baij_closed_form"; do timeout 200 python scripts/spmv_bench.py 64 30 "$o" 2>/dev/null | tee -a gpurun_out/r02_spmv_bench.json; done
