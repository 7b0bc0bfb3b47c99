"""Drop-in front end: reproduces the stdout of ./exSaddle{2d,3d}{,_lame} <options> (exSaddle.c:105-566)
for the solver trees this library covers, so the reference's testref/*.ref files can be diffed directly."""
from .api import ExSaddle

_REASON = {2: "CONVERGED_RTOL", 3: "CONVERGED_ATOL", -3: "DIVERGED_ITS", -4: "DIVERGED_DTOL", -5: "DIVERGED_BREAKDOWN"}


def monitor_short(v):
    """KSPMonitorDefaultShort number format."""
    if v > 1.e-9:
        return "%g" % v
    if v > 1.e-11:
        return "%5.3e" % v
    return "< 1.e-11"


def diagnostics_text(nsd, d):
    """SaddleReportSolutionDiagnostics (exSaddle_io.c:17-56)."""
    tag = "|u,v|" if nsd == 2 else "|u,v,w|"
    names = ("_1  ", "_2  ", "_inf", "_min", "_max")
    lines = []
    for k, nm in enumerate(names):
        lines.append("%s%s %s%s" % (tag, nm, " , ".join("%+1.6e" % d[k * nsd + c] for c in range(nsd)), " " if nsd == 2 else ""))
    for k, nm in enumerate(names):
        lines.append("|p|%s        %+1.6e" % (nm, d[5 * nsd + k]))
    return lines


def run_exsaddle(exe, options, options_file_dir=None, outdir=".", nranks=1):
    """exe in {exSaddle2d, exSaddle3d, exSaddle2d_lame, exSaddle3d_lame}; returns (stdout_text, ExSaddle, x).
    nranks: the `mpiexec -n` of the reference's command line; it only matters where the reference's algorithm depends on the rank
    count (one ASM element patch per rank) and is passed to the library as -xsb_ranks.
    -dump_solution / -dump_operator / -dump_scaled_mass_matrix write PETSc binary files into `outdir` under the
    reference's file names (exSaddle.c:488-501, 535-537)."""
    import os
    nsd = 2 if "2d" in exe else 3
    lame = "lame" in exe
    s = ExSaddle(nsd=nsd, lame=lame)
    toks = options.split()
    if "-options_file" in toks:
        i = toks.index("-options_file"); path = toks[i + 1]
        if not os.path.isabs(path) and options_file_dir:
            path = os.path.join(options_file_dir, path)
        s.set_options(" ".join(toks[:i] + toks[i + 2:]))
        s.set_options_file(path)
    else:
        s.set_options(options)
    if nranks > 1:
        s.set_options("-xsb_ranks %d" % nranks)
    out = []
    out.append(s.banner().rstrip("\n"))
    s.assemble()
    s.ksp_setup()
    x = s.solve()
    its, reason = s.iterations()
    if "-saddle_ksp_monitor_short" in toks:
        out.append("  Residual norms for saddle_ solve.")
        inner, why = s.inner_iterations(), s.inner_reasons()
        show_inner = "-saddle_fieldsplit_u_ksp_converged_reason" in toks
        for i, r in enumerate(s.history()):
            if show_inner and i > 0 and i - 1 < len(inner):
                w = why[i - 1] if i - 1 < len(why) else 2
                out.append("  Linear saddle_fieldsplit_u_ solve %s due to %s iterations %d" % ("converged" if w > 0 else "did not converge", _REASON.get(w, str(w)), inner[i - 1]))
            out.append("%3d KSP Residual norm %s " % (i, monitor_short(r)))
    if "-saddle_ksp_converged_reason" in toks:
        if reason > 0:
            out.append("Linear saddle_ solve converged due to %s iterations %d" % (_REASON.get(reason, str(reason)), its))
        else:
            out.append("Linear saddle_ solve did not converge due to %s iterations %d" % (_REASON.get(reason, str(reason)), its))
    if "-diagnostics" in toks:
        out.extend(diagnostics_text(nsd, s.diagnostics(x)))
    from .api import MAT_A, MAT_MP
    if "-view_fields" in toks:         # ViewFields(dm_saddle, X, ""), exSaddle.c:481-483
        s.view_fields(x, outdir, "")
    if "-dump_solution" in toks:       # DumpSolution, exSaddle_io.c:76-88
        out.append("Dumping solution vector to solution.petscbin.")
        s.dump_vector(x, os.path.join(outdir, "solution.petscbin"))
        out.append("Finished dumping vector to solution.petscbin.")
    if "-dump_operator" in toks:       # DumpOperator, exSaddle_io.c:61-73 (finest level: operator_<nlevels-1>)
        k = 0
        if "-nlevels" in toks:
            k = int(toks[toks.index("-nlevels") + 1]) - 1
        name = "operator_%d.petscbin" % k
        out.append("Dumping operator to %s. This could be very slow!" % name)
        s.dump_operator(MAT_A, os.path.join(outdir, name))
        out.append("Finished dumping operator to %s." % name)
    if "-dump_scaled_mass_matrix" in toks:
        if "-fs" not in toks:
            raise ValueError("-dump_scaled_mass_matrix without -fs")   # exSaddle.c:213
        out.append("Dumping operator to mpscaled.petscbin. This could be very slow!")
        s.dump_operator(MAT_MP, os.path.join(outdir, "mpscaled.petscbin"))
        out.append("Finished dumping operator to mpscaled.petscbin.")
    return "\n".join(out) + "\n", s, x
