#!/usr/bin/env python
"""Extract the reference's own known-answer values into tests/golden/testref_kat.json.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Sources: /root/reference/Makefile:254-513 (the option string of every regression test) and
/root/reference/testref/*.ref (the golden stdout of each).  Only numbers and option strings are
extracted (residual histories, converged reasons, diagnostics block, -ksp_view sizes).
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "testref_kat.json")


def makefile_cases():
    cases = {}
    txt = open(os.path.join(REF, "Makefile")).read()
    for m in re.finditer(r"^test_(\w+)\s*:\s*\n\s*-@\$\{MPIEXEC\}(.*?)>\s*\w+\.tmp", txt, re.M | re.S):
        name, cmd = m.group(1), " ".join(m.group(2).split())
        nranks = 1
        mm = re.match(r"-n (\d+) (.*)", cmd)
        if mm:
            nranks, cmd = int(mm.group(1)), mm.group(2)
        exe, opts = cmd.split(" ", 1)
        exe = exe.replace("./", "")
        cases[name] = {"exe": exe, "nsd": 2 if "2d" in exe else 3, "lame": "lame" in exe, "nranks": nranks, "options": opts.strip()}
    return cases


def parse_ref(path):
    out = {"residuals": [], "residuals_text": [], "inner_its": [], "diagnostics": [], "banner": []}
    for line in open(path):
        s = line.rstrip("\n")
        m = re.match(r"\s*(\d+) KSP Residual norm (\S+)", s)
        if m and s.startswith("  ") is not None and "Residual norms" not in s:
            if len(s) - len(s.lstrip()) <= 2:   # top-level monitor only
                out["residuals"].append(float(m.group(2)) if m.group(2) != "<" else 0.0)
                out["residuals_text"].append(m.group(2))
            continue
        m = re.match(r"\s*Linear saddle_fieldsplit_u_ solve converged due to (\w+) iterations (\d+)", s)
        if m:
            out["inner_its"].append(int(m.group(2))); continue
        m = re.match(r"Linear saddle_ solve (converged|did not converge) due to (\w+) iterations (\d+)", s)
        if m:
            out["reason"] = m.group(2); out["iterations"] = int(m.group(3)); continue
        if s.startswith("|u,v") or s.startswith("|p|"):
            out["diagnostics"].append(s); continue
        if s.startswith("Boundary Conditions") or s.startswith("ModelType") or s.startswith("  params:"):
            out["banner"].append(s); continue
        m = re.match(r"\s*eigenvalue estimates used:\s+min = (\S+), max = (\S+)", s)
        if m:
            out.setdefault("cheb_bounds", []).append([float(m.group(1)), float(m.group(2))]); continue
        m = re.match(r"\s*eigenvalues estimate via \w+ min (\S+), max (\S+)", s)
        if m:
            out.setdefault("cheb_ritz", []).append([float(m.group(1)), float(m.group(2))]); continue
        m = re.match(r"\s*rows=(\d+), cols=(\d+)(, bs=(\d+))?", s)
        if m:
            out.setdefault("mat_rows", []).append([int(m.group(1)), int(m.group(2)), int(m.group(4) or 1)]); continue
        m = re.match(r"\s*total: nonzeros=(\d+), allocated nonzeros=(\d+)", s)
        if m:
            out.setdefault("mat_nnz", []).append([int(m.group(1)), int(m.group(2))]); continue
    return out


def main():
    cases = makefile_cases()
    kat = {}
    for name, c in sorted(cases.items()):
        ref = os.path.join(REF, "testref", name + ".ref")
        if not os.path.exists(ref):
            continue
        c.update(parse_ref(ref))
        c["source"] = "testref/%s.ref" % name
        kat[name] = c
    kat["_abf_opts"] = [l.split("#")[0].strip() for l in open(os.path.join(REF, "abf.opts")) if l.split("#")[0].strip()]
    json.dump(kat, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(kat), "cases")


if __name__ == "__main__":
    main()
