"""Compare GPU residual histories at the BASELINE sizes against the committed oracle fixtures (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import exsaddle_b200 as X
out = {}
for mx in (32, 64):
    fx = json.load(open("tests/golden/oracle_%dcubed_history.json" % mx))
    ho = np.array(fx["hist"])
    for tag, extra in (("assembled", ""), ("operator_free", " -xsb_matrix_free full")):
        g = X.ExSaddle(fx["options"] + extra, nsd=3).assemble().ksp_setup()
        x = g.solve(); h = g.history(); n = min(len(h), len(ho))
        d = np.abs(h[:n] - ho[:n])
        bands = {}
        for lo in (1e-2, 1e-4, 1e-6, 0.0):
            m = ho[:n] >= lo * ho[0]
            bands["rel>=%g" % lo] = float(np.max(d[m] / ho[:n][m]))
        inner = g.inner_iterations()
        out["%d_%s" % (mx, tag)] = {"its": g.iterations(), "its_oracle": fx["its"], "ksp_rel": float(np.max(d) / ho[0]), "bands": bands,
                                    "inner_diff": [(i, a, b) for i, (a, b) in enumerate(zip(inner, fx["inner_its"])) if a != b],
                                    "tail_gpu": [float(v) for v in h[-3:]], "tail_oracle": [float(v) for v in ho[-3:]],
                                    "xnorm_rel": float(abs(np.linalg.norm(x) - fx["x_norm2"]) / fx["x_norm2"]),
                                    "cheb_rel": [float(abs(g.chebyshev(l)[1] - fx["cheb_emax_est"][l]) / fx["cheb_emax_est"][l]) for l in range(1, fx["levels"])]}
        g.close()
print(json.dumps(out, indent=1))
