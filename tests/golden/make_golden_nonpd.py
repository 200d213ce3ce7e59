"""Golden traces of FAILED factorisations, from the reference itself (run in the build container):

    make -C oracle ref && python tests/golden/make_golden_nonpd.py

LinearSolverCSparse::solve returns false when the up-looking Cholesky meets a pivot <= 0
(thirdparty/g2o/g2o/solvers/csparse/csparse_extension.cpp:115); OptimizationAlgorithmLevenberg then sets
tempChi = DBL_MAX, rejects the trial and inflates lambda (optimization_algorithm_levenberg.cpp:119-143); ten
rejected trials end the optimisation with Terminate (:147).  Two ways to get there, both stored:

  * `lambda30`  a gauge-free graph (no pose fixed, as backend.cpp:93-103) with _userLambdaInit = 1e-30: S + lambda I
                is numerically singular, every trial fails, optimize() stops after one iteration.
  * `indef_*`   a fraction of the edges carries a NEGATIVE-definite information matrix (the g2o API allows any
                symmetric matrix): the reduced system is genuinely indefinite until lambda has grown past its
                most negative eigenvalue, so the failures do not depend on the elimination order.

Inputs are regenerated from the recipe (synth config + numpy PCG64 seed); nonpd.json holds what the reference did:
per iteration (robust chi2, lambda, trials), the number of failed factorisations, the final chi2.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.bindings import RefOracle  # noqa: E402
from ssvio_b200 import synth  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import NONPD_CASES as CASES, nonpd_case_inputs as case_inputs  # noqa: E402


def main():
    ref = RefOracle()
    ref.lib.ssba_ref_set_extras.argtypes = [C.c_double, C.c_int32, C.POINTER(C.c_double)]
    ref.lib.ssba_ref_set_extras.restype = None
    os.chdir(tempfile.mkdtemp())
    out = {}
    for name in CASES:
        g, info, ul, iters = case_inputs(name)
        ref.lib.ssba_ref_set_extras(ul, 0, info.ctypes.data_as(C.POINTER(C.c_double)) if info is not None else None)
        r = ref.optimize(g, iters=iters, jacobian="analytic", trace=True)
        rep = r["report"]
        out[name] = dict(iterations=rep.iterations, cholesky_failures=rep.cholesky_failures, chi2_initial=rep.chi2_initial,
                         chi2_robust=rep.chi2_robust, chi2_plain=rep.chi2_plain, lambda_final=rep.lambda_,
                         trace=[list(t) for t in rep.trace()],
                         n_negative_edges=int((info[:, 0] < 0).sum()) if info is not None else 0)
        print(name, json.dumps(out[name]))
    json.dump(out, open(os.path.join(HERE, "nonpd.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
