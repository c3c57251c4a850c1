"""CPU oracle for the batched LinMPC / linear-MHE hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import, link or execute it, and there only as the checker
or as the timed CPU baseline -- never as the thing shipped.  The product path
(``modelpredictivecontrol.jl_b200`` -> ``libbmpc.so``) fails loudly when its CUDA
library is missing and never routes through this package.

What it is: a numpy fp64 restatement of the reference's algorithm for the path named by
``BASELINE.json.north_star`` -- ``moveinput!`` of ``LinMPC`` (reference
``src/controller/execute.jl:59-80``) and the linear ``MovingHorizonEstimator`` step
(``src/estimator/mhe/execute.jl:44-84``) -- each function citing the reference
file:line it follows.  The reference is 100 % Julia and neither Julia nor its solver
dependencies (OSQP.jl compat "0.8", JuMP "1.21", DAQP; no Manifest.toml, versions
unpinned, sources not vendored under /root/reference) exist in this image, so the
reference itself cannot be run here.

Pinning: the restatement is checked (``tests/test_oracle_*.py``) against every inline
known answer the reference's own tests/doctests hold for this path (SURVEY.md section 8c /
Appendix D): ``u ~ 1`` / ``dU ~ 2`` (test/3_test_predictive_control.jl:95-106), LQR
equivalence 1e-5 (:498-527), constraint activation (:391-464), move-blocking zeros
(:135-140), doctest ``u = 17.577311`` (ext/LinearMPCext.jl:252-261), MHE == KalmanFilter
1e-6 (test/2_test_state_estim.jl:1750-1784), MHE doctest 0.5
(src/estimator/mhe/execute.jl:134-144).  For iterate-level agreement with the OSQP
binary itself parity is UNPINNED beyond the 1e-3..1e-2 the reference's own assertions
tolerate: the QP is solved here to its exact optimum (KKT residual <= 1e-9).
"""
