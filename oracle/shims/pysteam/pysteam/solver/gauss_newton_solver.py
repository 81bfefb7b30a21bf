"""GaussNewtonSolver base of the reference's LM solver (lev_marq...py:6,10,155).
[3P-UNVERIFIED] switch: dense solve by Cholesky (raises LinAlgError when the damped Hessian is
not positive definite -> counted as a failed decomposition at lev_marq...py:47-48) or by LU."""
import os
import numpy as np
import numpy.linalg as npla
import scipy.linalg as spla
from .solver import Solver

DENSE_SOLVE = os.environ.get("DRE_ORACLE_DENSE_SOLVE", "cholesky")


class GaussNewtonSolver(Solver):
    def __init__(self, problem, **parameters):
        super().__init__(problem, **parameters)

    def linearize_solve_and_update(self):
        A, b = self._problem.build_gauss_newton_terms()
        grad_norm = npla.norm(b)
        perturbation = self.solve_gauss_newton(A, b)
        new_cost = self.propose_update(perturbation)
        self.accept_proposed_state()
        return True, new_cost, grad_norm

    def solve_gauss_newton(self, A, b):
        if DENSE_SOLVE == "cholesky":
            try:
                c = spla.cho_factor(A, lower=True, check_finite=False)
            except spla.LinAlgError as e:
                raise npla.LinAlgError(str(e))
            return spla.cho_solve(c, b, check_finite=False)
        return npla.solve(A, b)
