"""Outer optimisation loop + termination order of STEAM's Solver, as used by
reference pysteam_augmented/lev_marq_gauss_newton_custom_solver.py through GaussNewtonSolver."""
import abc
import os
from math import fabs

# test hook: record (iterations, final cost, termination) of every optimize() call
TRACE = [] if os.environ.get("DRE_ORACLE_TRACE") else None


class Solver(abc.ABC):
    def __init__(self, problem, **parameters):
        self._parameters = {
            "verbose": False,
            "max_iterations": 100,
            "absolute_cost_threshold": 0.0,
            "absolute_cost_change_threshold": 1e-4,
            "relative_cost_change_threshold": 1e-4,
        }
        self._parameters.update(parameters)
        self._problem = problem
        self._state_vector = problem.get_state_vector()
        self._state_vector_backup = self._state_vector.clone()
        self._pending_proposed_state = False
        self._curr_iteration = 0
        self._solver_converged = False
        self._term = None
        self._curr_cost = self._prev_cost = self._problem.cost()

    @property
    def curr_iteration(self):
        return self._curr_iteration

    def optimize(self):
        try:
            while not self._solver_converged:
                self.iterate()
        finally:
            if TRACE is not None:
                TRACE.append((self._curr_iteration, self._curr_cost, self._term))

    def iterate(self):
        if self._solver_converged:
            return
        self._curr_iteration += 1
        self._prev_cost = self._curr_cost
        step_success, self._curr_cost, grad_norm = self.linearize_solve_and_update()
        if (not step_success) and fabs(grad_norm) < 1e-6:
            self._term, self._solver_converged = "CONVERGED_ZERO_GRADIENT", True
        elif not step_success:
            self._term, self._solver_converged = "STEP_UNSUCCESSFUL", True
            raise RuntimeError("The solver terminated due to an unsuccessful step.")
        elif self._curr_iteration >= self._parameters["max_iterations"]:
            self._term, self._solver_converged = "MAX_ITERATIONS", True
        elif self._curr_cost <= self._parameters["absolute_cost_threshold"]:
            self._term, self._solver_converged = "CONVERGED_ABSOLUTE_ERROR", True
        elif fabs(self._prev_cost - self._curr_cost) <= self._parameters["absolute_cost_change_threshold"]:
            self._term, self._solver_converged = "CONVERGED_ABSOLUTE_CHANGE", True
        elif fabs(self._prev_cost - self._curr_cost) / self._prev_cost <= self._parameters["relative_cost_change_threshold"]:
            self._term, self._solver_converged = "CONVERGED_RELATIVE_CHANGE", True

    def propose_update(self, perturbation):
        if self._pending_proposed_state:
            raise RuntimeError("There is already a pending update, accept or reject.")
        self._state_vector_backup.copy_values(self._state_vector)
        self._state_vector.update(perturbation)
        self._pending_proposed_state = True
        return self._problem.cost()

    def accept_proposed_state(self):
        if not self._pending_proposed_state:
            raise RuntimeError("You must call propose_update before accept.")
        self._pending_proposed_state = False

    def reject_proposed_state(self):
        if not self._pending_proposed_state:
            raise RuntimeError("You must call propose_update before rejecting.")
        self._state_vector.copy_values(self._state_vector_backup)
        self._pending_proposed_state = False

    @abc.abstractmethod
    def linearize_solve_and_update(self):
        ...
