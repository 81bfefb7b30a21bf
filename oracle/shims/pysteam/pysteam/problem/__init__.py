"""Problem side of the pysteam stand-in: names imported at reference mpc.py:8 and
pysteam_augmented/extra_loss_func.py:1, lev_marq_gauss_newton_custom_solver.py:5."""
import numpy as np
import numpy.linalg as npla


class LossFunc:
    def cost(self, whitened_error_norm):
        raise NotImplementedError

    def weight(self, whitened_error_norm):
        raise NotImplementedError


class L2LossFunc(LossFunc):
    def cost(self, whitened_error_norm):
        return 0.5 * whitened_error_norm * whitened_error_norm

    def weight(self, whitened_error_norm):
        return 1.0


class CauchyLossFunc(LossFunc):
    def __init__(self, k=1.0):
        self._k = k

    def cost(self, whitened_error_norm):
        e_div_k = abs(whitened_error_norm) / self._k
        return 0.5 * self._k * self._k * np.log(1.0 + e_div_k * e_div_k)

    def weight(self, whitened_error_norm):
        e_div_k = abs(whitened_error_norm) / self._k
        return 1.0 / (1.0 + e_div_k * e_div_k)


class StaticNoiseModel:
    """Noise model given a covariance: sqrt information = chol(cov^-1)^T (upper)."""

    def __init__(self, matrix, type="covariance"):
        matrix = np.asarray(matrix, dtype=float)
        if type == "covariance":
            info = npla.inv(matrix)
        elif type == "information":
            info = matrix
        else:
            raise ValueError(type)
        self._sqrt_information = npla.cholesky(info).T

    def get_sqrt_information(self):
        return self._sqrt_information

    def get_whitened_error_norm(self, raw_error):
        return npla.norm(self._sqrt_information @ raw_error)

    def whiten_error(self, raw_error):
        return self._sqrt_information @ raw_error


class StateVector:
    class Entry:
        def __init__(self, state_var, indices):
            self.state_var = state_var
            self.indices = indices

    def __init__(self):
        self._state_vars = {}
        self._size = 0

    def add_state_var(self, state_var):
        assert not state_var.locked, "locked variables must not be added (mpc.py:88 note)"
        n = state_var.perturb_dim
        self._state_vars[state_var.key] = StateVector.Entry(state_var, slice(self._size, self._size + n))
        self._size += n

    def get_state_size(self):
        return self._size

    def get_state_indices(self, key):
        return self._state_vars[key].indices

    def clone(self):
        c = StateVector()
        for key, e in self._state_vars.items():
            c._state_vars[key] = StateVector.Entry(e.state_var.clone(), e.indices)
        c._size = self._size
        return c

    def copy_values(self, other):
        for key, e in self._state_vars.items():
            e.state_var.assign(other._state_vars[key].state_var.value)

    def update(self, perturbation):
        for e in self._state_vars.values():
            e.state_var.update(perturbation[e.indices])


class WeightedLeastSquareCostTerm:
    def __init__(self, error_function, noise_model, loss_function):
        self._error_function = error_function
        self._noise_model = noise_model
        self._loss_function = loss_function

    def cost(self):
        error = self._error_function.evaluate()
        return self._loss_function.cost(self._noise_model.get_whitened_error_norm(error))

    @property
    def related_var_keys(self):
        return self._error_function.related_var_keys

    def _evaluate_weighted_and_whitened(self):
        from ..evaluable import Jacobians
        jacs = Jacobians()
        raw_error = self._error_function.evaluate(self._noise_model.get_sqrt_information(), jacs)
        white_error = self._noise_model.whiten_error(raw_error)
        sqrt_w = np.sqrt(self._loss_function.weight(npla.norm(white_error)))
        error = sqrt_w * white_error
        return error, {k: sqrt_w * v for k, v in jacs.get_copy().items()}

    def build_gauss_newton_terms(self, state_vec, A, b):
        error, jacs = self._evaluate_weighted_and_whitened()
        keys = list(jacs.keys())
        for i, k1 in enumerate(keys):
            j1, i1 = jacs[k1], state_vec.get_state_indices(k1)
            b[i1] += -j1.T @ error
            for k2 in keys[i:]:
                j2, i2 = jacs[k2], state_vec.get_state_indices(k2)
                if i1.start <= i2.start:
                    A[i1, i2] += j1.T @ j2
                else:
                    A[i2, i1] += j2.T @ j1


class Problem:
    pass


class OptimizationProblem(Problem):
    def __init__(self, num_threads=1):
        self._cost_terms = []
        self._state_vars = []
        self._state_vector = StateVector()

    def add_state_var(self, *state_vars):
        for v in state_vars:
            self._state_vars.append(v)

    def add_cost_term(self, *cost_terms):
        self._cost_terms.extend(cost_terms)

    def cost(self):
        return sum(float(c.cost()) for c in self._cost_terms)

    def get_state_vector(self):
        self._state_vector = StateVector()
        for v in self._state_vars:
            if not v.locked:
                self._state_vector.add_state_var(v)
        return self._state_vector

    def get_num_of_cost_terms(self):
        return len(self._cost_terms)

    def build_gauss_newton_terms(self):
        n = self._state_vector.get_state_size()
        A = np.zeros((n, n))
        b = np.zeros((n, 1))
        for c in self._cost_terms:
            c.build_gauss_newton_terms(self._state_vector, A, b)
        A = np.triu(A) + np.triu(A, 1).T
        return A, b
