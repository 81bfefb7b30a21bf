from .solver import Solver
from .gauss_newton_solver import GaussNewtonSolver
