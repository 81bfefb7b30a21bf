"""SE(3) pose variable with left-multiplicative update T <- exp(delta^) T (mpc.py:10,83,199-204,233)."""
import copy
from pylgmath import Transformation
from ..state_var import StateVar
from ..evaluable import Node


class SE3StateVar(StateVar):
    def __init__(self, value, **kwargs):
        super().__init__(6, **kwargs)
        self._value = value

    def forward(self):
        return Node(self._value)

    def clone(self):
        c = copy.copy(self)
        c._value = Transformation(T_ba=self._value.matrix())
        return c

    @property
    def value(self):
        return self._value

    def assign(self, value):
        self._value = Transformation(T_ba=value.matrix())

    def update(self, perturbation):
        self._value = Transformation(xi_ab=perturbation) @ self._value
