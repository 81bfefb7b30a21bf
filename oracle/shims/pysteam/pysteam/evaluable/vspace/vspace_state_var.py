"""Vector-space variable, additive update (mpc.py:11,46-49,70,200,205; lev_marq...py:57)."""
import copy
import numpy as np
from ..state_var import StateVar
from ..evaluable import Node


class VSpaceStateVar(StateVar):
    def __init__(self, value, **kwargs):
        value = np.array(value, dtype=float)
        super().__init__(value.shape[0], **kwargs)
        self._value = value

    def forward(self):
        return Node(self._value)

    def clone(self):
        c = copy.copy(self)
        c._value = self._value.copy()
        return c

    @property
    def value(self):
        return self._value

    def assign(self, value):
        self._value = np.array(value, dtype=float)

    def update(self, perturbation):
        self._value = self._value + perturbation
