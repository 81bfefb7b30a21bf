"""State variables: unique key, perturbation dimension, lock flag (mpc.py:84 sets `.locked`)."""
import abc
import itertools
from .evaluable import Evaluable, Node

_key_counter = itertools.count()


class StateVar(Evaluable):
    def __init__(self, perturb_dim, *, name="", locked=False):
        self._perturb_dim = perturb_dim
        self._name = name
        self._locked = locked
        self._key = next(_key_counter)

    @property
    def key(self):
        return self._key

    @property
    def perturb_dim(self):
        return self._perturb_dim

    @property
    def locked(self):
        return self._locked

    @locked.setter
    def locked(self, v):
        self._locked = v

    @property
    def active(self):
        return not self._locked

    @property
    def related_var_keys(self):
        return {self._key}

    @abc.abstractmethod
    def clone(self):
        ...

    @abc.abstractmethod
    def assign(self, value):
        ...

    @abc.abstractmethod
    def update(self, perturbation):
        ...

    def backward(self, lhs, node, jacs):
        if self.active:
            jacs.add(self._key, lhs)
