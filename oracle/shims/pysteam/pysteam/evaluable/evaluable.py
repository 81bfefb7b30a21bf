"""Evaluator-graph core: forward() builds a Node tree, backward() pushes a left-hand matrix
down to the state variables and accumulates Jacobians (STEAM's reverse-mode scheme).
Serves reference pysteam_augmented/extra_se3_evaluators.py:3 and mpc.py:12."""
import abc
import numpy as np


class Node:
    def __init__(self, value, *children):
        self.value = value
        self.children = list(children)


class Jacobians:
    def __init__(self):
        self._jacs = {}

    def add(self, key, jac):
        if key in self._jacs:
            self._jacs[key] = self._jacs[key] + jac
        else:
            self._jacs[key] = jac

    def clear(self):
        self._jacs.clear()

    def get_copy(self):
        return dict(self._jacs)


class Evaluable(abc.ABC):
    def evaluate(self, lhs=None, jacs=None):
        end_node = self.forward()
        if lhs is not None:
            assert jacs is not None
            self.backward(lhs, end_node, jacs)
        return end_node.value

    @property
    @abc.abstractmethod
    def active(self):
        ...

    @property
    @abc.abstractmethod
    def related_var_keys(self):
        ...

    @abc.abstractmethod
    def forward(self):
        ...

    @abc.abstractmethod
    def backward(self, lhs, node, jacs):
        ...
