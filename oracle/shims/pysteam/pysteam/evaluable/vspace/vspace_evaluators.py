"""Vector-space evaluators imported at reference mpc.py:15."""
import numpy as np
from ..evaluable import Evaluable, Node
from ..se3.se3_evaluators import _Unary, _Binary


class NegationEvaluator(_Unary):
    def forward(self):
        ch = self._c.forward()
        return Node(-ch.value, ch)

    def backward(self, lhs, node, jacs):
        if self._c.active:
            self._c.backward(-lhs, node.children[0], jacs)


class AdditionEvaluator(_Binary):
    def forward(self):
        a, b = self._a.forward(), self._b.forward()
        return Node(a.value + b.value, a, b)

    def backward(self, lhs, node, jacs):
        if self._a.active:
            self._a.backward(lhs, node.children[0], jacs)
        if self._b.active:
            self._b.backward(lhs, node.children[1], jacs)


class ScalarMultEvaluator(_Unary):
    def __init__(self, child, scalar):
        super().__init__(child)
        self._s = scalar

    def forward(self):
        ch = self._c.forward()
        return Node(self._s * ch.value, ch)

    def backward(self, lhs, node, jacs):
        if self._c.active:
            self._c.backward(self._s * lhs, node.children[0], jacs)


class MatrixMultEvaluator(_Unary):
    def __init__(self, child, matrix):
        super().__init__(child)
        self._m = np.asarray(matrix, dtype=float)

    def forward(self):
        ch = self._c.forward()
        return Node(self._m @ ch.value, ch)

    def backward(self, lhs, node, jacs):
        if self._c.active:
            self._c.backward(lhs @ self._m, node.children[0], jacs)


class VSpaceErrorEvaluator(_Unary):
    def __init__(self, child, meas):
        super().__init__(child)
        self._meas = np.asarray(meas, dtype=float)

    def forward(self):
        ch = self._c.forward()
        return Node(self._meas - ch.value, ch)

    def backward(self, lhs, node, jacs):
        if self._c.active:
            self._c.backward(-lhs, node.children[0], jacs)
