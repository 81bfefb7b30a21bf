"""SE(3) evaluators imported at reference mpc.py:13. Jacobian conventions follow STEAM:
left perturbations, lhs matrices pushed right-to-left through the chain."""
import numpy as np
from pylgmath import Transformation, se3op
from ..evaluable import Evaluable, Node
import os

# [3P-UNVERIFIED] switch: STEAM's SE3ErrorEvaluator back-propagates with -J_l^{-1}(e); the exact
# derivative of ln(T_meas * (exp(d^) T)^-1) is -J_r^{-1}(e). Default follows STEAM.
POSE_ERROR_EXACT_JACOBIAN = os.environ.get("DRE_ORACLE_POSE_JAC", "steam") == "exact"


class _Unary(Evaluable):
    def __init__(self, child):
        self._c = child

    @property
    def active(self):
        return self._c.active

    @property
    def related_var_keys(self):
        return self._c.related_var_keys


class _Binary(Evaluable):
    def __init__(self, a, b):
        self._a, self._b = a, b

    @property
    def active(self):
        return self._a.active or self._b.active

    @property
    def related_var_keys(self):
        return self._a.related_var_keys | self._b.related_var_keys


class LogMapEvaluator(_Unary):
    def forward(self):
        ch = self._c.forward()
        return Node(ch.value.vec(), ch)

    def backward(self, lhs, node, jacs):
        if self._c.active:
            self._c.backward(lhs @ se3op.vec2jacinv(node.value), node.children[0], jacs)


class ExpMapEvaluator(_Unary):
    def forward(self):
        ch = self._c.forward()
        return Node(Transformation(xi_ab=ch.value), ch)

    def backward(self, lhs, node, jacs):
        if self._c.active:
            self._c.backward(lhs @ se3op.vec2jac(node.children[0].value), node.children[0], jacs)


class InverseEvaluator(_Unary):
    def forward(self):
        ch = self._c.forward()
        return Node(ch.value.inverse(), ch)

    def backward(self, lhs, node, jacs):
        if self._c.active:
            self._c.backward(-lhs @ node.value.adjoint(), node.children[0], jacs)


class ComposeEvaluator(_Binary):
    def forward(self):
        a, b = self._a.forward(), self._b.forward()
        return Node(a.value @ b.value, a, b)

    def backward(self, lhs, node, jacs):
        if self._a.active:
            self._a.backward(lhs, node.children[0], jacs)
        if self._b.active:
            self._b.backward(lhs @ node.children[0].value.adjoint(), node.children[1], jacs)


class ComposeInverseEvaluator(_Binary):
    """T_a * T_b^{-1}"""

    def forward(self):
        a, b = self._a.forward(), self._b.forward()
        return Node(a.value @ b.value.inverse(), a, b)

    def backward(self, lhs, node, jacs):
        if self._a.active:
            self._a.backward(lhs, node.children[0], jacs)
        if self._b.active:
            self._b.backward(-lhs @ node.value.adjoint(), node.children[1], jacs)


class SE3ErrorEvaluator(Evaluable):
    """e = ln(T_meas * T^{-1})"""

    def __init__(self, T_ab, T_ab_meas):
        self._T, self._meas = T_ab, T_ab_meas

    @property
    def active(self):
        return self._T.active

    @property
    def related_var_keys(self):
        return self._T.related_var_keys

    def forward(self):
        ch = self._T.forward()
        return Node((self._meas @ ch.value.inverse()).vec(), ch)

    def backward(self, lhs, node, jacs):
        if self._T.active:
            Jinv = se3op.vec2jacinv(node.value)
            if POSE_ERROR_EXACT_JACOBIAN:
                Jinv = Jinv @ se3op.tranAd(se3op.vec2tran(node.value))
            self._T.backward(-lhs @ Jinv, node.children[0], jacs)
