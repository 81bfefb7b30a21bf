"""TEST INFRASTRUCTURE -- stand-in for `asrl-pylgmath==1.0.3` (reference environment.yml:41).

Serves `from pylgmath import Transformation` at reference
environment/path_tracking/utils/mpc.py:7 (used at mpc.py:78-83, 228-235) and the
SE(3) operations the pysteam stand-in needs. The real package is absent from this
container; this restates the published closed forms (Barfoot, "State Estimation for
Robotics", ch. 7: exp/log maps, left Jacobian, Q matrix, adjoint) with xi = (rho; phi),
translation first. Small-angle cases use series instead of the cancellation-prone closed
forms. PARITY UNPINNED.
"""
import numpy as np

__all__ = ["Transformation", "se3op", "so3op"]


class so3op:
    @staticmethod
    def hat(v):
        v = np.asarray(v, dtype=float).reshape(3)
        return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])

    @staticmethod
    def vec2rot(phi):
        phi = np.asarray(phi, dtype=float).reshape(3)
        ang = np.linalg.norm(phi)
        if ang < 1e-12:
            return np.eye(3) + so3op.hat(phi)
        a = phi / ang
        c, s = np.cos(ang), np.sin(ang)
        return c * np.eye(3) + (1.0 - c) * np.outer(a, a) + s * so3op.hat(a)

    @staticmethod
    def rot2vec(C):
        # angle from the trace, axis from the skew part (general case); near 0 -> zeros;
        # near pi -> eigenvector of eigenvalue +1.
        tr = np.clip(0.5 * (np.trace(C) - 1.0), -1.0, 1.0)
        ang = np.arccos(tr)
        s = np.sin(ang)
        if abs(s) > 1e-9:
            axis = np.array([C[2, 1] - C[1, 2], C[0, 2] - C[2, 0], C[1, 0] - C[0, 1]])
            return (0.5 * ang / s) * axis
        if abs(ang) > 1e-9:
            w, v = np.linalg.eigh(0.5 * (C + C.T))
            k = int(np.argmin(np.abs(w - 1.0)))
            return ang * v[:, k]
        return np.zeros(3)

    @staticmethod
    def _coeffs(ang):
        """(sin a / a, (1-cos a)/a^2, (a - sin a)/a^3) with series for small a."""
        if ang < 1e-4:
            a2 = ang * ang
            return (1.0 - a2 / 6.0 + a2 * a2 / 120.0, 0.5 - a2 / 24.0 + a2 * a2 / 720.0,
                    1.0 / 6.0 - a2 / 120.0 + a2 * a2 / 5040.0)
        s, c = np.sin(ang), np.cos(ang)
        return s / ang, (1.0 - c) / (ang * ang), (ang - s) / (ang ** 3)

    @staticmethod
    def vec2jac(phi):
        phi = np.asarray(phi, dtype=float).reshape(3)
        ang = np.linalg.norm(phi)
        _, b, c = so3op._coeffs(ang)
        P = so3op.hat(phi)
        return np.eye(3) + b * P + c * (P @ P)

    @staticmethod
    def vec2jacinv(phi):
        phi = np.asarray(phi, dtype=float).reshape(3)
        ang = np.linalg.norm(phi)
        P = so3op.hat(phi)
        if ang < 1e-4:
            k = 1.0 / 12.0 + ang * ang / 720.0
        else:
            half = 0.5 * ang
            k = (1.0 - half / np.tan(half)) / (ang * ang)
        return np.eye(3) - 0.5 * P + k * (P @ P)


class se3op:
    @staticmethod
    def _split(xi):
        xi = np.asarray(xi, dtype=float).reshape(6)
        return xi[:3], xi[3:]

    @staticmethod
    def vec2tran(xi):
        rho, phi = se3op._split(xi)
        T = np.eye(4)
        T[:3, :3] = so3op.vec2rot(phi)
        T[:3, 3] = so3op.vec2jac(phi) @ rho
        return T

    @staticmethod
    def tran2vec(T):
        phi = so3op.rot2vec(T[:3, :3])
        rho = so3op.vec2jacinv(phi) @ T[:3, 3]
        return np.concatenate([rho, phi]).reshape(6, 1)

    @staticmethod
    def tranAd(T):
        C, r = T[:3, :3], T[:3, 3]
        Ad = np.zeros((6, 6))
        Ad[:3, :3] = C
        Ad[:3, 3:] = so3op.hat(r) @ C
        Ad[3:, 3:] = C
        return Ad

    @staticmethod
    def curlyhat(xi):
        rho, phi = se3op._split(xi)
        M = np.zeros((6, 6))
        M[:3, :3] = so3op.hat(phi)
        M[:3, 3:] = so3op.hat(rho)
        M[3:, 3:] = so3op.hat(phi)
        return M

    @staticmethod
    def vec2Q(xi):
        """Barfoot eq. 7.86b; coefficients by series when the angle is small."""
        rho, phi = se3op._split(xi)
        ang = np.linalg.norm(phi)
        rx, px = so3op.hat(rho), so3op.hat(phi)
        a2 = ang * ang
        if ang < 1e-2:
            m2 = 1.0 / 6.0 - a2 / 120.0 + a2 * a2 / 5040.0
            m3 = 1.0 / 24.0 - a2 / 720.0 + a2 * a2 / 40320.0
            m4 = 1.0 / 120.0 - a2 / 2520.0 + a2 * a2 / 120960.0
        else:
            s, c = np.sin(ang), np.cos(ang)
            m2 = (ang - s) / ang ** 3
            m3 = (0.5 * a2 + c - 1.0) / ang ** 4
            m4 = (ang - 1.5 * s + 0.5 * ang * c) / ang ** 5
        pxrx, rxpx = px @ rx, rx @ px
        pxrxpx = pxrx @ px
        t2 = pxrx + rxpx + pxrxpx
        t3 = px @ pxrx + rxpx @ px - 3.0 * pxrxpx
        t4 = pxrxpx @ px + px @ pxrxpx
        return 0.5 * rx + m2 * t2 + m3 * t3 + m4 * t4

    @staticmethod
    def vec2jac(xi):
        rho, phi = se3op._split(xi)
        J = so3op.vec2jac(phi)
        out = np.zeros((6, 6))
        out[:3, :3] = J
        out[:3, 3:] = se3op.vec2Q(xi)
        out[3:, 3:] = J
        return out

    @staticmethod
    def vec2jacinv(xi):
        rho, phi = se3op._split(xi)
        Ji = so3op.vec2jacinv(phi)
        out = np.zeros((6, 6))
        out[:3, :3] = Ji
        out[:3, 3:] = -Ji @ se3op.vec2Q(xi) @ Ji
        out[3:, 3:] = Ji
        return out


class Transformation:
    """SE(3) element stored as a 4x4 matrix T_ba."""

    def __init__(self, *, T_ba=None, xi_ab=None, C_ba=None, r_ab_inb=None, transformation=None):
        if transformation is not None:
            self._T = transformation._T.copy()
        elif T_ba is not None:
            self._T = np.array(T_ba, dtype=float).reshape(4, 4).copy()
        elif xi_ab is not None:
            self._T = se3op.vec2tran(xi_ab)
        elif C_ba is not None:
            self._T = np.eye(4)
            self._T[:3, :3] = C_ba
            if r_ab_inb is not None:
                self._T[:3, 3] = np.asarray(r_ab_inb).reshape(3)
        else:
            self._T = np.eye(4)

    def matrix(self):
        return self._T.copy()

    def C_ba(self):
        return self._T[:3, :3].copy()

    def r_ab_inb(self):
        return self._T[:3, 3:4].copy()

    def vec(self):
        return se3op.tran2vec(self._T)

    def inverse(self):
        C, r = self._T[:3, :3], self._T[:3, 3]
        Ti = np.eye(4)
        Ti[:3, :3] = C.T
        Ti[:3, 3] = -C.T @ r
        return Transformation(T_ba=Ti)

    def adjoint(self):
        return se3op.tranAd(self._T)

    def __matmul__(self, other):
        if isinstance(other, Transformation):
            return Transformation(T_ba=self._T @ other._T)
        return self._T @ other

    def __deepcopy__(self, memo):
        return Transformation(T_ba=self._T)
