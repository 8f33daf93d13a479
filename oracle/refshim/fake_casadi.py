"""Numeric stand-in for the handful of `casadi` symbols the reference's model
files touch (test infrastructure; see refshim/__init__.py).

Everything is evaluated eagerly on float64 numpy values, so the reference's
symbolic code paths (`f(X, casadi=True)`, `step(..., casadi=True)`,
`agent_barrier_dt`) become ordinary numeric functions.  One semantic
difference is worth recording: with real CasADi, `angle_normalize` takes the
`ca.fmod` branch (C fmod, e.g. robots/dynamic_unicycle2D.py:17-19); with
numeric inputs it takes the numpy floored-`%` branch (:14-16).  The wrapped
angle only ever enters the discrete barriers through cos/sin (DU/KB) or not
at all (Quad3D uses x,y only), so barrier values are unaffected.
"""
import numpy as np

pi = np.pi


class _Sym:  # placeholder types so `isinstance(x, (ca.SX, ca.MX, ca.DM))` works
    @staticmethod
    def zeros(r, c=1):
        return np.zeros((r, c))


class SX(_Sym):
    pass


class MX(_Sym):
    pass


class DM(_Sym):
    """`ca.DM([...])` -> a plain float64 ndarray; still a type for isinstance checks (double_integrator2D.py:84)."""

    def __new__(cls, x):
        return np.array(x, dtype=float)


def vertcat(*args):
    arrs = [np.asarray(a, dtype=float) for a in args]
    if arrs and all(a.ndim == 2 and a.shape[0] == 1 and a.shape[1] > 1 for a in arrs):
        return np.vstack(arrs)                     # rows from horzcat -> matrix (kinematic_bicycle2D_dpcbf.py:114-117)
    parts = [np.atleast_1d(a).reshape(-1) for a in arrs]
    return np.concatenate(parts).reshape(-1, 1)


def horzcat(*args):
    parts = [np.atleast_1d(np.asarray(a, dtype=float)).reshape(-1) for a in args]
    return np.concatenate(parts).reshape(1, -1)


def mtimes(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    out = np.asarray(args[0], dtype=float)
    for a in args[1:]:
        out = out @ np.asarray(a, dtype=float)
    return out


cos = np.cos
sin = np.sin
tan = np.tan
sqrt = np.sqrt
fabs = np.abs
fmax = np.maximum
fmin = np.minimum
fmod = np.fmod
atan2 = np.arctan2
hypot = np.hypot
exp = np.exp
log = np.log


def power(a, b):
    return np.power(a, b)


def if_else(c, a, b):
    return np.where(c, a, b)


def norm_2(x):
    return np.linalg.norm(np.asarray(x, dtype=float))
