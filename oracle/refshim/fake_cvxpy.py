"""Minimal stand-in for the slice of `cvxpy` the reference's QP controllers use
(position_control/cbf_qp.py:47-106,190; optimal_decay_cbf_qp.py:56-130,156).
TEST INFRASTRUCTURE (see refshim/__init__.py).

Expressions are lazy trees evaluated numerically; because every constraint is
affine and the objective a weighted sum of squares of affine maps, (P, q, G, h)
are recovered by probing the trees at 0 and at unit vectors each solve (so
in-place Parameter updates such as `self.A1.value[row, :] = ...` are honoured).
The resulting strictly convex QP goes to oracle/qp_exact.py.
"""
import numpy as np

from ..qp_exact import solve_qp_exact, OPTIMAL

GUROBI = "GUROBI"
OSQP = "OSQP"
SCS = "SCS"


class Expr:
    __array_ufunc__ = None       # make numpy scalars defer to our operators
    __array_priority__ = 1000

    def ev(self, env):
        raise NotImplementedError

    # arithmetic -----------------------------------------------------------
    def __add__(self, o): return _Bin(np.add, self, o)
    def __radd__(self, o): return _Bin(np.add, o, self)
    def __sub__(self, o): return _Bin(np.subtract, self, o)
    def __rsub__(self, o): return _Bin(np.subtract, o, self)
    def __mul__(self, o): return _Bin(np.multiply, self, o)
    def __rmul__(self, o): return _Bin(np.multiply, o, self)
    def __matmul__(self, o): return _Bin(np.matmul, self, o)
    def __rmatmul__(self, o): return _Bin(np.matmul, o, self)
    def __neg__(self): return _Bin(np.multiply, -1.0, self)
    def __getitem__(self, idx): return _Index(self, idx)
    # relations (canonical form: expr <= 0) ------------------------------
    def __le__(self, o): return _constraint_le(self, o)
    def __ge__(self, o): return _constraint_le(o, self)


def _ev(x, env):
    if isinstance(x, Expr):
        return x.ev(env)
    return np.asarray(x, dtype=float)


class _Bin(Expr):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b

    def ev(self, env):
        return self.op(_ev(self.a, env), _ev(self.b, env))


class _Index(Expr):
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def ev(self, env):
        return np.atleast_1d(_ev(self.a, env)[self.idx])


class Parameter(Expr):
    def __init__(self, shape=(), value=None, **kw):
        self.shape = shape
        self.value = None if value is None else np.array(value, dtype=float)

    def ev(self, env):
        return np.asarray(self.value, dtype=float)


class Variable(Expr):
    def __init__(self, shape=(), **kw):
        self.shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.value = None

    def ev(self, env):
        return env[id(self)]


class _Abs(Expr):
    def __init__(self, a):
        self.a = a

    def __le__(self, o):      # |a| <= c   ->   a <= c  and  -a <= c
        return [_constraint_le(self.a, o), _constraint_le(-self.a, o)]


def abs(a):  # noqa: A001  (mirrors cvxpy.abs)
    return _Abs(a)


class _Quad:
    """sum_i w_i * ||e_i||^2"""
    def __init__(self, terms):
        self.terms = terms

    def __add__(self, o): return _Quad(self.terms + o.terms)
    def __mul__(self, c): return _Quad([(w * float(c), e) for w, e in self.terms])
    __rmul__ = __mul__


def sum_squares(e): return _Quad([(1.0, e)])
def square(e): return _Quad([(1.0, e)])


class Minimize:
    def __init__(self, quad):
        self.quad = quad


class _LeZero:
    def __init__(self, e):
        self.e = e


def _constraint_le(a, b):
    a_ = a if isinstance(a, Expr) else np.asarray(a, dtype=float)
    return _LeZero(_Bin(np.subtract, a_, b))


def _variables(node, acc):
    if isinstance(node, Variable):
        if all(node is not v for v in acc):
            acc.append(node)
    elif isinstance(node, _Bin):
        _variables(node.a, acc); _variables(node.b, acc)
    elif isinstance(node, (_Index, _Abs)):
        _variables(node.a, acc)
    elif isinstance(node, _LeZero):
        _variables(node.e, acc)


class Problem:
    def __init__(self, objective, constraints):
        self.objective = objective
        flat = []
        for c in constraints:
            flat.extend(c if isinstance(c, list) else [c])
        self.constraints = flat
        self.vars = []
        for _, e in objective.quad.terms:
            _variables(e, self.vars)
        for c in self.constraints:
            _variables(c, self.vars)
        self.status = None
        self.value = None

    def _affine(self, e):
        """-> (J, c) with e(x) = J x + c, by probing."""
        n = sum(v.size for v in self.vars)

        def at(x):
            env, k = {}, 0
            for v in self.vars:
                env[id(v)] = x[k:k + v.size].reshape(v.shape)
                k += v.size
            return np.asarray(e.ev(env), dtype=float).reshape(-1)

        c = at(np.zeros(n))
        J = np.stack([at(np.eye(n)[j]) - c for j in range(n)], axis=1)
        return J, c

    def solve(self, solver=None, **kw):
        n = sum(v.size for v in self.vars)
        P = np.zeros((n, n)); q = np.zeros(n)
        for w, e in self.objective.quad.terms:
            J, c = self._affine(e)
            P += 2.0 * w * J.T @ J
            q += 2.0 * w * J.T @ c
        Gs, hs = [], []
        for con in self.constraints:
            J, c = self._affine(con.e)      # J x + c <= 0
            Gs.append(J); hs.append(-c)
        G = np.vstack(Gs); h = np.concatenate(hs)
        res = solve_qp_exact(P, q, G, h)
        self.qp = dict(P=P, q=q, G=G, h=h, res=res)
        if res["status"] == OPTIMAL:
            self.status = "optimal"
            k = 0
            for v in self.vars:
                v.value = res["x"][k:k + v.size].reshape(v.shape)
                k += v.size
            self.value = res["obj"]
        else:
            self.status = "infeasible"
            for v in self.vars:
                v.value = None
            self.value = np.inf
        return self.value
