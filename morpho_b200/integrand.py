"""Integrand expressions for LineIntegral / AreaIntegral.

The reference evaluates a Morpho closure ``fn(x, q1, q2, ...)`` at every quadrature point by re-entering the VM
(integral_integrandfn, functional.c:5146-5180).  The device library has no VM; it runs a postfix program of
``mcu_expr_op`` (include/morpho_cuda.h).  This module records such a program by calling a Python callable once with
symbolic arguments: ``x`` is a tuple of :class:`Expr` (one per space dimension), each field value an :class:`Expr`
(scalar fields) or a tuple of them (vector / matrix fields, column-major like the reference's Matrix prototype).
"""
from __future__ import annotations

import numpy as np

OPS = {"const": 0, "pos": 1, "field": 2, "add": 3, "sub": 4, "mul": 5, "div": 6, "neg": 7, "powi": 8, "sqrt": 9,
       "sin": 10, "cos": 11, "exp": 12, "log": 13, "abs": 14, "pow": 15, "tan": 16, "asin": 17, "acos": 18,
       "atan": 19, "atan2": 20, "sinh": 21, "cosh": 22, "tanh": 23, "floor": 24, "ceil": 25, "log10": 26,
       "tangent": 27, "normal": 28, "gradq": 29, "lt": 30, "le": 31, "eq": 32, "ne": 33, "not": 34, "select": 35, "store": 36, "load": 37}
BINARY = ("add", "sub", "mul", "div", "pow", "atan2", "lt", "le", "eq", "ne")

OP_DTYPE = np.dtype([("op", "<i4"), ("a", "<i4"), ("c", "<f8")])

__all__ = ["Expr", "trace", "sqrt", "sin", "cos", "exp", "log", "tan", "asin", "acos", "atan", "atan2", "sinh", "cosh",
           "tanh", "floor", "ceil", "log10", "tangent", "normal", "grad", "trace_point", "OPS", "OP_DTYPE", "where", "lt", "le"]


class Expr:
    """A node of the integrand; holds its own postfix program."""
    __slots__ = ("prog",)

    def __init__(self, prog):
        self.prog = prog

    @staticmethod
    def lift(v):
        if isinstance(v, Expr):
            return v
        return Expr([(OPS["const"], 0, float(v))])

    def _bin(self, other, op, swap=False):
        o = Expr.lift(other)
        a, b = (o, self) if swap else (self, o)
        return Expr(a.prog + b.prog + [(OPS[op], 0, 0.0)])

    def _un(self, op, a=0):
        return Expr(self.prog + [(OPS[op], a, 0.0)])

    def __add__(self, o): return self._bin(o, "add")
    def __radd__(self, o): return self._bin(o, "add", True)
    def __sub__(self, o): return self._bin(o, "sub")
    def __rsub__(self, o): return self._bin(o, "sub", True)
    def __mul__(self, o): return self._bin(o, "mul")
    def __rmul__(self, o): return self._bin(o, "mul", True)
    def __truediv__(self, o): return self._bin(o, "div")
    def __rtruediv__(self, o): return self._bin(o, "div", True)
    def __neg__(self): return self._un("neg")
    def __pos__(self): return self
    def __abs__(self): return self._un("abs")

    def __pow__(self, n):
        if isinstance(n, (int, np.integer)):
            return self._un("powi", int(n))
        return self._bin(n, "pow")

    def __rpow__(self, base):
        return self._bin(base, "pow", True)


def _fn(name):
    def f(x):
        return Expr.lift(x)._un(name)
    f.__name__ = name
    return f


sqrt, sin, cos, exp, log = _fn("sqrt"), _fn("sin"), _fn("cos"), _fn("exp"), _fn("log")
tan, asin, acos, atan, sinh, cosh, tanh = (_fn(n) for n in ("tan", "asin", "acos", "atan", "sinh", "cosh", "tanh"))
floor, ceil, log10 = _fn("floor"), _fn("ceil"), _fn("log10")


def atan2(y, x):
    return Expr.lift(y)._bin(x, "atan2")


def lt(a, b):
    """1.0 where a < b, else 0.0 — with the VM's tolerance for equal floats (include/morpho_cuda.h, MCU_X_LT)."""
    return Expr.lift(a)._bin(b, "lt")


def le(a, b):
    return Expr.lift(a)._bin(b, "le")


def where(cond, a, b):
    """``a`` where ``cond`` is non-zero, else ``b`` (MCU_X_SELECT): what an ``if`` in a Morpho closure translates to."""
    c, x, y = Expr.lift(cond), Expr.lift(a), Expr.lift(b)
    return Expr(c.prog + x.prog + y.prog + [(OPS["select"], 0, 0.0)])


class _Q(Expr):
    """The interpolated value of a scalar Field argument (what ``grad`` recognises)."""
    __slots__ = ("qbase",)


class _QVec(tuple):
    """The interpolated value of a vector / matrix Field argument."""
    qbase = 0


def tangent():
    """Unit tangent of the line element (the reference's ``tangent()``, functional.c:4402-4425)."""
    return tuple(Expr([(OPS["tangent"], k, 0.0)]) for k in range(3))


def normal():
    """Unit normal of the triangle (the reference's ``normal()``, functional.c:4444-4470)."""
    return tuple(Expr([(OPS["normal"], k, 0.0)]) for k in range(3))


def grad(q):
    """P1 gradient of a Field argument on the triangle (``grad(q)``, functional.c:4638-4745): for a scalar field a
    tuple (d/dx0, d/dx1, d/dx2); for a vector field a list over the directions of tuples over the components."""
    if isinstance(q, _Q):
        return tuple(Expr([(OPS["gradq"], 3 * q.qbase + k, 0.0)]) for k in range(3))
    if isinstance(q, _QVec):
        return [tuple(Expr([(OPS["gradq"], 3 * (q.qbase + c) + k, 0.0)]) for c in range(len(q))) for k in range(3)]
    raise TypeError("grad() takes a Field argument of the integrand")


def trace(fn, dim, field_sizes):
    """Call ``fn`` with symbolic arguments and return its program as a structured numpy array."""
    x = tuple(Expr([(OPS["pos"], k, 0.0)]) for k in range(dim))
    args, q = [], 0
    for ps in field_sizes:
        if ps == 1:
            a = _Q([(OPS["field"], q, 0.0)])
        else:
            a = _QVec(Expr([(OPS["field"], q + k, 0.0)]) for k in range(ps))
        a.qbase = q
        q += ps
        args.append(a)
    out = Expr.lift(fn(x, *args))
    prog = np.zeros(len(out.prog), dtype=OP_DTYPE)
    for i, t in enumerate(out.prog):
        prog[i] = t
    return prog


def trace_point(fn, dim):
    """Programs of ``fn(x, y[, z])`` (ScalarPotential): one if it returns a number, ``dim`` if it returns a sequence."""
    out = fn(*(Expr([(OPS["pos"], k, 0.0)]) for k in range(dim)))
    outs = list(out) if isinstance(out, (tuple, list)) else [out]
    progs = []
    for o in outs:
        e = Expr.lift(o)
        prog = np.zeros(len(e.prog), dtype=OP_DTYPE)
        for i, t in enumerate(e.prog):
            prog[i] = t
        progs.append(prog)
    return progs
