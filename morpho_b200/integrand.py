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
       "atan": 19, "atan2": 20, "sinh": 21, "cosh": 22, "tanh": 23, "floor": 24, "ceil": 25, "log10": 26}
BINARY = ("add", "sub", "mul", "div", "pow", "atan2")

OP_DTYPE = np.dtype([("op", "<i4"), ("a", "<i4"), ("c", "<f8")])

__all__ = ["Expr", "trace", "sqrt", "sin", "cos", "exp", "log", "tan", "asin", "acos", "atan", "atan2", "sinh", "cosh",
           "tanh", "floor", "ceil", "log10", "OPS", "OP_DTYPE"]


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


def trace(fn, dim, field_sizes):
    """Call ``fn`` with symbolic arguments and return its program as a structured numpy array."""
    x = tuple(Expr([(OPS["pos"], k, 0.0)]) for k in range(dim))
    args, q = [], 0
    for ps in field_sizes:
        comps = tuple(Expr([(OPS["field"], q + k, 0.0)]) for k in range(ps))
        q += ps
        args.append(comps[0] if ps == 1 else comps)
    out = Expr.lift(fn(x, *args))
    prog = np.zeros(len(out.prog), dtype=OP_DTYPE)
    for i, t in enumerate(out.prog):
        prog[i] = t
    return prog
