"""TEST INFRASTRUCTURE -- a stand-in for Theano 0.6 so the reference's own Python can run here.

The reference (slinderman/theano_pyglm) builds every quantity on the hot path as a Theano graph
(`glm.py:28-63`, `impulse.py:45-58`, `priors.py`, ...) and evaluates it with `theano.function`
(`utils/theano_func_wrapper.py:12-51`).  Theano is not installable in this image.  This module
provides just enough of the `theano` / `theano.tensor` API, as lazily evaluated expression nodes
executed with torch float64 (and `torch.autograd` for `T.grad`), for `oracle/ref_loader.py` to import
the UNMODIFIED reference sources from /root/reference and run them: the reference's graph
construction code decides what is computed, torch only supplies the float64 arithmetic.

What this does NOT reproduce: Theano's graph optimiser (e.g. its `log(1+exp(x)) -> softplus`
rewrite), its reduction order and its elementwise fusion -- the expression is evaluated literally as
written in the reference source.  Only `oracle/ref_fixtures.py` (the golden-vector generator) uses
this; nothing under `theano_pyglm_b200/` may import it.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

_F64 = torch.float64


def _to_tensor(v):
    if isinstance(v, torch.Tensor):
        return v
    if isinstance(v, (bool, int, float, np.number)):
        if isinstance(v, (float, np.floating)):
            return torch.tensor(float(v), dtype=_F64)
        return torch.tensor(int(v))
    a = np.asarray(v)
    if a.dtype.kind == "f":
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
    if a.dtype.kind in "iub":
        return torch.as_tensor(np.ascontiguousarray(a).astype(np.int64))
    raise TypeError("cannot convert %r to a tensor" % (type(v),))


class _Ctx:
    """One evaluation: leaf bindings (function inputs / givens) and a memo of computed nodes."""

    def __init__(self, feeds):
        self.feeds = feeds          # id(Var) -> tensor
        self.memo = {}


class Var:
    __array_ufunc__ = None          # numpy defers to our reflected operators
    __array_priority__ = 1000

    def __init__(self, kind, fn=None, args=(), name=None, dtype="float64", ndim=None, value=None):
        self.kind = kind            # 'input' | 'shared' | 'const' | 'op' | 'grad'
        self.fn = fn
        self.args = args
        self.name = name
        self.dtype = dtype
        self.ndim = ndim
        self.value = value

    # ---- identity
    def __str__(self):
        return self.name if self.name is not None else "<%s>" % self.kind

    __repr__ = __str__

    def __hash__(self):
        return id(self)

    def __eq__(self, other):        # identity, as Theano variables hash/compare
        return self is other

    def __ne__(self, other):
        return self is not other

    # ---- shared-variable API
    def get_value(self, borrow=False):
        return self.value

    def set_value(self, v, borrow=False):
        self.value = v

    # ---- evaluation
    def _eval(self, ctx):
        key = id(self)
        if key in ctx.feeds:
            return ctx.feeds[key]
        if key in ctx.memo:
            return ctx.memo[key]
        if self.kind == "input":
            raise KeyError("no value bound to input %s" % self)
        if self.kind in ("shared", "const"):
            out = _to_tensor(self.value)
        elif self.kind == "op":
            out = self.fn(*[_ev(a, ctx) for a in self.args])
        elif self.kind == "grad":
            out = self.fn(ctx)
        else:
            raise RuntimeError(self.kind)
        ctx.memo[key] = out
        return out

    # ---- operators
    def __add__(self, o): return _op(torch.add, self, o)
    def __radd__(self, o): return _op(torch.add, o, self)
    def __sub__(self, o): return _op(torch.sub, self, o)
    def __rsub__(self, o): return _op(torch.sub, o, self)
    def __mul__(self, o): return _op(torch.mul, self, o)
    def __rmul__(self, o): return _op(torch.mul, o, self)
    def __truediv__(self, o): return _op(torch.true_divide, self, o)
    def __rtruediv__(self, o): return _op(torch.true_divide, o, self)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __pow__(self, o): return _op(torch.pow, self, o)
    def __rpow__(self, o): return _op(torch.pow, o, self)
    def __neg__(self): return _op(torch.neg, self)
    def __abs__(self): return _op(torch.abs, self)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        dyn = [i for i in idx if isinstance(i, Var)]

        def f(x, *dv):
            it = iter(dv)
            real = []
            for i in idx:
                if isinstance(i, Var):
                    v = next(it)
                    real.append(v.long() if v.dim() else int(v))
                elif isinstance(i, np.ndarray):
                    real.append(torch.as_tensor(i.astype(np.int64)))
                else:
                    real.append(i)
            return x[tuple(real)]
        return Var("op", f, (self,) + tuple(dyn))

    # ---- tensor methods the reference calls
    def sum(self, axis=None): return sum(self, axis=axis)
    def flatten(self): return flatten(self)
    def reshape(self, shape): return reshape(self, shape)
    def take(self, indices): return _op(lambda x, i: x.reshape(-1)[torch.as_tensor(np.asarray(i), dtype=torch.long)], self, _Raw(indices))
    def dimshuffle(self, *pattern):
        def f(x):
            perm = [p for p in pattern if p != 'x']
            y = x.permute(*perm) if perm else x
            for pos, p in enumerate(pattern):
                if p == 'x':
                    y = y.unsqueeze(pos)
            return y
        return Var("op", f, (self,))

    @property
    def T(self): return transpose(self)

    @property
    def shape(self): return _ShapeOf(self)


class _ShapeOf:
    def __init__(self, v): self.v = v
    def __getitem__(self, i): return Var("op", lambda x: torch.tensor(x.shape[i]), (self.v,), dtype="int64", ndim=0)


class _Raw:
    """Wraps a python / numpy argument that must reach the op untouched."""
    def __init__(self, v): self.v = v


def _ev(a, ctx):
    if isinstance(a, Var):
        return a._eval(ctx)
    if isinstance(a, _Raw):
        return a.v
    if isinstance(a, (list, tuple)) and any(isinstance(x, Var) for x in a):
        return [_ev(x, ctx) for x in a]
    return a


def _coerce(x):
    return x if isinstance(x, torch.Tensor) else _to_tensor(x)


def _op(fn, *args):
    def f(*vals):
        vals = [_coerce(v) for v in vals]
        if len(vals) == 2 and vals[0].dtype != vals[1].dtype and (vals[0].is_floating_point() or vals[1].is_floating_point()):
            vals = [v.to(_F64) for v in vals]
        return fn(*vals)
    return Var("op", f, args)


# ---------------------------------------------------------------------------------------------
# theano.tensor
# ---------------------------------------------------------------------------------------------
def _input(dtype, ndim):
    def make(name=None):
        return Var("input", name=name, dtype=dtype, ndim=ndim)
    return make


dscalar, dvector, dmatrix = _input("float64", 0), _input("float64", 1), _input("float64", 2)
lscalar, lvector = _input("int64", 0), _input("int64", 1)
bmatrix = _input("int8", 2)


def constant(v, dtype=None, name=None):
    return Var("const", value=v, name=name, ndim=np.ndim(v))


def ones(shape):
    return Var("const", value=np.ones(shape), ndim=len(shape))


def eye(n):
    return Var("const", value=np.eye(n), ndim=2)


def arange(n):
    return _op(lambda k: torch.arange(int(k)), n)


def exp(x): return _op(torch.exp, x)
def log(x): return _op(torch.log, x)
def sqrt(x): return _op(torch.sqrt, x)
def sgn(x): return _op(torch.sign, x)
def abs_(x): return _op(torch.abs, x)
def pow(x, y): return _op(torch.pow, x, y)
def lt(a, b): return _op(lambda x, y: (x < y).to(_F64), a, b)
def gt(a, b): return _op(lambda x, y: (x > y).to(_F64), a, b)
def or_(a, b): return _op(lambda x, y: ((x != 0) | (y != 0)).to(_F64), a, b)
def any(x): return _op(lambda v: (v != 0).any().to(_F64), x)
def clip(x, lo, hi): return _op(lambda v, a, b: torch.clamp(v, float(a), float(b)), x, lo, hi)
def switch(c, a, b): return _op(lambda cc, x, y: torch.where(cc != 0, x, y), c, a, b)
where = switch


def sum(x, axis=None):
    if axis is None:
        return _op(lambda v: v.sum(), x)
    return _op(lambda v: v.sum(dim=axis), x)


def dot(a, b):
    def f(x, y):
        x, y = x.to(_F64), y.to(_F64)
        if x.dim() == 0 or y.dim() == 0:
            return x * y
        return x @ y
    return _op(f, a, b)


def tensordot(a, b, axes=2):
    return _op(lambda x, y: torch.tensordot(x.to(_F64), y.to(_F64), dims=axes), a, b)


def reshape(x, shape, ndim=None):
    shp = _Raw(tuple(int(s) for s in shape)) if not builtins_any(isinstance(s, Var) for s in shape) else shape
    return _op(lambda v, s: v.reshape(tuple(int(k) for k in s)), x, shp)


def transpose(x): return _op(lambda v: v.t() if v.dim() == 2 else v.permute(*reversed(range(v.dim()))), x)
def flatten(x, outdim=1): return _op(lambda v: v.reshape(-1), x)
def addbroadcast(x, *axes): return x
def shape(x): return _op(lambda v: torch.tensor(v.shape), x)
def tile(x, reps, ndim=None): return _op(lambda v, r: v.repeat(*[int(k) for k in r]), x, _Raw(reps))


def shape_padright(x, n_ones=1):
    def f(v):
        for _ in range(n_ones):
            v = v.unsqueeze(-1)
        return v
    out = _op(f, x)
    out.ndim = (x.ndim or 0) + n_ones if isinstance(x, Var) and x.ndim is not None else None
    return out


def shape_padleft(x, n_ones=1):
    def f(v):
        for _ in range(n_ones):
            v = v.unsqueeze(0)
        return v
    return _op(f, x)


def concatenate(lst, axis=0):
    return Var("op", lambda vals: torch.cat([_coerce(v).to(_F64) for v in vals], dim=axis), (list(lst),))


import builtins as _b
builtins_any = _b.any


def grad(cost, wrt, **kw):
    """d cost / d wrt by torch.autograd; `wrt` must be bound (function inputs) at evaluation time."""
    single = not isinstance(wrt, (list, tuple))
    wl = [wrt] if single else list(wrt)
    group = {}

    def run(ctx):
        key = id(group)
        if key in ctx.memo:
            return ctx.memo[key]
        feeds = dict(ctx.feeds)
        leaves = []
        for w in wl:
            if id(w) not in feeds:
                raise KeyError("T.grad: %s is not an input of the compiled function" % w)
            leaf = feeds[id(w)].detach().clone().to(_F64).requires_grad_(True)
            feeds[id(w)] = leaf
            leaves.append(leaf)
        sub = _Ctx(feeds)
        c = cost._eval(sub)
        gs = torch.autograd.grad(c, leaves, allow_unused=True)
        out = [g if g is not None else torch.zeros_like(l) for g, l in zip(gs, leaves)]
        ctx.memo[key] = out
        return out

    outs = [Var("grad", (lambda ctx, i=i: run(ctx)[i]), name="grad(%s)" % w, ndim=w.ndim) for i, w in enumerate(wl)]
    return outs[0] if single else outs


# ---------------------------------------------------------------------------------------------
# theano
# ---------------------------------------------------------------------------------------------
def shared(value=None, name=None, **kw):
    return Var("shared", value=value, name=name, ndim=np.ndim(value),
               dtype=str(np.asarray(value).dtype))


def function(inputs, outputs, givens=(), on_unused_input=None, **kw):
    inputs = list(inputs)

    def call(*vals):
        if len(vals) != len(inputs):
            raise TypeError("expected %d arguments, got %d" % (len(inputs), len(vals)))
        feeds = {id(s): _to_tensor(v) for s, v in zip(inputs, vals)}
        ctx = _Ctx(feeds)
        for s, v in list(givens):
            feeds[id(s)] = v._eval(ctx) if isinstance(v, Var) else _to_tensor(v)
        with torch.enable_grad():
            if isinstance(outputs, (list, tuple)):
                return [_out(o, ctx) for o in outputs]
            return _out(outputs, ctx)
    return call


def _out(o, ctx):
    if not isinstance(o, Var):
        return np.asarray(o)
    return o._eval(ctx).detach().numpy().copy()


def install():
    """Register fake `theano`, `theano.tensor` (and inert `hips` stubs) in sys.modules."""
    me = sys.modules[__name__]
    th = types.ModuleType("theano")
    tt = types.ModuleType("theano.tensor")
    for k, v in vars(me).items():
        if not k.startswith("_"):
            setattr(tt, k, v)
    tt.sum, tt.any, tt.pow, tt.abs_ = sum, any, pow, abs_
    th.tensor = tt
    th.shared = shared
    th.function = function
    th.config = types.SimpleNamespace(floatX="float64")
    sys.modules["theano"] = th
    sys.modules["theano.tensor"] = tt

    # un-vendored sampler library: only the names are needed to import gibbs.py; calling them is an error
    def _missing(*a, **k):
        raise NotImplementedError("hips is not vendored in the reference tree (parity for ARS / HMC draws is unpinned)")
    for modname, names in (("hips", ()), ("hips.inference", ()), ("hips.inference.ars", ("adaptive_rejection_sample",)),
                           ("hips.inference.hmc", ("hmc",))):
        m = types.ModuleType(modname)
        for n in names:
            setattr(m, n, _missing)
        sys.modules[modname] = m
    sys.modules["hips"].inference = sys.modules["hips.inference"]
    sys.modules["hips.inference"].ars = sys.modules["hips.inference.ars"]
    sys.modules["hips.inference"].hmc = sys.modules["hips.inference.hmc"]
