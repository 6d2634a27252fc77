"""TEST INFRASTRUCTURE -- import the reference's own Python (`/root/reference/pyglm`) in this interpreter.

The reference is Python 2 + Theano 0.6 + NumPy of 2014.  Nothing is copied into this repository: an
import hook reads each `pyglm.*` source file where it lies under /root/reference, applies a purely
mechanical Python-2 -> 3 source transform in memory (print statements, tuple parameters, implicit
relative imports, `dict.has_key` ...), and executes it with `theano` replaced by `oracle/theano_shim.py`
and a few removed NumPy aliases restored.  Used only by `oracle/ref_fixtures.py` to mint
`tests/golden/ref_*.npz` (the reference cannot travel to the GPU box; the fixtures do).
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import re
import sys

REF_ROOT = os.environ.get("PYGLM_REFERENCE_ROOT", "/root/reference")


# ---------------------------------------------------------------------------------------------
# mechanical Python 2 -> 3 source transform
# ---------------------------------------------------------------------------------------------
_PRINT = re.compile(r"^(\s*)print(\s+|$)(?!\()(.*)$")
_PRINT_PAREN = re.compile(r"^(\s*)print\s*\((.*)\)\s*$")


def _balanced(s):
    depth = 0
    for ch in s:
        depth += ch in "([{"
        depth -= ch in ")]}"
    return depth <= 0


def _untuple_defs(src):
    """`def f(a, (b, c), d):` -> `def f(a, _tup1, d):` + `b, c = _tup1` as the first body line."""
    out, pos = [], 0
    for m in re.finditer(r"^([ ]*)def\s+\w+\s*\(", src, flags=re.M):
        if m.start() < pos:
            continue
        i, depth = m.end(), 1
        while depth:
            depth += src[i] in "([{"
            depth -= src[i] in ")]}"
            i += 1
        params, close = src[m.end():i - 1], i - 1
        parts, d, cur = [], 0, ""
        for ch in params:
            if ch == "," and d == 0:
                parts.append(cur); cur = ""
                continue
            d += ch in "([{"
            d -= ch in ")]}"
            cur += ch
        parts.append(cur)
        unpack = []
        for k, prm in enumerate(parts):
            if prm.strip().startswith("("):
                name = "_tup%d" % k
                unpack.append("%s = %s" % (prm.strip()[1:-1], name))
                parts[k] = " " + name
        if not unpack:
            continue
        colon = src.index(":", close)
        eol = src.index("\n", colon)
        indent = m.group(1) + "    "
        out.append(src[pos:m.end()] + ",".join(parts) + src[close:eol + 1] +
                   "".join(indent + u + "\n" for u in unpack))
        pos = eol + 1
    out.append(src[pos:])
    return "".join(out)


def py2to3(src, package_modules=()):
    lines = src.replace("\t", "        ").split("\n")
    out = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = _PRINT.match(ln)
        if m and not ln.lstrip().startswith("#"):
            indent, body = m.group(1), m.group(3)
            # continuation lines (backslash or open brackets)
            while body.rstrip().endswith("\\") or not _balanced(body):
                i += 1
                body = body.rstrip().rstrip("\\") + " " + lines[i].strip()
            body = body.rstrip()
            end = ""
            if body.endswith(","):
                body, end = body[:-1], ", end=' '"
            out.append("%sprint(%s%s)" % (indent, body, end))
            i += 1
            continue
        out.append(ln)
        i += 1
    src = "\n".join(out)
    # tuple parameters:  def f(self, vars, (a,b), dt)  /  lambda (k,v): expr
    src = _untuple_defs(src)
    src = re.sub(r"lambda \((\w+),\s*(\w+)\):\s*\(([^)]*)\)", r"lambda _kv: (lambda \1, \2: (\3))(*_kv)", src)
    # NumPy of 2014 evaluated `ndarray == []` to the scalar False; today it is an (ambiguous) empty array
    src = src.replace("elif val == []:", "elif isinstance(val, list) and val == []:")
    src = src.replace(".has_key(", ".__contains__(")
    src = src.replace(".iteritems()", ".items()").replace(".itervalues()", ".values()").replace(".iterkeys()", ".keys()")
    src = re.sub(r"\bxrange\(", "range(", src)
    src = re.sub(r"^(\s*)import cPickle\b", r"\1import pickle as cPickle", src, flags=re.M)
    src = re.sub(r"except (\w+), (\w+):", r"except \1 as \2:", src)
    # implicit relative imports of sibling modules
    for mod in package_modules:
        src = re.sub(r"^(\s*)from %s import" % re.escape(mod), r"\1from .%s import" % mod, src, flags=re.M)
        src = re.sub(r"^(\s*)import %s\s*$" % re.escape(mod), r"\1from . import %s" % mod, src, flags=re.M)
    return src


# ---------------------------------------------------------------------------------------------
# import hook
# ---------------------------------------------------------------------------------------------
class _RefLoader(importlib.abc.Loader):
    def __init__(self, path, is_pkg):
        self.path, self.is_pkg = path, is_pkg

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(self.path, "r") as f:
            src = f.read()
        pkg_dir = os.path.dirname(self.path)
        siblings = [os.path.splitext(n)[0] for n in os.listdir(pkg_dir) if n.endswith(".py") and n != "__init__.py"]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            code = compile(py2to3(src, siblings), self.path, "exec")
        exec(code, module.__dict__)


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != "pyglm" and not fullname.startswith("pyglm."):
            return None
        rel = fullname.split(".")
        base = os.path.join(REF_ROOT, *rel)
        if os.path.isdir(base):
            init = os.path.join(base, "__init__.py")
            spec = importlib.util.spec_from_loader(fullname, _RefLoader(init, True), is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        if os.path.isfile(base + ".py"):
            return importlib.util.spec_from_loader(fullname, _RefLoader(base + ".py", False))
        return None


_installed = False


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "pyglm", "population.py"))


def install():
    """Make `import pyglm...` resolve to the reference tree (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found under %s" % REF_ROOT)
    import numpy as np
    import scipy
    import scipy.special
    from . import theano_shim
    theano_shim.install()
    # NumPy / SciPy names of 2014 that have since been removed (pure aliases)
    for name, val in (("int", int), ("float", float), ("bool", bool), ("Inf", np.inf)):
        if name not in np.__dict__:
            setattr(np, name, val)
    if "scipy.misc" not in sys.modules or not hasattr(sys.modules["scipy.misc"], "logsumexp"):
        import types
        misc = types.ModuleType("scipy.misc")
        misc.logsumexp = scipy.special.logsumexp
        sys.modules["scipy.misc"] = misc
        scipy.misc = misc
    # np.linspace(0, 1, <float>) was accepted (truncated) by the NumPy the reference was written for
    _linspace = np.linspace

    def linspace(start, stop, num=50, *a, **k):
        return _linspace(start, stop, int(num), *a, **k)
    np.linspace = linspace
    sys.meta_path.insert(0, _RefFinder())
    _installed = True
