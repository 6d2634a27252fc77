"""Nested dict <-> flat vector helpers with the reference's sorted-key ordering.

Interface of pyglm/utils/packvec.py (:3-115).  The ordering is part of the drop-in contract:
optimisers see parameters as `bias` < `bkgd` < `imp` (theano_func_wrapper.py:53-67, packvec.py:23).
"""
from __future__ import annotations

import numpy as np


def pack(var_list):
    shapes = [np.shape(v) for v in var_list]
    flat = [np.ravel(v) for v in var_list]
    return (np.concatenate(flat) if flat else np.zeros((0,))), shapes


def unpack(vec, shapes):
    out, off = [], 0
    for shp in shapes:
        sz = int(np.prod(shp))
        out.append(np.reshape(vec[off:off + sz], shp))
        off += sz
    assert off == len(vec), "Unpack was called with incorrect shapes!"
    return out


def packdict(var_dict, on_unpackable_type='raise'):
    pieces, shapes = [], {}
    for key in sorted(var_dict):
        val = var_dict[key]
        if isinstance(val, dict):
            sub, subshapes = packdict(val, on_unpackable_type)
            pieces.append(sub)
            shapes[key] = subshapes
        elif isinstance(val, list) and len(val) == 0:
            continue
        else:
            if not isinstance(val, np.ndarray):
                if on_unpackable_type.lower() == 'raise':
                    raise Exception("Can only pack numpy arrays!")
                val = np.asarray(val)
            shapes[key] = val.shape
            pieces.append(val.reshape(-1).astype(np.float64, copy=False))
    return (np.concatenate(pieces) if pieces else np.zeros((0,))), shapes


def _unpackdict(vec, shapes, offset):
    out, used = {}, 0
    for key in sorted(shapes):
        shp = shapes[key]
        if isinstance(shp, dict):
            out[key], n = _unpackdict(vec, shp, offset)
        elif isinstance(shp, tuple):
            n = int(np.prod(shp))
            out[key] = np.reshape(vec[offset:offset + n], shp)
        else:
            raise Exception("Can only unpack shape tuples!")
        offset += n
        used += n
    return out, used


def unpackdict(vec, shapes):
    return _unpackdict(vec, shapes, 0)[0]


def get_vars(syms, vars):
    """Sub-dictionary of `vars` with the key structure of `syms`."""
    out = {}
    for k, v in syms.items():
        assert k in vars, "ERROR: syms key %s not found in vars!" % k
        out[k] = get_vars(v, vars[k]) if isinstance(v, dict) else vars[k]
    return out


def set_vars(syms, vars, vals):
    if isinstance(syms, dict):
        for k, v in syms.items():
            assert k in vars, "ERROR: syms key %s not found in vars!" % k
            assert k in vals, "ERROR: syms key %s not found in vals!" % k
            vars[k] = set_vars(v, vars[k], vals[k]) if isinstance(v, dict) else vals[k]
    elif syms in vars:
        vars[syms] = vals
    else:
        raise Exception("Can only set variables for a dictionary of symbolic vars or a specific key in vars")
    return vars


def get_shapes(x, syms):
    return packdict(get_vars(syms, x))[1]
