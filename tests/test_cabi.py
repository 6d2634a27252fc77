"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports every symbol
include/pyglm_b200.h declares.  No compute calls (no GPU here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pyglm_b200.h")).read()
    return sorted(set(re.findall(r"PYGLM_B200_API[^;(]*?\b(pyglm_b200_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for name in ("pyglm_b200_dataset_create", "pyglm_b200_ll_grad", "pyglm_b200_ll_grad_dev",
                 "pyglm_b200_gibbs_delta_ll", "pyglm_b200_gibbs_commit"):
        assert name in syms


def test_library_exports_every_declared_symbol(engine_lib):
    from theano_pyglm_b200 import engine
    for name in declared_symbols():
        assert hasattr(engine_lib, name), name
    assert sorted(engine.EXPORTS) == declared_symbols()
    assert engine_lib.pyglm_b200_abi_version() == 2


def test_header_cites_reference_lines():
    src = open(os.path.join(ROOT, "include", "pyglm_b200.h")).read()
    for cite in ("utils/basis.py:201-236", "glm.py:52", "gibbs.py:812-864", "theano_func_wrapper.py:12-51"):
        assert cite in src


def test_sass_is_sm100a_only():
    """The shipped library carries sm_100a code only (no multi-arch fatbin)."""
    import subprocess
    from theano_pyglm_b200 import engine
    out = subprocess.run(["cuobjdump", "-lelf", engine.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
