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
    assert engine_lib.pyglm_b200_abi_version() == 5


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


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under theano_pyglm_b200/ may import it (no CPU fallback in the product)."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "theano_pyglm_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.+oracle)", re.M)
    for dirpath, _dirs, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), os.path.join(dirpath, f)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from theano_pyglm_b200 import engine
    monkeypatch.setattr(engine, "_lib", None)
    monkeypatch.setattr(engine, "LIB_PATH", str(tmp_path / "libpyglm_b200.so"))
    import pytest
    with pytest.raises(engine.EngineError, match="no CPU fallback"):
        engine.load_library()
