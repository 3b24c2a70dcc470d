"""Drop-in surface check against the reference's own run scripts (python/run_*.py), read from /root/reference where it is
mounted and never copied: every `class_files` import they make resolves in this package, every constructor call uses
keywords this package's constructors accept, and every attribute or method they touch on a System / iLQR object exists
here.  (The scripts cannot be EXECUTED against the CUDA backend anywhere: they live on the machine without a GPU, and the
GPU box has no /root/reference -- DESIGN.md section 9.  examples/ re-types their loops for the GPU tests.)"""
import ast
import glob
import importlib
import inspect
import os

import pytest

REF = "/root/reference/python"
SCRIPTS = sorted(glob.glob(os.path.join(REF, "run_*.py")))
pytestmark = pytest.mark.skipif(not SCRIPTS, reason="the reference sources are not mounted here")
OUT_OF_SCOPE = ("class_files.animations",)          # VTK / matplotlib animation helpers (out of scope)


def _instance_attrs(cls):
    """names set on self anywhere in the class hierarchy's sources + class attributes / properties / methods"""
    names = set(dir(cls))
    for k in cls.__mro__:
        try:
            tree = ast.parse(inspect.getsource(k).lstrip() if not inspect.getsource(k).startswith("class") else inspect.getsource(k))
        except (OSError, TypeError, IndentationError, SyntaxError):
            continue
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "self" \
                    and isinstance(node.ctx, ast.Store):
                names.add(node.attr)
    return names


@pytest.mark.parametrize("script", SCRIPTS, ids=[os.path.basename(s) for s in SCRIPTS])
def test_run_script_only_uses_what_this_package_provides(script):
    tree = ast.parse(open(script).read())
    classes = {}                                      # local name -> class object of THIS package
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("class_files"):
            if node.module.startswith(OUT_OF_SCOPE):
                continue
            mod = importlib.import_module(node.module)            # resolves to iterative-linear-quadratic-regulator_b200/
            assert "iterative-linear-quadratic-regulator_b200" in mod.__file__
            for alias in node.names:
                assert hasattr(mod, alias.name), (node.module, alias.name)
                classes[alias.asname or alias.name] = getattr(mod, alias.name)
    assert classes, "the script imports nothing from class_files?"
    # constructor calls: keywords accepted, positional count within the signature; remember what each variable holds
    holds = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Call) and isinstance(node.value.func, ast.Name) \
                and node.value.func.id in classes:
            cls = classes[node.value.func.id]
            sig = inspect.signature(cls.__init__)
            params = list(sig.parameters)[1:]
            has_kw = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values())
            for kw in node.value.keywords:
                assert kw.arg is None or kw.arg in params or has_kw, (cls.__name__, kw.arg)
            assert len(node.value.args) <= len(params), (cls.__name__, len(node.value.args))
            for tgt in node.targets:
                if isinstance(tgt, ast.Name):
                    holds[tgt.id] = cls
    assert holds, "no System / iLQR object is constructed?"
    # every attribute read, written or called on those objects
    touched = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in holds:
            cls = holds[node.value.id]
            assert node.attr in _instance_attrs(cls), f"{os.path.basename(script)}: {cls.__name__}.{node.attr} is missing here"
            touched += 1
    assert touched > 0
    # results are used as arrays with .block_until_ready(): the returned host arrays provide it
    from class_files._device import HostArray
    assert hasattr(HostArray, "block_until_ready")
