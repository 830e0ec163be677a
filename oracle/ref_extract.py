"""TEST INFRASTRUCTURE ONLY -- loads the reference's own pure-NumPy greedy functions, unmodified.

`/root/reference/placement_algorithm2.py` cannot be imported (its module top imports TensorFlow,
TFP and matplotlib, `placement_algorithm2.py:1-22`), so the function definitions that only need NumPy
are cut out of the file with `ast` and exec'd verbatim with `np` in scope:

    argmax_cache_linear   placement_algorithm2.py:53-67
    argmax_               placement_algorithm2.py:105-125
    placement_algorithm_1 placement_algorithm2.py:128-145
    placement_algorithm_2 placement_algorithm2.py:151-219
    nominator             placement_algorithm2.py:371-388
    make_slice            placement_algorithm2.py:391-396
    call_pinv             placement_algorithm2.py:399-405
    denominator           placement_algorithm2.py:408-413
    dg_create_random_cov  placement_algorithm2.py:441-444
    cov_vv_4x4            placement_algorithm2.py:473-479

Nothing is copied into this repository: the source is read where it lies at call time.  The reference
tree exists only in the build container, never on the GPU box, so this module is used by
`tests/golden/make_golden.py` (fixture generation) and by CPU tests that skip when the tree is absent.
No product code imports it.
"""
import ast
import contextlib
import io
import os

import numpy as np

REFERENCE_ROOT = os.environ.get("VGPOSP_REFERENCE_ROOT", "/root/reference")
_WANTED = (
    "argmax_cache_linear", "argmax_", "placement_algorithm_1", "placement_algorithm_2",
    "nominator", "make_slice", "call_pinv", "denominator", "dg_create_random_cov", "cov_vv_4x4",
)


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "placement_algorithm2.py"))


def load():
    """Return a dict name -> function holding the reference's NumPy greedy, exec'd unmodified."""
    path = os.path.join(REFERENCE_ROOT, "placement_algorithm2.py")
    with open(path, "r") as fh:
        src = fh.read()
    tree = ast.parse(src, filename=path)
    keep = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name in _WANTED]
    missing = set(_WANTED) - {n.name for n in keep}
    if missing:
        raise RuntimeError("reference is missing %s" % sorted(missing))
    module = ast.Module(body=keep, type_ignores=[])
    namespace = {"np": np}
    exec(compile(module, path, "exec"), namespace)
    return {name: namespace[name] for name in _WANTED}


def run_quiet(fn, *args):
    """Call a reference function, returning (result, captured stdout) -- alg. 2 prints per evaluation
    (`placement_algorithm2.py:188,205`)."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*args)
    return out, buf.getvalue()
