"""Import the upstream reference read-only from /root/reference (authoring container only).

The GPU box has no /root/reference: nothing under tests/ -m gpu, smoke() or bench.py may
call this.  It exists for ``oracle/make_golden.py`` and ``oracle/check_against_reference.py``.

``skimage`` is not installed (and not installable offline); ``utils/matcher.py:4`` imports
``skimage.feature.match_descriptors``, so a module shim exposing the oracle's restatement of
that one function is registered before import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('KB_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'utils'))


def load():
    """Returns a namespace with the reference modules on the hot path."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    sys.dont_write_bytecode = True
    from oracle.ref_ops import match_descriptors
    if 'skimage.feature' not in sys.modules:
        sk = types.ModuleType('skimage')
        skf = types.ModuleType('skimage.feature')
        skf.match_descriptors = match_descriptors
        sk.feature = skf
        sys.modules['skimage'] = sk
        sys.modules['skimage.feature'] = skf
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.extracter = importlib.import_module('utils.extracter')
    ns.matcher = importlib.import_module('utils.matcher')
    ns.projection = importlib.import_module('utils.projection')
    ns.repeatability = importlib.import_module('tasks.repeatability')
    ns.mha = importlib.import_module('tasks.MHA')
    return ns


def load_lightglue_extract():
    """The LightGlue-style extract helpers of ``models/lightglue.py`` live inside its
    ``if __name__ == '__main__'`` demo block (:904-979) and the module itself needs packages that
    are not installed, so the function definitions are taken from the parsed source (never
    copied into this repo) and executed in a namespace that holds only torch."""
    import ast
    import torch
    path = os.path.join(REFERENCE_ROOT, 'models', 'lightglue.py')
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    want = {'sample_descriptors', 'simple_nms', 'top_k_kps', 'extract'}
    found = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {n.name for n in found} == want, [n.name for n in found]
    mod = ast.Module(body=found, type_ignores=[])
    ns = {'torch': torch}
    exec(compile(mod, path, 'exec'), ns)
    return types.SimpleNamespace(**{k: ns[k] for k in want})
