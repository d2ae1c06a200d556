"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/kb_b200.h declares,
and its pure-host entry points (version, error strings, workspace sizing, argument validation) behave.
No kernels are launched here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'kb_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r'KB_API\s+[\w\s\*]+?\b(kb_\w+)\s*\(', text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ('kb_fast_nms', 'kb_select', 'kb_detect', 'kb_sample_desc', 'kb_match_mnn', 'kb_warp_homography',
                 'kb_repeat_counts', 'kb_corner_error', 'kb_version', 'kb_error_string'):
        assert must in names, must


def test_library_exports_every_declared_symbol_and_binding_matches():
    from keypoint_bench_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f'{name} declared in include/kb_b200.h but not exported'
        assert name in _lib.PROTOTYPES, f'{name} has no ctypes prototype in _lib.py'
    for name in _lib.PROTOTYPES:
        assert name in declared_symbols(), f'{name} bound in _lib.py but not declared in the header'


def test_host_only_entry_points():
    from keypoint_bench_b200 import _lib
    lib = _lib.lib
    assert _lib.version() >= 100
    assert lib.kb_error_string(0) == b'ok'
    assert b'workspace' in lib.kb_error_string(_lib.KB_ERR_WORKSPACE)
    assert b'argument' in lib.kb_error_string(_lib.KB_ERR_BAD_ARG)
    # workspace sizing is monotone in the batch and zero for empty problems
    assert lib.kb_fast_nms_workspace_bytes(0, 480, 640) == 0
    a = lib.kb_detect_workspace_bytes(1, 480, 640, 6, 1000, 0.0)
    b = lib.kb_detect_workspace_bytes(8, 480, 640, 6, 1000, 0.0)
    assert 0 < a < b
    assert lib.kb_match_workspace_bytes(1, 1000, 1000, 256, 1) > 0
    assert lib.kb_match_workspace_bytes(1, 1000, 1000, 256, -1) >= lib.kb_match_workspace_bytes(1, 1000, 1000, 256, 0)
    assert lib.kb_repeat_workspace_bytes(2, 1000, 1000) > 0
    # argument validation happens before any CUDA call
    assert lib.kb_detect(None, 1, 480, 640, 6, 8, 0.0, 0.0, 1000, None, None, None, None, None, 0, None) == _lib.KB_ERR_BAD_ARG
    assert lib.kb_match_mnn(None, None, None, None, 1, 10, 10, 8, 1.0, 1, 0, None, None, None, None, 0, None) == _lib.KB_ERR_BAD_ARG
    assert lib.kb_sample_desc(None, 1, 8, 4, 4, None, 2, None, 4, 0, 0, 8, None, None) == _lib.KB_ERR_BAD_ARG


def test_ops_refuse_cpu_tensors():
    import torch
    from keypoint_bench_b200 import _lib, ops
    with pytest.raises(_lib.KbError):
        ops.detect_batched(torch.rand(1, 1, 32, 32), dict(nms_dist=4, threshold=0.0, border_dist=4, top_k=10, min_score=0.0))
    with pytest.raises(_lib.KbError):
        ops.match_batched(torch.rand(1, 4, 8), torch.rand(1, 4, 8))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'keypoint_bench_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('the oracle', '').replace("oracle's", ''), os.path.join(dirpath, f)
