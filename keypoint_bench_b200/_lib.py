"""ctypes binding of the C-ABI library ``csrc/libkb_b200.so`` (declared in include/kb_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or cannot be loaded, importing
this module raises, and every op raises when CUDA is unavailable.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# KB_LIB selects another build of the same library (the debug build with bounds asserts: `make -C csrc debug`)
LIB_PATH = os.path.abspath(os.environ['KB_LIB']) if os.environ.get('KB_LIB') else os.path.join(_HERE, 'csrc', 'libkb_b200.so')

KB_OK = 0
KB_ERR_BAD_ARG = -1
KB_ERR_WORKSPACE = -2
KB_ERR_UNSUPPORTED = -3
KB_KNOB_TC_CLUSTER, KB_KNOB_TC_DEBUG, KB_KNOB_REP_NO_SORT, KB_KNOB_TC_ONE_PASS, KB_KNOB_SPARSE_PROF, KB_KNOB_TC_BF16X3 = 1, 2, 3, 4, 5, 6
KB_KNOB_SAMPLE_4CH = 7

# name -> (restype, argtypes); mirrors include/kb_b200.h one to one
PROTOTYPES = {
    'kb_version': (c_int, []),
    'kb_error_string': (ctypes.c_char_p, [c_int]),
    'kb_debug_knob': (c_int, [c_int, c_int]),
    'kb_debug_sparse_prof': (c_int, [c_void_p]),
    'kb_fast_nms_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'kb_fast_nms': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                            c_size_t, c_void_p]),
    'kb_select_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'kb_select': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'kb_simple_nms_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'kb_simple_nms': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    'kb_lk_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'kb_lk_track': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                            c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'kb_detect_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_float]),
    'kb_detect': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'kb_detect_phases': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    'kb_sample_desc': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                               c_int, c_void_p, c_void_p]),
    'kb_sample_desc_operands_supported': (c_int, [c_int, c_int, c_int, c_int]),
    'kb_sample_desc_operands': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_size_t, c_void_p]),
    'kb_match_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'kb_match_mnn': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_int,
                             c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'kb_match_mnn_phases': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_int,
                                    c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    'kb_match_tc_debug_offsets': (c_int, [c_int, c_int, c_int, c_int, c_void_p]),
    'kb_warp_homography': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    'kb_warp_se3': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p]),
    'kb_repeat_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'kb_repeat_counts': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                 c_void_p]),
    'kb_corner_error': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                c_void_p, c_void_p, c_void_p]),
}


class KbError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or `make -C keypoint_bench_b200/csrc`). There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str = '') -> None:
    if rc != KB_OK:
        msg = lib.kb_error_string(int(rc)).decode()
        raise KbError(f'{what or "kb_b200"} failed with code {rc}: {msg}')


def version() -> int:
    return int(lib.kb_version())
