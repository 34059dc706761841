"""The test-only build of the library (tests/libb200rl_testhooks.so): the product objects plus the `b200rl_hosttest_*` host entry points
(`#ifdef B200RL_TEST_HOOKS` in csrc/*.cu: per-thread kernel bodies and row functions run over HOST arrays).  Built by
`rl_games_b200/csrc/build.py::build_test_hooks()` (called from `__graft_entry__.build()`); the product library does not export these
symbols and nothing under rl_games_b200/ loads this file."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOOKS_LIB_PATH = os.path.join(ROOT, 'tests', 'libb200rl_testhooks.so')
_cdll = None


def load():
    global _cdll
    if _cdll is None:
        if not os.path.exists(HOOKS_LIB_PATH):
            from rl_games_b200.csrc import build as _b
            _b.build_test_hooks(verbose=False)
        _cdll = ctypes.CDLL(HOOKS_LIB_PATH)
    return _cdll
