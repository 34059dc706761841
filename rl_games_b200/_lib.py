"""ctypes binding of libb200rl.so -- the C-ABI boundary (include/b200rl.h).

The prototypes are parsed from the header itself so the binding cannot drift from the declared ABI.
There is NO CPU fallback: if the library is missing or was not built for sm_100a, importing the ops
raises, loudly (the product path never routes through oracle/).
"""
import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
HEADER = os.path.join(_ROOT, 'include', 'b200rl.h')
# B200RL_LIB_PATH: developer override for A/B-timing two builds of the same ABI inside one GPU session
LIB_PATH = os.environ.get('B200RL_LIB_PATH') or os.path.join(_PKG, 'libb200rl.so')

_SCALARS = {
    'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64, 'uint32_t': ctypes.c_uint32,
    'float': ctypes.c_float, 'double': ctypes.c_double,
}


def parse_header(path=HEADER):
    """Return {name: [(argname, ctype), ...]} for every `int b200rl_*(...)` declaration."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'//[^\n]*', ' ', src)
    protos = {}
    for m in re.finditer(r'\b(?:int|int64_t)\s+(b200rl_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        sig = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a:
                    sig.append((a.split('*')[-1].strip(), ctypes.c_void_p))
                else:
                    toks = [t for t in a.split(' ') if t not in ('const', 'unsigned')]
                    sig.append((toks[-1], _SCALARS[toks[0]]))
        protos[name] = sig
    return protos


RET_I64 = ('b200rl_tc_pack_bytes', 'b200rl_tc_xtile_bytes')


class _Lib:
    def __init__(self):
        self._cdll = None
        self._prof = None
        self.protos = parse_header()

    def load(self):
        if self._cdll is not None:
            return self._cdll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m rl_games_b200.csrc.build` '
                '(or __graft_entry__.build()).  rl_games_b200 has no CPU fallback.')
        cdll = ctypes.CDLL(LIB_PATH)
        for name, sig in self.protos.items():
            fn = getattr(cdll, name)   # AttributeError => header/library drift, fail loudly
            fn.restype = ctypes.c_int64 if name in RET_I64 else ctypes.c_int
            fn.argtypes = [t for _, t in sig]
        if cdll.b200rl_built_arch() != 100:
            raise RuntimeError('libb200rl.so was not built for sm_100a')
        self._cdll = cdll
        return cdll

    def __getattr__(self, name):
        if name.startswith('b200rl_'):
            fn = getattr(self.load(), name)
            if self.__dict__.get('_prof') is None:
                return fn
            return self._profiled(name, fn)
        raise AttributeError(name)

    # ---- optional per-call CUDA-event timing (bench.py kernel breakdown; never on in the product path) ----
    def start_profile(self):
        self.__dict__['_prof'] = []

    def _profiled(self, name, fn):
        import torch

        def wrapped(*args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            self.__dict__['_prof'].append((name, s, e))
            return rc
        return wrapped

    def stop_profile(self):
        import torch
        torch.cuda.synchronize()
        recs, out = self.__dict__.get('_prof') or [], {}
        self.__dict__['_prof'] = None
        for name, s, e in recs:
            d = out.setdefault(name.replace('b200rl_', ''), {'n': 0, 'ms': 0.0})
            d['n'] += 2 if name == 'b200rl_prepare_batch_f32' else 1
            d['ms'] += s.elapsed_time(e)
        return out


lib = _Lib()


class B200RLError(RuntimeError):
    pass


def check(rc, what=''):
    if rc != 0:
        if rc > 0:
            raise B200RLError(f'{what}: CUDA error {rc}')
        raise B200RLError(f'{what}: argument error {rc}')


def ptr(t):
    """device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
