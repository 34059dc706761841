"""Helper for tests/test_bench_dryrun_cpu.py (run as a script in its own process): executes bench.py's B200 arm on a machine without a
GPU with every CUDA touch point replaced -- the C-ABI library by the prototype-checking recorder of test_abi_calls_cpu.py (no compute),
torch.cuda by no-ops, 'cuda:N' devices by the CPU -- so that the whole Python path of the measurement (agent construction through the
registries, warm-up, timed loop, kernel breakdown, roofline / e2e assembly, the one JSON line) is executed end to end."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import test_agent_host_cpu as H
    from test_abi_calls_cpu import _Recorder
    from rl_games_b200 import ops, _lib
    rec = _Recorder(_lib.lib)
    ops.lib = rec
    real_device = torch.device

    class _Meta(type):
        def __instancecheck__(cls, obj):
            return isinstance(obj, real_device)

    class _Dev(metaclass=_Meta):
        def __new__(cls, *a, **k):
            if a and isinstance(a[0], str) and a[0].startswith('cuda'):
                return real_device('cpu')
            return real_device(*a, **k)
    torch.device = _Dev

    def on_cpu(fn):
        def wrapped(*a, **k):
            if isinstance(k.get('device'), str) and k['device'].startswith('cuda'):
                k['device'] = 'cpu'
            return fn(*a, **k)
        return wrapped
    for name in ('empty', 'zeros', 'ones', 'full', 'tensor', 'arange', 'randn', 'rand'):
        setattr(torch, name, on_cpu(getattr(torch, name)))
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.current_device = lambda: 0
    torch.cuda.Event = H._Event
    torch.cuda.current_stream = lambda *a: H._Stream()
    torch.Tensor.pin_memory = lambda self: self
    import bench
    sys.argv = ['bench.py'] + sys.argv[1:]
    bench.main()
    print('RECORDED ' + json.dumps(rec.calls), file=sys.stderr)


if __name__ == '__main__':
    main()
