"""Build libb200rl.so (sm_100a only) with nvcc.  In-tree output: rl_games_b200/libb200rl.so
(git-ignored, but shipped to the GPU box by gpurun).  Cross-compiles without a GPU."""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, 'libb200rl.so')
OBJ = os.path.join(HERE, '_obj')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith('.cu'))


def _digest(path):
    h = hashlib.sha1()
    extra = [os.path.join(HERE, 'mlp_tc.cu')] if os.path.basename(path) in ('mlp_tc_relu.cu', 'mlp_tc_tanh.cu') else []     # they #include it
    for dep in [path] + extra + [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith(('.cuh', '.h'))] + \
            [os.path.join(os.path.dirname(PKG), 'include', 'b200rl.h')]:
        with open(dep, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    path = os.path.join(HERE, src)
    obj = os.path.join(OBJ, src[:-3] + '.o')
    stamp = obj + '.sha1'
    dig = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + FLAGS + ['-c', path, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(dig)
    return obj, True


def build(verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(_compile, srcs))
    objs = [o for o, _ in res]
    rebuilt = any(ch for _, ch in res) or not os.path.exists(OUT)
    if rebuilt:
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', OUT] + objs + ['-lcuda']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    if verbose:
        print('libb200rl: %s (%d sources, %s)' % (OUT, len(srcs), 'rebuilt' if rebuilt else 'up to date'))
    return OUT


if __name__ == '__main__':
    build()
