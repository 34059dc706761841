import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box via gpurun)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """The C-ABI library and the oracle's C restatement are build artefacts (git-ignored): build them if absent
    (nvcc cross-compiles without a GPU), so the CPU suite never depends on a prior manual build step."""
    lib = os.path.join(ROOT, 'rl_games_b200', 'libb200rl.so')
    orc = os.path.join(ROOT, 'oracle', '_build', 'libgae_oracle.so')
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__ as ge
        ge.build()
    yield
