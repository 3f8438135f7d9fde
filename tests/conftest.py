import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = ['push_nominal', 'push_M10', 'push_recovery', 'push_badstate', 'push_null_action', 'peg_nominal',
                'push_mlp_notransform', 'pendulum']
# online GP residual: M genuinely different replicas, injected GP draws
GP_CASES = ['push_gp_latent', 'push_gp_local']


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    import numpy as np
    from tampc_b200 import spec as S
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    return z, S.spec_from_arrays(z)


@pytest.fixture(scope='session')
def cuda_lib():
    """The CUDA library must be the thing that runs: build it if the in-tree .so is missing, never fall back."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from tampc_b200 import _lib, build
    if not os.path.exists(_lib.lib_path()):
        build.build()
    return _lib.load()
