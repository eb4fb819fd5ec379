import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, 'tests', 'golden')
for p in (REPO, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        # fp32 parity: keep cuBLAS / cuDNN (the C-GCN LSTM) out of TF32
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_adj():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'adjacency.npz'))


@pytest.fixture(scope='session')
def golden_model():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'model.npz'))


@pytest.fixture(scope='session')
def golden_deprel():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'deprel.npz'))
