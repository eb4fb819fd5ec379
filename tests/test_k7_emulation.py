"""CPU: the SOURCE of csrc/update.cu (K7: global-norm clip + SGD + zero_grad over one flat parameter buffer and the live
word-embedding rows, two launches) executed on the host (tests/emu) through ops.update_sqnorm / update_apply against
torch.nn.utils.clip_grad_norm_ + torch.optim.SGD (/root/reference/train.py:224-227)."""
import ctypes
import os
import sys

import pytest
import torch

from gcn_over_pruned_trees_b200 import _lib, ops

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
NAMES = ('gpt_update_partials', 'gpt_update_sqnorm', 'gpt_update_apply')


@pytest.fixture(scope='module', autouse=True)
def emulated():
    import emu_build
    handle = ctypes.CDLL(emu_build.build())
    for name in NAMES:
        getattr(handle, name).argtypes = _lib.SIGNATURES[name]
        getattr(handle, name).restype = ctypes.c_int
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, '_lib', handle)
    mp.setattr(ops, '_stream', lambda: None)
    yield handle
    mp.undo()


@pytest.mark.parametrize('n,max_norm', [(281_242, 5.0), (1027, 0.1), (6, 5.0)])
@pytest.mark.parametrize('with_rows', (False, True))
def test_k7_source_clip_sgd_zero_grad_equals_torch(n, max_norm, with_rows):
    g = torch.Generator().manual_seed(n)
    param = torch.randn(n, generator=g)
    grad = torch.randn(n, generator=g) * 0.05
    V, E, lr = 90, 300, 0.3
    sparse, emb, dense_emb_grad = None, None, None
    if with_rows:
        emb = torch.randn(V, E, generator=g)
        sparse = ops.SparseEmbeddingState(emb, V - 10)                  # topn: the last 10 rows are frozen
        words = torch.randint(0, V, (4, 37), generator=g)
        sparse.words = words
        dense_emb_grad = torch.zeros(V, E)
        for row, w in enumerate(words.flatten().tolist()):               # what K5's backward leaves behind
            if w == 0 or w >= V - 10:
                continue
            dense_emb_grad[w] += torch.randn(E, generator=g) * 0.05
            sparse.owner[w] = min(int(sparse.owner[w]), row)
        sparse.G.copy_(dense_emb_grad)
    # torch side
    p_ref = param.clone().requires_grad_()
    p_ref.grad = grad.clone()
    tensors = [p_ref]
    if with_rows:
        e_ref = emb.clone().requires_grad_()
        e_ref.grad = dense_emb_grad.clone()
        tensors.append(e_ref)
    total = torch.nn.utils.clip_grad_norm_(tensors, max_norm)
    torch.optim.SGD(tensors, lr=lr).step()
    # K7
    partials = torch.zeros(ops.update_partials(n, sparse.words.numel() if sparse else 0))
    total_norm = torch.zeros(())
    counter = torch.tensor([7, 41], dtype=torch.int64)
    ops.update_sqnorm(grad, sparse, partials)
    ops.update_apply(param, grad, sparse, emb, partials, max_norm, lr, total_norm=total_norm, step_counter=counter[1:])
    assert abs(float(total_norm) - float(total)) <= 1e-5 * float(total)
    assert float((param - p_ref.detach()).abs().max()) <= 1e-6 * float(p_ref.detach().abs().max())
    assert float(grad.abs().max()) == 0.0                                # zero_grad is part of the update
    assert counter.tolist() == [7, 42]                                   # the dropout step word advanced
    if with_rows:
        assert float((emb - e_ref.detach()).abs().max()) <= 1e-6 * float(e_ref.detach().abs().max())
        assert float(sparse.G.abs().max()) == 0.0 and int(sparse.owner.min()) == 0x7fffffff
