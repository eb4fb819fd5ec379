"""CPU: the SOURCE of csrc/embed.cu (K5: embedding gather + concat + input dropout, scatter backward, row-sparse
clip/SGD tail) executed on the host (tests/emu) through ops.embed_concat against torch's embedding + autograd.  The
`-m gpu` tests of test_gpu_parity.py run the same checks on the device."""
import ctypes
import os
import sys

import pytest
import torch

from gcn_over_pruned_trees_b200 import _lib, ops, synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
NAMES = ('gpt_embed_fwd', 'gpt_embed_bwd', 'gpt_embed_rows_sqnorm', 'gpt_embed_rows_sgd')


@pytest.fixture(scope='module', autouse=True)
def emulated():
    import emu_build
    handle = ctypes.CDLL(emu_build.build())
    for name in NAMES:
        getattr(handle, name).argtypes = _lib.SIGNATURES[name]
        getattr(handle, name).restype = ctypes.c_int
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, '_lib', handle)
    mp.setattr(ops, '_dev', lambda t, dtype, name: t.contiguous() if t.dtype == dtype else (_ for _ in ()).throw(
        TypeError('%s must be %s' % (name, dtype))))
    mp.setattr(ops, '_stream', lambda: None)
    yield handle
    mp.undo()


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _torch_embed(words, pos, ner, tabs):
    parts = [torch.nn.functional.embedding(words, tabs[0], padding_idx=0), torch.nn.functional.embedding(pos, tabs[1])]
    if tabs[2] is not None:
        parts.append(torch.nn.functional.embedding(ner, tabs[2]))
    return torch.cat(parts, 2)


@pytest.mark.parametrize('dataset', ('tacred', 'semeval'))
def test_k5_source_embed_concat_vs_torch(dataset):
    V, E, Dp, Dn = 300, 300, 30, 30 if dataset == 'tacred' else 0
    batch = synth.make_batch(61, batch_size=12, vocab_size=V, dataset=dataset)
    words, pos = batch[0], batch[2]
    ner = batch[3] if dataset == 'tacred' else None
    g = torch.Generator().manual_seed(1)
    tabs = [torch.randn(V, E, generator=g).requires_grad_(), torch.randn(47, Dp, generator=g).requires_grad_(),
            torch.randn(15, 30, generator=g).requires_grad_() if Dn else None]
    x = ops.embed_concat(words, pos, ner, tabs[0], tabs[1], tabs[2])
    r = torch.randn(x.shape, generator=g)
    (x * r).sum().backward()
    ref_t = [t.detach().clone().requires_grad_() if t is not None else None for t in tabs]
    xr = _torch_embed(words, pos, ner, ref_t)
    (xr * r).sum().backward()
    assert torch.equal(x, xr)
    for a, b in zip(tabs, ref_t):
        if a is not None:
            assert _rel(a.grad, b.grad) < 1e-5


def test_k5_source_dropout_mask_is_replayed_in_backward_and_topn_freezes_rows():
    V, E = 200, 64
    batch = synth.make_batch(62, batch_size=20, vocab_size=V)
    words, pos, ner = batch[0], batch[2], batch[3]
    g = torch.Generator().manual_seed(2)
    emb = (torch.rand(V, E, generator=g) + 0.5).requires_grad_()
    pw = (torch.rand(47, 8, generator=g) + 0.5).requires_grad_()
    nw = (torch.rand(15, 8, generator=g) + 0.5).requires_grad_()
    rng = torch.tensor([99, 5], dtype=torch.int64)
    x = ops.embed_concat(words, pos, ner, emb, pw, nw, drop_p=0.5, rng_state=rng, subseq=3, topn=150)
    keep = (x != 0)
    assert abs(keep[words != 0].float().mean().item() - 0.5) < 0.03
    r = torch.randn(x.shape, generator=g)
    (x * r).sum().backward()
    want = torch.zeros(V, E)
    want.index_put_((words.flatten(),), (r * keep * 2.0)[..., :E].reshape(-1, E), accumulate=True)
    want[0] = 0
    want[150:] = 0                                          # frozen rows (topn)
    assert _rel(emb.grad, want) < 1e-5
    x2 = ops.embed_concat(words, pos, ner, emb, pw, nw, drop_p=0.5, rng_state=rng + torch.tensor([0, 1]), subseq=3)
    assert not torch.equal(x2 != 0, keep)


def test_k5_source_row_sparse_clip_and_sgd_tail_equals_the_dense_update():
    """SparseEmbeddingState path (engine.GraphedTrainStep): gradients land in the all-zero-between-steps buffer G, the
    norm of the live rows is added to the clip, live rows are updated and re-zeroed -- same result as dense
    clip_grad_norm_ + SGD on the full [V, E] gradient."""
    V, E = 120, 40
    batch = synth.make_batch(63, batch_size=10, vocab_size=V)
    words, pos, ner = batch[0], batch[2], batch[3]
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(V, E, generator=g).requires_grad_()
    pw = torch.randn(47, 8, generator=g).requires_grad_()
    nw = torch.randn(15, 8, generator=g).requires_grad_()
    state = ops.SparseEmbeddingState(emb.data, V)
    x = ops.embed_concat(words, pos, ner, emb, pw, nw, sparse=state)
    r = torch.randn(x.shape, generator=g)
    (x * r).sum().backward()
    assert emb.grad is None                                  # the gradient went into state.G
    dense = torch.zeros(V, E)
    dense.index_put_((words.flatten(),), r[..., :E].reshape(-1, E), accumulate=True)
    dense[0] = 0
    assert _rel(state.G, dense) < 1e-5
    state.sq.zero_()
    ops.embed_rows_sqnorm(state)
    assert abs(float(state.sq) - float((dense.double() ** 2).sum())) <= 1e-4 * float((dense.double() ** 2).sum())
    max_norm, lr = 0.5, 0.3
    before = emb.data.clone()
    ops.embed_rows_sgd(state, emb.data, state.sq, max_norm, lr)
    coef = min(1.0, max_norm / (float(state.sq) ** 0.5 + 1e-6))
    assert _rel(emb.data, before - lr * coef * dense) < 1e-5
    assert float(state.G.abs().max()) == 0.0 and int(state.owner.min()) == 0x7fffffff     # re-armed for the next step
