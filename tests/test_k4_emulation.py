"""CPU: the SOURCE of csrc/pool3.cu (K4, fused sentence / subject / object pooling) executed on the host (tests/emu),
on the CSR the emulated K1 produces, through ops.pool3 (forward + autograd backward) against the reference's pool()
(/root/reference/model/gcn.py:473-483, restated in oracle.gcn_oracle.masked_pool).  The `-m gpu` tests of
test_gpu_parity.py run the same cases on the device."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

import cases
from gcn_over_pruned_trees_b200 import _lib, ops, synth
from oracle import gcn_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
NAMES = ('gpt_prune_csr', 'gpt_pool3_fwd', 'gpt_pool3_bwd', 'gpt_pool3_bwd_masked')


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
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('kind', ('max', 'avg', 'sum'))
@pytest.mark.parametrize('H', (200, 37))
def test_k4_source_vs_reference_pool(kind, H):
    batch = synth.make_batch(41, batch_size=16)
    csr = ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], 1)
    B, T = batch[0].shape
    g = torch.Generator().manual_seed(H)
    h = (torch.rand(B, T, H, generator=g) + 0.01).requires_grad_()      # no ties: argmax is unique
    gout = torch.randn(B, 3 * H, generator=g)
    got = ops.pool3(h, csr, kind)
    (got * gout).sum().backward()
    hd = h.detach().clone().requires_grad_()
    m = [csr.pool_mask(), batch[6].ne(0).unsqueeze(2), batch[7].ne(0).unsqueeze(2)]
    ref = torch.cat([gcn_oracle.masked_pool(hd, mi, kind) for mi in m], dim=1)
    (ref * gout).sum().backward()
    assert _rel(got.detach(), ref.detach()) < 1e-6
    assert _rel(h.grad, hd.grad) < 1e-6


def test_k4_source_fully_masked_pool_is_minus_1e12():
    # same-token subject/object -> singleton tree -> empty adjacency -> h_out = -1e12
    head = torch.tensor([[2, 0, 2, 3]])
    pos = torch.from_numpy(cases.positions([3], 4))[None]
    csr = ops.prune_csr(head, pos, pos, torch.tensor([[5, 11, 6, 7]]), torch.zeros((1, 4), dtype=torch.bool), 1)
    h = torch.rand(1, 4, 8)
    out = ops.pool3(h, csr, 'max')
    assert torch.all(out[:, :8] == -1e12)
    assert torch.equal(out[:, 8:16], h[:, 3]) and torch.equal(out[:, 16:], h[:, 3])


def test_k4_source_masked_backward_is_the_plain_backward_times_the_k2_prologue():
    """gpt_pool3_bwd_masked (K4's backward fused with the first step of K2's): dh * dropscale * [out > 0] / denom."""
    batch = synth.make_batch(43, batch_size=8)
    csr = ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], 1)
    B, T = batch[0].shape
    H = 72
    g = torch.Generator().manual_seed(1)
    out = torch.randn(B, T, H, generator=g).clamp_min(0.0)
    _, argmax = ops.pool3_fwd(out, csr, 0)
    gout = torch.randn(B, 3 * H, generator=g)
    act = torch.zeros((B, (H + 31) // 32, T), dtype=torch.int32)
    bits = (out > 0).permute(0, 2, 1).reshape(B, H, T)                 # [B, H, T]
    for c in range(H):
        act[:, c // 32, :] |= (bits[:, c, :].to(torch.int32) << (c % 32))
    plain = ops.pool3_bwd(gout, argmax, csr, 0, H)
    fused = ops.pool3_bwd_masked(gout, argmax, csr, 0, H, act.reshape(-1), 0.5)
    want = plain * (out > 0).float() * ops.drop_scale(0.5) / csr.denom.unsqueeze(2)
    assert _rel(fused, want) < 1e-6
