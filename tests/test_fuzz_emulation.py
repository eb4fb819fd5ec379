"""CPU: randomised shapes through K1 -> K2 (forward, backward from `out` and from the activation bits) -> K4 (max / avg /
sum), all from source on the host (tests/emu), against the dense formulation: sentence widths 1..97 incl. the warp
boundaries, hidden sizes 1..200 incl. unaligned ones, star trees (one row with T-1 neighbours), every slice width."""
import ctypes
import os
import random
import sys

import numpy as np
import pytest
import torch

from gcn_over_pruned_trees_b200 import _lib, ops, synth
from oracle import gcn_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
NAMES = ('gpt_prune_csr', 'gpt_gcn_aggregate_fwd', 'gpt_gcn_aggregate_bwd', 'gpt_pool3_fwd', 'gpt_pool3_bwd')


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
    scale = float(b.double().abs().max())
    return float((a.double() - b.double()).abs().max()) / scale if scale > 0 else float(a.double().abs().max())


def _star_batch(B, T, seed):
    """Token 0 is the root, every other token hangs under it; subject = token 1 (or 0), object = the last token."""
    g = np.random.default_rng(seed)
    head = np.zeros((B, T), dtype=np.int64)
    head[:, 1:] = 1
    deprel = g.integers(2, 42, size=(B, T))
    deprel[:, 0] = 11
    subj = np.full((B, T), 5, dtype=np.int64)
    obj = np.full((B, T), 5, dtype=np.int64)
    subj[:, min(1, T - 1)] = 0
    obj[:, T - 1] = 0
    return [torch.from_numpy(x) for x in (head, subj, obj, deprel)] + [torch.zeros((B, T), dtype=torch.bool)]


@pytest.mark.parametrize('chunk', range(4))
def test_random_shapes_k1_k2_k4_from_source_vs_dense(chunk):
    rnd = random.Random(1000 + chunk)
    for trial in range(8):
        B, k = rnd.choice([1, 2, 5]), rnd.choice([-1, 0, 1, 2])
        H = rnd.choice([1, 2, 3, 4, 5, 8, 31, 32, 33, 36, 100, 200])
        if rnd.random() < 0.4:
            T = rnd.choice([1, 2, 3, 17, 31, 32, 33, 64, 97])
            head, subj, obj, deprel, masks = _star_batch(B, T, trial)
        else:
            batch = synth.make_batch(100 * chunk + trial, batch_size=B, mean_len=rnd.choice([9, 20, 36]))
            head, subj, obj, deprel, masks = batch[5], batch[6], batch[7], batch[4], batch[1]
            T = head.shape[1]
        tag = dict(chunk=chunk, trial=trial, B=B, T=T, H=H, k=k)
        csr = ops.prune_csr(head, subj, obj, deprel, masks, k)
        assert int((csr.err & ops.TREE_ERR_FATAL).sum()) == 0, tag
        g = torch.Generator().manual_seed(trial)
        y = torch.randn(B * T, H, generator=g)
        bias = torch.randn(H, generator=g)
        gout = torch.randn(B, T, H, generator=g)
        mask = ((torch.rand(B * T, H, generator=g) < 0.5).float() * 2.0) if rnd.random() < 0.5 else None
        fv = rnd.choice([0, 0, 1, 2, 4])
        out, act = ops.aggregate_fwd(y, csr, bias, drop_mask=mask, force_vec=fv, want_act=True)
        adj = (csr.to_dense() != 0).double()
        yd = y.view(B, T, H).double().requires_grad_()
        bd = bias.double().requires_grad_()
        ref = torch.relu((adj.bmm(yd) + yd + 2 * bd) / csr.denom.double().unsqueeze(2)) * (csr.flags != 0).unsqueeze(2)
        if mask is not None:
            ref = ref * mask.view(B, T, H).double()
        (ref * gout.double()).sum().backward()
        assert _rel(out, ref.detach()) < 1e-5, tag
        dy, db = ops.aggregate_bwd(gout, out, csr, drop_mask=mask, force_vec=fv)
        dy2, _ = ops.aggregate_bwd(gout, None, csr, drop_mask=mask, force_vec=fv, act=act)
        assert _rel(dy.view(B, T, H), yd.grad) < 1e-5 and _rel(db, bd.grad) < 2e-5 and torch.equal(dy, dy2), tag
        pool_masks = [(csr.flags & bit).eq(0).unsqueeze(2) for bit in (1, 2, 4)]
        for kind in ('max', 'avg', 'sum'):
            got = ops.pool3(out, csr, kind)
            want = torch.cat([gcn_oracle.masked_pool(out, m, kind) for m in pool_masks], dim=1)
            assert torch.equal(torch.isnan(got), torch.isnan(want)), (tag, kind)      # 0/0 of an empty avg pool
            ok = torch.isfinite(want)
            assert not ok.any() or _rel(got[ok], want[ok]) < 1e-5, (tag, kind)
