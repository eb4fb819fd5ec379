"""CPU: randomised shapes through K1 -> K2 (forward, backward from `out` and from the activation bits) -> K4 (max / avg /
sum), all from source on the host (tests/emu), against the dense formulation: sentence widths 1..97 incl. the warp
boundaries, hidden sizes 1..200 incl. unaligned ones, star trees (one row with T-1 neighbours), every slice width; and
randomised configurations of the relation-aware model (K1, K5, K3, K10, K4 from source) against the oracle."""
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
NAMES = ('gpt_prune_csr', 'gpt_gcn_aggregate_fwd', 'gpt_gcn_aggregate_bwd', 'gpt_pool3_fwd', 'gpt_pool3_bwd',
         'gpt_embed_fwd', 'gpt_embed_bwd', 'gpt_linear_fwd_f32', 'gpt_linear_dgrad_f32', 'gpt_linear_wgrad_f32',
         'gpt_relmix_fwd', 'gpt_relmix_bwd', 'gpt_diagmix_fwd', 'gpt_diagmix_bwd', 'gpt_agg3_fwd', 'gpt_agg3_bwd',
         'gpt_colsum_acc', 'gpt_live_rows', 'gpt_gather_rows', 'gpt_scatter_rows', 'gpt_relmix_fwd_rows', 'gpt_relmix_bwd_rows',
       'gpt_colsum_acc_rows', 'gpt_linear_wgrad_rows_f32')


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


@pytest.mark.parametrize('chunk', range(2))
def test_random_relation_aware_configurations_from_source_vs_oracle(chunk):
    """hidden 1..33, 1..7 relation slots, 1..3 layers, every switch of the mode, all three poolings: loss and every
    gradient of one train-mode step (dropout probabilities 0, nothing injected or replaced)."""
    import weights
    from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
    rnd = random.Random(2000 + chunk)
    for trial in range(6):
        H, D, L = rnd.choice([1, 2, 3, 4, 5, 8, 12, 33]), rnd.choice([1, 2, 3, 7]), rnd.choice([1, 2, 3])
        mode = rnd.choice(['full_deprel', 'full_deprel', 'diagonal_deprel'])
        over = dict(adj_type=mode, deprel_emb_dim=D, hidden_dim=H, num_layers=L, prune_k=rnd.choice([-1, 0, 1, 2]),
                    vocab_size=60, input_dropout=0.0, gcn_dropout=0.0, gemm_mode='fp32',
                    deprel_max_depth=rnd.choice([1, 2, 3]), deprel_directed=rnd.random() < 0.3,
                    deprel_self_loop=rnd.random() < 0.8, pooling=rnd.choice(['max', 'avg', 'sum']),
                    mlp_layers=rnd.choice([1, 2]))
        if mode == 'full_deprel':                   # the shared Linear needs in_dim == hidden_dim
            over.update(emb_dim=H, pos_dim=0, ner_dim=0)
        else:
            over.update(emb_dim=rnd.choice([3, 8]), pos_dim=rnd.choice([0, 2]), ner_dim=rnd.choice([0, 3]))
        opt = synth.tacred_opt(**over)
        batch = synth.make_batch(10 * chunk + trial, batch_size=rnd.choice([1, 3]), vocab_size=60,
                                 mean_len=rnd.choice([9, 20]))
        state = {k: torch.from_numpy(v) for k, v in weights.make_state(opt, 100 + trial).items()}
        trainer = GCNTrainer(dict(opt))
        trainer.model.load_state_dict(state)
        oracle = gcn_oracle.DenseClassifier(opt)
        oracle.load_state_dict(state)
        trainer.model.train()
        oracle.train()
        loss = trainer.update(batch)
        loss.backward()
        ref_loss, _ = oracle.loss(batch)
        ref_loss.backward()
        tag = dict(over, chunk=chunk, trial=trial)
        assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item()), tag
        got = dict(trainer.model.named_parameters())
        for key, p in oracle.named_parameters():
            if p.grad is not None:
                assert got[key].grad is not None and _rel(got[key].grad, p.grad) <= 1e-4, (tag, key)
