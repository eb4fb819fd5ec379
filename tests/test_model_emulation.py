"""CPU: the north-star hot path from its CUDA sources on the host (tests/emu) -- K1 prune->CSR, K5 embedding stage, K3
projection (fp32 FFMA mode), K2 aggregation, K4 pooling -- underneath the product's own GCNTrainer / GCNClassifier /
autograd Functions, against the outputs recorded from the UNMODIFIED reference (tests/golden/model.npz) and against the
oracle's gradients with identical injected dropout masks.  Nothing of the path is replaced except cuDNN's LSTM
(torch's CPU LSTM) for the C-GCN case and, where masks are injected, the embedding lookup (K5 draws its own Philox mask).
The `-m gpu` tests of test_gpu_parity.py run the same cases on the device, in both projection modes."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

import cases
import weights
from gcn_over_pruned_trees_b200 import _lib, ops
from gcn_over_pruned_trees_b200 import synth
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
from oracle import gcn_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
NAMES = ('gpt_prune_csr', 'gpt_embed_fwd', 'gpt_embed_bwd', 'gpt_linear_fwd_f32', 'gpt_linear_dgrad_f32',
         'gpt_linear_wgrad_f32', 'gpt_gcn_aggregate_fwd', 'gpt_gcn_aggregate_bwd', 'gpt_pool3_fwd', 'gpt_pool3_bwd')


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
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _setup(golden_adj, name, extra=None):
    over, source, wseed = cases.MODEL_CASES[name]
    over = dict(over, gemm_mode='fp32', **(extra or {}))     # K3's FFMA kernels; tcgen05 cannot be emulated
    if source[0] == 'split':
        over = dict(over, vocab_size=int(golden_adj['vocab_size']))
    batch = cases.make_case_batch(source, over, golden_adj)
    opt = synth.tacred_opt(**over)
    state = {k: torch.from_numpy(v) for k, v in weights.make_state(opt, wseed).items()}
    trainer = GCNTrainer(dict(opt))
    trainer.model.load_state_dict(state)
    oracle = gcn_oracle.DenseClassifier(opt)
    oracle.load_state_dict(state)
    return opt, batch, trainer, oracle


@pytest.mark.parametrize('name', sorted(cases.MODEL_CASES))
def test_hot_path_from_source_matches_reference_outputs(golden_adj, golden_model, name):
    opt, batch, trainer, _ = _setup(golden_adj, name)
    trainer.model.eval()
    with torch.no_grad():
        logits, h_out = trainer.model(list(batch[:-2]))
        loss = trainer._loss(logits, h_out, batch[-2])          # what update() adds to the forward (trainer.py:93-100)
    assert _rel(logits, golden_model['%s/logits' % name]) <= 1e-5
    assert _rel(h_out, golden_model['%s/h_out' % name]) <= 1e-5
    assert abs(loss.item() - float(golden_model['%s/eval_loss' % name])) <= 1e-5 * abs(loss.item())
    if cases.MODEL_CASES[name][1][0] == 'split':
        csr = trainer.model.gcn_model.last_csr                  # K1's adjacency, bit-exact against the reference's
        want = golden_adj['%s/adj_k%d' % (cases.MODEL_CASES[name][1][1], opt['prune_k'])].astype(np.float32)
        assert np.array_equal(csr.to_dense().numpy(), want)
        preds, probs, ploss = trainer.predict(batch)            # eval.py's call: unsorted labels, probabilities, CE
        assert preds == golden_model['%s/pred' % name].tolist()
        assert _rel(np.asarray(probs), golden_model['%s/probs' % name]) <= 1e-5
        assert abs(ploss - float(golden_model['%s/predict_loss' % name])) <= 1e-5 * abs(ploss)


def _compare_grads(trainer, oracle, tol):
    got = dict(trainer.model.named_parameters())
    checked = 0
    for key, p in oracle.named_parameters():
        if p.grad is None:
            assert got[key].grad is None or float(got[key].grad.abs().max()) == 0.0, key
            continue
        assert got[key].grad is not None, key
        assert _rel(got[key].grad, p.grad) <= tol, (key, _rel(got[key].grad, p.grad))
        checked += 1
    return checked


@pytest.mark.parametrize('name', cases.GRAD_CASES)
def test_hot_path_from_source_train_grads_match_oracle_with_injected_masks(golden_adj, name):
    opt, batch, trainer, oracle = _setup(golden_adj, name)
    B, T = batch[0].shape
    g = torch.Generator().manual_seed(99)
    in_dim = opt['emb_dim'] + opt['pos_dim'] + (opt['ner_dim'] if opt['dataset'] == 'tacred' else 0)

    def drop(shape, p):
        return (torch.rand(shape, generator=g) >= p).float() / (1 - p)
    masks = {'in': drop((B, T, in_dim), opt['input_dropout'])}
    if opt['rnn']:
        masks['rnn'] = drop((B, T, 2 * opt['rnn_hidden']), opt['rnn_dropout'])
    for l in range(opt['num_layers'] - 1):
        masks['gcn%d' % l] = drop((B, T, opt['hidden_dim']), opt['gcn_dropout'])
    oracle.train()
    ref_loss, _ = oracle.loss(batch, masks)
    ref_loss.backward()
    trainer.model.train()
    trainer.model.gcn_model.gcn.injected_masks = masks
    loss = trainer.update(batch)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle, 1e-4) >= 8


def test_hot_path_from_source_train_step_without_dropout_nothing_replaced(golden_adj):
    """train.py:213-227 with dropout probabilities 0: K1, K5 (forward + scatter backward), K3, K2 (forward, activation
    mask, backward), K4 from source under the autograd path; loss, every gradient, and the parameters after
    clip_grad_norm_ + SGD against the oracle's."""
    opt, batch, trainer, oracle = _setup(golden_adj, 'cfg2_synth_k1', dict(input_dropout=0.0, gcn_dropout=0.0))
    assert trainer.model.gcn_model.gcn.injected_masks is None
    trainer.model.train()
    oracle.train()
    optim = torch.optim.SGD([p for p in oracle.parameters() if p.requires_grad], lr=opt['lr'])
    ref_loss = gcn_oracle.train_step(oracle, optim, batch, opt['max_grad_norm'])
    trainer.optimizer.zero_grad()
    loss = trainer.update(batch)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert _compare_grads(trainer, oracle, 1e-4) >= 8       # (the oracle's .grad are the clipped ones: same coefficient)
    torch.nn.utils.clip_grad_norm_(trainer.model.parameters(), opt['max_grad_norm'])
    trainer.optimizer.step()
    want = dict(oracle.named_parameters())
    for key, p in trainer.model.named_parameters():
        assert _rel(p.detach(), want[key].detach()) <= 1e-5, key
