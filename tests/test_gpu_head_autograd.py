"""GPU test of K6 (csrc/head.cu) on the per-op autograd path (ops.head_loss, GCNTrainer._forward_loss): the classifier head
and its loss in two launches, under autograd, against the nn.Linear / CrossEntropyLoss kernels it replaces
(/root/reference/model/gcn.py:64-68,122, model/trainer.py:94-100).

Tolerances: loss <= 1e-6 relative, every gradient <= 1e-5 relative of its largest element (fp32 both ways; the two sum in
different orders).
"""
import pytest
import torch

from gcn_over_pruned_trees_b200 import synth
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _grads(tr, batch, fused, scale=1.0):
    tr.fused_head = fused
    tr.fast_update = False                       # the per-op path is what is under test
    tr.optimizer.zero_grad(set_to_none=True)
    loss = tr.update(batch)
    (loss * scale).backward()
    return float(loss), {n: p.grad.detach().clone() for n, p in tr.model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize('over', [
    dict(),                                                         # train_gcn.sh: 2 mlp layers, pooling_l2 0.003
    dict(mlp_layers=1, pooling_l2=0.0),
    dict(mlp_layers=3, conv_l2=0.01, pooling='avg'),
    dict(adj_type='diagonal_deprel'),
    dict(rnn=True, hidden_dim=64, rnn_hidden=64),
])
def test_fused_head_on_the_autograd_path_equals_the_per_op_head(over):
    torch.manual_seed(3)
    opt = synth.tacred_opt(vocab_size=3000, cuda=True, gemm_mode='fp32', input_dropout=0.0, gcn_dropout=0.0,
                           rnn_dropout=0.0, **over)
    tr = GCNTrainer(opt)
    tr.model.train()
    batch = tuple(t.cuda() if torch.is_tensor(t) else t for t in synth.make_batch(77, batch_size=23, vocab_size=3000))
    ref_loss, ref = _grads(tr, batch, fused=False)
    loss, got = _grads(tr, batch, fused=True)
    assert abs(loss - ref_loss) <= 1e-6 * abs(ref_loss)
    assert set(got) == set(ref) and len(ref) >= 8
    for n, g in ref.items():
        assert _rel(got[n], g) <= 1e-5, n
    # the incoming gradient scales every gradient of the head (two losses summed, (loss * 0.5).backward(), ...)
    _, half = _grads(tr, batch, fused=True, scale=0.5)
    for n, g in ref.items():
        assert _rel(half[n], 0.5 * g) <= 1e-5, n


def test_fused_head_is_not_taken_in_eval_or_without_grad():
    opt = synth.tacred_opt(vocab_size=3000, cuda=True)
    tr = GCNTrainer(opt)
    batch = tuple(t.cuda() if torch.is_tensor(t) else t for t in synth.make_batch(5, batch_size=7, vocab_size=3000))
    tr.fast_update = False
    tr.model.eval()
    a = float(tr.update(batch))
    tr.fused_head = False
    b = float(tr.update(batch))
    assert abs(a - b) <= 1e-6 * abs(b)
