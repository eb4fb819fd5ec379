"""GPU tests of the reduced-precision projection mode (gemm_mode='bf16': tcgen05.mma kind::f16 with bf16 operands, fp32
accumulation) -- north_star: "logits and losses must match ... within 1e-5 relative in fp32, with the bf16 tolerance
stated separately".  The tolerances of this mode, stated here:

  * GEMM vs float64 on the bf16-ROUNDED operands: <= 2e-6 of the largest element (the products are exact in fp32, only
    the accumulation order differs) -- this pins the kernel's arithmetic, not the rounding
  * GEMM vs float64 on the unrounded operands: <= 1e-2 of the largest element (two roundings of 2^-9 per product)
  * model, eval mode: logits within 3e-2 of the fp32 mode's (relative to the largest logit), loss within 1e-2,
    at least 90 % of the predictions equal; a training run lowers the loss
"""
import numpy as np
import pytest
import torch

from gcn_over_pruned_trees_b200 import ops, synth
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.fixture(autouse=True)
def _restore():
    yield
    ops.gemm_persist_config(2, 65536)


@pytest.mark.parametrize('M,N,K,cg', [(2750, 200, 360, 2), (4800, 200, 200, 2), (50, 64, 40, 2), (300, 512, 512, 2),
                                      (70000, 512, 360, 2), (70000, 512, 360, 1), (66000, 200, 200, 2),
                                      (65536, 512, 512, 0)])
def test_bf16_projection_and_dgrad(M, N, K, cg):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / np.sqrt(K)
    dy = torch.randn(M, N, device=DEV, generator=g)
    ops.gemm_persist_config(cg, 65536)
    ws = ops.weight_prep(w, 'bf16')
    assert ws is not None and ws.dtype == torch.bfloat16
    assert torch.equal(ws[0].view(N, K), w.to(torch.bfloat16)) and torch.equal(ws[1].view(K, N), w.t().to(torch.bfloat16))
    y = ops.linear_fwd(x, w, 'bf16', ws)
    dx = ops.linear_dgrad(dy, w, 'bf16', ws)
    torch.cuda.synchronize()
    xr, wr, dyr = (t.to(torch.bfloat16).double() for t in (x, w, dy))
    assert _rel(y, xr @ wr.t()) <= 2e-6
    assert _rel(dx, dyr @ wr) <= 2e-6
    assert _rel(y, x.double() @ w.double().t()) <= 1e-2
    assert _rel(dx, dy.double() @ w.double()) <= 1e-2


def test_bf16_falls_back_to_fp32_where_tma_cannot_describe_the_weight():
    x = torch.randn(500, 330, device=DEV)                   # SemEval input width: bf16 rows of 660 B are not 16-byte multiples
    w = torch.randn(200, 330, device=DEV)
    assert ops.weight_prep(w, 'bf16') is None
    assert _rel(ops.linear_fwd(x, w, 'bf16'), x.double() @ w.double().t()) <= 2e-6


def test_model_in_bf16_mode_tracks_the_fp32_mode():
    batch = synth.make_batch(77, batch_size=50, vocab_size=800)
    outs = {}
    for mode in ('fp32', 'bf16'):
        torch.manual_seed(23)
        tr = GCNTrainer(synth.tacred_opt(vocab_size=800, cuda=True, gemm_mode=mode))
        tr.model.eval()
        with torch.no_grad():
            logits, _ = tr.model([t.to(DEV) for t in batch[:-2]])
        preds, probs, loss = tr.predict(batch)
        outs[mode] = (logits.cpu(), np.array(preds), loss)
    la, lb = outs['fp32'][0], outs['bf16'][0]
    assert float((la - lb).abs().max() / la.abs().max()) <= 3e-2
    assert float((la - lb).abs().max()) > 0                                  # it IS a different arithmetic
    assert abs(outs['fp32'][2] - outs['bf16'][2]) <= 1e-2 * abs(outs['fp32'][2])
    assert (outs['fp32'][1] == outs['bf16'][1]).mean() >= 0.9


@pytest.mark.parametrize('engine', ('update', 'train_step'))
def test_training_in_bf16_mode_lowers_the_loss(engine):
    torch.manual_seed(29)
    tr = GCNTrainer(synth.tacred_opt(vocab_size=600, cuda=True, gemm_mode='bf16', input_dropout=0.0, gcn_dropout=0.0))
    tr.model.train()
    batch = synth.make_batch(78, batch_size=50, vocab_size=600)
    losses = []
    for _ in range(8):
        if engine == 'train_step':
            losses.append(float(tr.train_step(batch)))
        else:
            tr.optimizer.zero_grad()
            loss = tr.update(batch)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(tr.model.parameters(), 5.0)
            tr.optimizer.step()
            losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
