#!/usr/bin/env python
"""Per-entry-point device time of one eager training step of the relation-aware modes (K10: full_deprel with 50 relation
slots on a 200-wide input, diagonal_deprel), B=50 TACRED-shaped sentences, k=1 -- where the step's 2.4 ms go."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops, synth  # noqa: E402
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer  # noqa: E402


def run(name, over, steps=10):
    torch.manual_seed(0)
    sys.stdout = open(os.devnull, 'w')
    tr = GCNTrainer(synth.tacred_opt(vocab_size=50000, cuda=True, gemm_mode='tf32x3', prune_k=1, **over))
    sys.stdout = sys.__stdout__
    tr.model.train()
    batch = tuple(t.cuda() if torch.is_tensor(t) else t for t in synth.make_batch(2000, batch_size=50, vocab_size=50000))

    def step():
        tr.optimizer.zero_grad()
        loss = tr.update(batch)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(tr.model.parameters(), 5.0)
        tr.optimizer.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ops.TIMER = ops.KernelTimer()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    summary = ops.TIMER.summary()
    ops.TIMER = None
    ms = a.elapsed_time(b) / steps
    print('%s: eager step %.3f ms (incl. launch gaps); library entry points:' % (name, ms))
    for k, (c, t) in sorted(summary.items(), key=lambda kv: -kv[1][1]):
        print('  %-28s %5.1f calls/step %8.1f us/call %6.1f us/step' % (k, c / steps, t / c * 1e3, t / steps * 1e3))
    g = [float(tr.train_step(batch)) for _ in range(8)]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    torch.cuda.synchronize()
    for x, y in ev:
        x.record()
        tr.train_step(batch)
        y.record()
    torch.cuda.synchronize()
    print('  graphed train_step: %.3f ms/step (%s)' % (sum(x.elapsed_time(y) for x, y in ev) / len(ev),
                                                       type(tr._graphed).__name__))


if __name__ == '__main__':
    run('full_deprel D=50', dict(adj_type='full_deprel', deprel_emb_dim=50, emb_dim=140))
    run('diagonal_deprel', dict(adj_type='diagonal_deprel'))
