#!/usr/bin/env python
"""Whole training step at the large synthetic shape (BASELINE.json configs[4]: 512-token trees, H=512, 2 layers):
FusedTrainStep launched eagerly, per-entry-point device time (at this size launch gaps are negligible)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops, synth  # noqa: E402
from gcn_over_pruned_trees_b200.engine import FusedTrainStep  # noqa: E402
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=4096)
    ap.add_argument('--T', type=int, default=512)
    ap.add_argument('--H', type=int, default=512)
    ap.add_argument('--k', type=int, default=-1)
    ap.add_argument('--steps', type=int, default=5)
    a = ap.parse_args()
    torch.manual_seed(0)
    opt = synth.tacred_opt(vocab_size=50000, cuda=True, hidden_dim=a.H, prune_k=a.k)
    tr = GCNTrainer(opt)
    tr.model.train()
    eng = FusedTrainStep(tr)
    batch = synth.make_batch_torch(5, a.B, a.T, device='cuda')
    inputs, labels = list(batch[:-2]), batch[-2]
    with torch.no_grad():
        for _ in range(2):
            eng._run(inputs, labels)
        torch.cuda.synchronize()
        ops.TIMER = ops.KernelTimer()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss, _ = eng._run(inputs, labels)
        e1.record()
        summary = ops.TIMER.summary()
        ops.TIMER = None
    ms = e0.elapsed_time(e1) / a.steps
    print('large step B=%d T=%d H=%d k=%d: %.2f ms/step  %.0f sentences/s  loss %.4f  peak mem %.1f GB' %
          (a.B, a.T, a.H, a.k, ms, a.B / ms * 1e3, float(loss), torch.cuda.max_memory_allocated() / 1e9))
    for k, (c, t) in sorted(summary.items(), key=lambda kv: -kv[1][1]):
        print('  %-34s %5.1f calls/step %9.3f ms/call %5.1f%%' % (k, c / a.steps, t / c, 100 * t / a.steps / ms))


if __name__ == '__main__':
    main()
