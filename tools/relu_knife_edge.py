#!/usr/bin/env python
"""How far do two correct fp32 training runs drift apart?  Trains the same model twice from parameters that differ by
2e-7 relative noise and reports the largest relative parameter deviation after 8 steps.  At TACRED batch sizes a
pre-activation of the output MLP within rounding of zero falls on either side of the ReLU, and that sentence's whole
contribution to the unit's gradient row flips with it: deviations of 1e-4..6e-3 are common, which
is why tests/dp_worker.py compares data-parallel and single-process runs in lockstep."""
import os
import sys

import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import synth
from gcn_over_pruned_trees_b200.engine import FusedTrainStep
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
over = dict(vocab_size=1500, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode='fp32')
def run(base, B, steps, perturb):
    batches = [synth.make_batch(base + i, batch_size=B, vocab_size=1500) for i in range(3)]
    torch.manual_seed(11)
    tr = GCNTrainer(synth.tacred_opt(**over)); tr.model.train()
    eng = FusedTrainStep(tr)
    if perturb:
        g = torch.Generator(device='cuda').manual_seed(perturb)
        for p in tr.model.parameters():
            p.data.mul_(1 + 2e-7 * torch.randn(p.shape, generator=g, device='cuda'))
    for s in range(steps):
        eng(batches[s % 3])
    torch.cuda.synchronize()
    return {k: v.detach().clone() for k, v in tr.model.state_dict().items()}
for base in (300, 700, 900, 1100):
    for B in (100, 200, 400):
        a = run(base, B, 8, 0)
        worst = 0.0
        for pert in (1, 2, 3):
            b = run(base, B, 8, pert)
            worst = max(worst, max(float((a[k] - b[k]).abs().max() / a[k].abs().max().clamp_min(1e-30)) for k in a))
        print('base %d B %d: worst rel deviation under 2e-7 parameter noise: %.2e' % (base, B, worst), flush=True)
