#!/usr/bin/env python
"""CPU baseline of the relation-aware adjacency modes: the oracle port of the reference's dense algorithm
(oracle/gcn_oracle.py, gcn.py:272-386) timed on this host's cores on the configurations `tools/config_sweep.py
--relation` times on the GPU.  One step = zero_grad + loss (train mode) + backward + clip + SGD step on 50 synthetic
TACRED-shaped sentences (numpy tree pruning + dense [B,T,T] adjacency + torch CPU).  Lives under tests/ because only
tests/, smoke() and bench.py's CPU legs may import oracle/.

    python tests/bench_relation_cpu.py [--steps 10]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import synth  # noqa: E402
from oracle import gcn_oracle                  # noqa: E402


def run(name, over, steps):
    torch.manual_seed(0)
    opt = synth.tacred_opt(vocab_size=50000, **over)
    model = gcn_oracle.DenseClassifier(opt)
    model.train()
    optim = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=opt['lr'])
    batches = [synth.make_batch(2000 + i, batch_size=50, vocab_size=50000) for i in range(4)]
    for b in batches[:2]:
        gcn_oracle.train_step(model, optim, b, opt['max_grad_norm'])
    t0 = time.perf_counter()
    for i in range(steps):
        gcn_oracle.train_step(model, optim, batches[i % len(batches)], opt['max_grad_norm'])
    dt = (time.perf_counter() - t0) / steps
    out = {'config': name, 'kind': 'port', 'cores': torch.get_num_threads(), 'ms_per_step': dt * 1e3,
           'sentences_per_s': 50 / dt, 'steps': steps}
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    run('gcn_full_deprel_d50_k1', dict(prune_k=1, adj_type='full_deprel', deprel_emb_dim=50, emb_dim=140), a.steps)
    run('gcn_diagonal_deprel_k1', dict(prune_k=1, adj_type='diagonal_deprel'), a.steps)
